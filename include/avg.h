/* avg.h -- per-bin frame averaging interface of libglfer_b200 (drop-in).
 *
 * Same names, struct layout and call sequence as the reference's avg.h:28-43, as used by
 * g_main.c:1153-1183, source.c:311-312 and g_options.c:329-330.  The sliding sums run
 * on the GPU over a device-resident ring of the last avgdepth PSD rows; avg[] (the only
 * array the callers read, g_main.c:1173,1193,1199) is filled on every update, effdepth
 * is kept as the reference keeps it.  cum[] and avgarray[][] are allocated (so
 * alloc/delete pair up as in avg.c:38-78) but the running sums live on the device.
 */
#ifndef GLFER_B200_AVG_H
#define GLFER_B200_AVG_H

#ifdef __cplusplus
extern "C" {
#endif

/* replaces avg.h:28-36 */
typedef struct {
  int avgwidth;        /* spectrum width to average */
  int avgdepth;        /* number of spectra to average */
  int effdepth;        /* number of spectra averaged so far */
  double *avg;         /* current average array */
  double *cum;         /* (device-resident in this implementation) */
  double **avgarray;   /* (device-resident in this implementation) */
} avg_data_t;

/* replaces avg.h:38-43 */
extern void init_avg(avg_data_t *avgdata);
extern void alloc_avg(avg_data_t *avgdata, int width, int depth);
extern void delete_avg(avg_data_t *avgdata);
extern double update_avg_plain(avg_data_t *avgdata, int N, float *psd, int minbin, int maxbin, int *peakbin);
extern double update_avg_sumextreme(avg_data_t *avgdata, int N, float *psd, int max0, int minbin, int maxbin,
                                    int *peakbin);
extern double update_avg_sumavg(avg_data_t *avgdata, int N, float *psd, int max0, int minbin, int maxbin,
                                int *peakbin, double *variance);

#ifdef __cplusplus
}
#endif
#endif
