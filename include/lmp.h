/* lmp.h -- LMP estimator interface of libglfer_b200 (drop-in).
 *
 * Same names, struct layout and ownership as the reference's lmp.h:37-50 so source.c:155-157,
 * 292-293,394-400 link unchanged.  lmp_do runs the rectangular-window FFT of the frame
 * (prepare_audio's RA9MB / limiter output is overwritten by the raw frame, lmp.c:112-114) and
 * the per-bin detector statistic over the ring of the last `avg` PSD rows (lmp.c:131-160) on
 * the GPU.  The ring index restarts at 0 with every lmp_init (the reference keeps it in a
 * function-static that no init resets, lmp.c:102: a second init with a smaller `avg` indexes
 * its ring out of bounds); several instances may coexist (lmp.c:47-49 are file-static).
 */
#ifndef GLFER_B200_LMP_H
#define GLFER_B200_LMP_H

#include "fft.h"

#ifdef __cplusplus
extern "C" {
#endif

/* replaces lmp.h:37-46 */
typedef struct {
  fft_params_t fft;
  int avg;           /* rows in the ring (opt.lmp_av, source.c:399); >= 2 */
  double **window;   /* unused by lmp.c */
  double *sig;       /* unused by lmp.c */
  float w;           /* unused by lmp.c */
  int kmax;          /* unused by lmp.c */
} lmp_params_t;

/* replaces lmp.h:49-51 */
void lmp_init(lmp_params_t *params);
void lmp_do(float *audio_buf, float *psd_buf, float *phase_buf, lmp_params_t *params);
void lmp_close(lmp_params_t *params);

#ifdef __cplusplus
}
#endif
#endif
