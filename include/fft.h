/* fft.h -- periodogram estimator interface of libglfer_b200 (drop-in for the reference).
 *
 * Same type, field and function names, argument meaning and ownership rules as the
 * reference's fft.h so that source.c:141-146,320-325, glfer.c:143, g_main.c:1109 and
 * g_scope.c:186-197 compile and link unchanged against this library instead of
 * fft.c + fft_radix2.c.  Both layouts of the reference exist:
 *   default               the no-FFTW variant (fft.h:51-63: float buffers, outbuf aliasing inbuf_fft,
 *                         fft.c:178-180)                                   -> libglfer_b200.so
 *   -DGLFER_FFTW_LAYOUT   the HAVE_LIBRFFTW variant (fft.h:36-48: `fftw_plan plan` first, fftw_real =
 *                         double buffers, outbuf a buffer of its own, fft.c:171-176; g_scope.c:189-197
 *                         reads them as double *)                          -> libglfer_b200_fftw.so
 * Compile the caller with the same definition as the reference build it replaces.
 *
 * What changes underneath: fft_do() runs window multiply + real FFT + |X|^2 on the
 * GPU (one frame per call here; the batched path is glfer_b200.h).  There is no CPU
 * implementation: with no CUDA device the calls print an error and abort the way the
 * reference does on allocation failure (fft.c:249-252).
 */
#ifndef GLFER_B200_FFT_H
#define GLFER_B200_FFT_H

#ifdef __cplusplus
extern "C" {
#endif

#ifdef GLFER_FFTW_LAYOUT
/* replaces fft.h:36-48 (and the two FFTW 2 types it needs: `typedef double fftw_real`, an opaque plan) */
typedef double fftw_real;
typedef void *fftw_plan;
typedef fftw_real glfer_real;
typedef struct {
  fftw_plan plan;       /* never dereferenced by callers; the library keeps its engine handle here */
  fftw_real *inbuf_audio;
  fftw_real *inbuf_fft;
  fftw_real *outbuf;    /* half-complex spectrum in a buffer of its own (fft.c:173) */
#else
typedef float glfer_real;
/* replaces fft.h:51-63 */
typedef struct {
  float *inbuf_audio;   /* N samples: (N - hop) of history + hop new ones (fft.c:98-113) */
  float *inbuf_fft;     /* windowed frame as handed to the FFT (fft.c:127-156) */
  float *outbuf;        /* half-complex spectrum, aliases inbuf_fft (fft.c:180) */
#endif
  int n;                /* FFT size */
  float *window;        /* unit-energy window, N floats (fft.c:309-360) */
  int window_type;      /* enum below */
  float overlap;        /* fraction of overlap between blocks */
  float a;              /* RA9MB non-linear processing parameter (fft.c:127) */
  int limiter;          /* fft.c:151 */
  int sub_mean;         /* set by fft_init from opt.autoscale (fft.c:186) */
} fft_params_t;

/* replaces fft.h:67 */
enum { HANNING_WINDOW = 0, BLACKMAN_WINDOW, GAUSSIAN_WINDOW, WELCH_WINDOW, BARTLETT_WINDOW,
       RECTANGULAR_WINDOW, HAMMING_WINDOW, KAISER_WINDOW };

/* replaces fft.h:69-75; table and count exported as fft.c:48-60 does */
typedef struct _fft_window_t fft_window_t;
struct _fft_window_t {
  char *name;
  int type;
};
extern fft_window_t fft_windows[];
extern int num_fft_windows;

/* replaces fft.h:77-81 */
void prepare_audio(float *audio_buf, fft_params_t *params);
void fft_init(fft_params_t *params);
void fft_do(float *audio_buf, fft_params_t *params);
void fft_psd(float *psd_buf, float *phase_buf, fft_params_t *params);
void fft_close(fft_params_t *params);

/* Extension (not in the reference): nblocks consecutive hop blocks per call -- exactly the sequence
 *   for b: fft_do(audio_blocks + b * hop, params); fft_psd(psd_rows + b * (n/2+1), NULL, params); glfer.first_buffer = FALSE;
 * (hop = (int)(n * (1.0 - overlap)), fft.c:70) with the launch and copy latency of the GPU paid once per call
 * instead of once per block.  State left in params (history, outbuf, block means removed in place from
 * audio_blocks) is that of the last single call; glfer.first_buffer is cleared after the first block. */
void fft_do_batch(float *audio_blocks, int nblocks, float *psd_rows, fft_params_t *params);

/* replaces fft.h:83 (display/AGC statistics of one PSD row) */
void compute_floor(float *psd_buf, int n, float *sig_pwr_p, float *floor_pwr_p, float *peak_pwr_p,
                   unsigned int *peak_bin_p);

/* window table generator (fft.c:63, non-static in the reference) */
void compute_window(fft_params_t *params);

#ifdef __cplusplus
}
#endif
#endif
