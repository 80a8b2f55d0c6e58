/* mtm.h -- Thomson multitaper estimator interface of libglfer_b200 (drop-in).
 *
 * Same names and ownership as the reference's mtm.h:36-49 so source.c:147-149,343-350
 * link unchanged.  mtm_init computes the DPSS tapers on the host in double (a fresh
 * implementation of the Gauss-Legendre method of g-l_dpss.c:288-347) into the same
 * NR-style 1-offset matrix window[1..n][0..kmax] (mtm.c:118) and uploads a float copy
 * with the 1/lambda_k weights folded in; mtm_do runs all kmax+1 tapered FFTs of the
 * frame and their weighted sum (mtm.c:189-220) in one GPU kernel.
 * Unlike the reference (file-static state, mtm.c:47-60) several instances may coexist.
 */
#ifndef GLFER_B200_MTM_H
#define GLFER_B200_MTM_H

#include "fft.h"

#ifdef __cplusplus
extern "C" {
#endif

/* replaces mtm.h:36-44 */
typedef struct {
  fft_params_t fft;
  double **window;   /* window[i + 1][k]: taper k at sample i */
  double *sig;       /* sig[k] = lambda_k - 1 (g-l_dpss.c:342-344) */
  float w;           /* N*W, the time-bandwidth product (g-l_dpss.c:295-297) */
  int kmax;          /* kmax + 1 tapers are used (mtm.c:189) */
} mtm_params_t;

/* replaces mtm.h:47-49 */
void mtm_init(mtm_params_t *params);
void mtm_do(float *audio_buf, float *psd_buf, float *phase_buf, mtm_params_t *params);
void mtm_close(mtm_params_t *params);

/* extension: nblocks hop blocks per call, = nblocks mtm_do calls (see fft_do_batch in fft.h) */
void mtm_do_batch(float *audio_blocks, int nblocks, float *psd_rows, mtm_params_t *params);

/* replaces g-l_dpss.h:23; v is an NR-style matrix v[1..n][0..kmax] */
int gl_dpss(int nmax, int kmax, int n, double w, double **v, double *sig, int *totit);

#ifdef __cplusplus
}
#endif
#endif
