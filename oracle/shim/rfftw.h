/* Local stand-in for FFTW 2.x <rfftw.h>, TEST INFRASTRUCTURE ONLY.
 *
 * The reference's double-precision path (fft.c:171,196,302; mtm.c:96,171,196,250,
 * compiled with -DHAVE_LIBRFFTW) calls three entry points of FFTW 2 (version not
 * pinned by the reference: configure.in:17-18 only probes -lfftw/-lrfftw).  FFTW 2
 * is not in this image, so the golden build links this shim instead.  The contract
 * at that boundary is the mathematical DFT in FFTW's "half-complex" layout:
 *   out[k] = Re X[k] (0 <= k <= n/2),  out[n-k] = Im X[k] (0 < k < n/2),
 *   X[k] = sum_j in[j] exp(-2 pi i j k / n), un-normalised.
 */
#ifndef ORACLE_RFFTW_SHIM_H
#define ORACLE_RFFTW_SHIM_H

typedef double fftw_real;
typedef struct oracle_rfftw_plan_s *fftw_plan;
typedef fftw_plan rfftw_plan;
typedef enum { FFTW_REAL_TO_COMPLEX = -1, FFTW_COMPLEX_TO_REAL = 1 } fftw_direction;
#define FFTW_ESTIMATE 0
#define FFTW_MEASURE 1

rfftw_plan rfftw_create_plan(int n, fftw_direction dir, int flags);
void rfftw_one(rfftw_plan plan, fftw_real *in, fftw_real *out);
void rfftw_destroy_plan(rfftw_plan plan);

#endif
