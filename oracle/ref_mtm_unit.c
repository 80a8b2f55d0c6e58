/* The reference's mtm.c as its own translation unit WITH an accessor for its file-static
 * F-test buffer.  TEST INFRASTRUCTURE ONLY (oracle).
 *
 * mtm_do computes Thomson's harmonic F-test into `static float *ftest` (mtm.c:59,165-174,
 * 204-210,222-233) and nothing in glfer ever reads it.  To pin the product's live F-test
 * output against the reference's own arithmetic the unmodified source is #included here
 * (found through -I$(REF)), so this unit can hand the static out.  In the FFTW (double)
 * build `mu` is written by rfftw_one (mtm.c:170-171) and the statistic is well defined;
 * in the no-FFTW float build the FFT lands in inbuf_fft and `mu` is never written.
 */
#include "mtm.c"

const float *refh_mtm_ftest_buffer(void) { return ftest; }
