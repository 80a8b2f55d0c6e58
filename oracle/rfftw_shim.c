/* Accurate double-precision real DFT behind the FFTW 2 rfftw entry points.
 * TEST INFRASTRUCTURE ONLY (oracle); never linked into the product library.
 *
 * Power-of-two n: iterative radix-2 complex FFT on long double twiddles computed
 * directly (no recurrences), so the result is the DFT to ~1e-16 relative.
 * Other n: direct O(n^2) DFT in long double (only used for tiny test sizes).
 */
#include <math.h>
#include <stdlib.h>
#include "rfftw.h"

struct oracle_rfftw_plan_s {
  int n;
  int pow2;
  long double *cs; /* cos(2 pi k / n), k < n */
  long double *sn; /* sin(2 pi k / n) */
  long double *wr, *wi; /* work */
};

rfftw_plan rfftw_create_plan(int n, fftw_direction dir, int flags)
{
  (void) dir; (void) flags;
  struct oracle_rfftw_plan_s *p = calloc(1, sizeof *p);
  p->n = n;
  p->pow2 = (n > 0) && ((n & (n - 1)) == 0);
  p->cs = malloc(sizeof(long double) * n);
  p->sn = malloc(sizeof(long double) * n);
  p->wr = malloc(sizeof(long double) * n);
  p->wi = malloc(sizeof(long double) * n);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < n; k++) {
    p->cs[k] = cosl(two_pi * k / n);
    p->sn[k] = sinl(two_pi * k / n);
  }
  return p;
}

void rfftw_destroy_plan(rfftw_plan p)
{
  if (!p) return;
  free(p->cs); free(p->sn); free(p->wr); free(p->wi); free(p);
}

static void fft_pow2(struct oracle_rfftw_plan_s *p)
{
  int n = p->n;
  long double *xr = p->wr, *xi = p->wi;
  /* bit reversal */
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      long double t = xr[i]; xr[i] = xr[j]; xr[j] = t;
      t = xi[i]; xi[i] = xi[j]; xi[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1, step = n / len;
    for (int base = 0; base < n; base += len) {
      for (int k = 0; k < half; k++) {
        /* forward transform: exp(-2 pi i k / len) */
        long double c = p->cs[k * step], s = -p->sn[k * step];
        long double ur = xr[base + k], ui = xi[base + k];
        long double vr = xr[base + k + half] * c - xi[base + k + half] * s;
        long double vi = xr[base + k + half] * s + xi[base + k + half] * c;
        xr[base + k] = ur + vr; xi[base + k] = ui + vi;
        xr[base + k + half] = ur - vr; xi[base + k + half] = ui - vi;
      }
    }
  }
}

void rfftw_one(rfftw_plan p, fftw_real *in, fftw_real *out)
{
  int n = p->n;
  if (p->pow2 && n > 1) {
    for (int i = 0; i < n; i++) { p->wr[i] = in[i]; p->wi[i] = 0.0L; }
    fft_pow2(p);
  } else {
    for (int k = 0; k <= n / 2; k++) {
      long double ar = 0, ai = 0;
      for (int j = 0; j < n; j++) {
        int idx = (int) (((long long) j * k) % n);
        ar += in[j] * p->cs[idx];
        ai -= in[j] * p->sn[idx];
      }
      p->wr[k] = ar; p->wi[k] = ai;
    }
  }
  out[0] = (double) p->wr[0];
  for (int k = 1; k < (n + 1) / 2; k++) {
    out[k] = (double) p->wr[k];
    out[n - k] = (double) p->wi[k];
  }
  if (n % 2 == 0 && n > 1) out[n / 2] = (double) p->wr[n / 2];
}
