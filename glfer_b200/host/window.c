/* window.c -- FFT window tables (host, init-time).
 *
 * Same eight shapes, denominators, float storage and unit-energy normalisation as the
 * reference's compute_window (fft.c:309-360), including its arithmetic types: shapes in
 * double stored to float, Kaiser's t / alpha / radicand in float (fft.c:312-313,344-346),
 * the power sum in a float accumulator (fft.c:314,353-356).  The Kaiser I0 is the
 * Abramowitz & Stegun 9.8.1 / 9.8.2 polynomial pair the reference uses (util.c:222-237).
 */
#include <math.h>
#include <stdlib.h>
#include "glb_host.h"

static double horner(const double *c, int n, double y)
{
  double acc = c[n - 1];
  for (int i = n - 2; i >= 0; i--) acc = c[i] + y * acc;
  return acc;
}

double glb_bessel_i0(double x)
{
  static const double small_c[] = { 1.0, 3.5156229, 3.0899424, 1.2067492, 0.2659732, 0.360768e-01, 0.45813e-02 };
  static const double large_c[] = { 0.39894228, 0.1328592e-01, 0.225319e-02, -0.157565e-02, 0.916281e-02,
                                    -0.2057706e-01, 0.2635537e-01, -0.1647633e-01, 0.392377e-02 };
  const double ax = fabs(x);
  if (ax < 3.75) {
    double y = x / 3.75;
    y *= y;
    return horner(small_c, 7, y);
  }
  return (exp(ax) / sqrt(ax)) * horner(large_c, 9, 3.75 / ax);
}

void glb_window_table(int n, int window_type, float *w)
{
  const double nm1 = n - 1.0;
  const float kt = (n - 1.0) / 2.0;                /* Kaiser centre, a C float in the reference */
  const float kalpha = 6.0 / kt;
  for (int i = 0; i < n; i++) {
    const double c = (2.0 * i - n + 1.0) / nm1;    /* centred abscissa in [-1, 1] */
    double v;
    switch (window_type) {
    case HANNING_WINDOW:
      v = 0.5 - 0.5 * cos(2.0 * M_PI * i / nm1);
      break;
    case BLACKMAN_WINDOW:
      v = 0.42 - 0.5 * cos(2.0 * M_PI * i / nm1) + 0.08 * cos(4.0 * M_PI * i / nm1);
      break;
    case GAUSSIAN_WINDOW:
      v = exp(-1.0 * (2.0 * i - n + 1.0) * (2.0 * i - n + 1.0) / (nm1 * nm1));
      break;
    case WELCH_WINDOW:
      v = 1.0 - c * c;
      break;
    case BARTLETT_WINDOW:
      v = 1.0 - fabs(c);
      break;
    case HAMMING_WINDOW:
      v = 0.54 - 0.46 * cos(2.0 * M_PI * i / nm1);
      break;
    case KAISER_WINDOW: {
      const float d = i - kt;
      const float rad = kt * kt - d * d;           /* float arithmetic, as fft.c:346 */
      const float at = kalpha * kt;
      v = glb_bessel_i0(kalpha * sqrt(rad)) / glb_bessel_i0(at);
      break;
    }
    case RECTANGULAR_WINDOW:
    default:
      v = 1.0;
    }
    w[i] = v;
  }
  float pwr = 0.0f;
  for (int i = 0; i < n; i++) pwr += w[i] * w[i];
  const double norm = sqrt(pwr);
  for (int i = 0; i < n; i++) w[i] /= norm;
}
