// tables.hpp -- host-side (double precision) builders for the constant tables the
// gram kernels read: per-pass twiddles, the real-FFT split factors and the plan
// geometry.  Values are computed directly from the angle (never by recurrence, the
// accuracy limiter of the reference's fft_radix2.c:127-140) and rounded once to float.
#pragma once
#include <cmath>
#include <vector>
#include "fft_core.cuh"

namespace glb {

inline void unit_root(long long e, long long L, double &c, double &s) {
  // exp(-2 pi i e / L) with exact octant handling of the trivial angles
  e %= L;
  if (e < 0) e += L;
  if (e == 0) { c = 1; s = 0; return; }
  if (4 * e == L) { c = 0; s = -1; return; }
  if (2 * e == L) { c = -1; s = 0; return; }
  if (4 * e == 3 * L) { c = 0; s = 1; return; }
  const double a = -2.0 * M_PI * (double) e / (double) L;
  c = std::cos(a);
  s = std::sin(a);
}

template <int M, int P>
inline void fill_pass(std::vector<float2> &tw) {
  if constexpr (P > 0 && P < Plan<M>::NP) {
    constexpr int R = PlanRadix<M, P>::R, Ns = PlanRadix<M, P>::Ns;
    for (int r = 1; r < R; r++)
      for (int k = 0; k < Ns; k++) {
        double c, s;
        unit_root((long long) k * r, (long long) Ns * R, c, s);
        tw[TwOffset<M, P>::value + (r - 1) * Ns + k] = make_float2((float) c, (float) s);
      }
  }
}

template <int M>
inline std::vector<float2> build_twiddles() {
  std::vector<float2> tw(TwTotal<M>::value > 0 ? TwTotal<M>::value : 1);
  fill_pass<M, 1>(tw);
  fill_pass<M, 2>(tw);
  fill_pass<M, 3>(tw);
  return tw;
}

// roots[k] = exp(-2 pi i k / M), k = 0..M-1
inline std::vector<float2> build_roots(int M) {
  std::vector<float2> v(M);
  for (int k = 0; k < M; k++) {
    double c, s;
    unit_root(k, M, c, s);
    v[k] = make_float2((float) c, (float) s);
  }
  return v;
}

// vtab[k] = -i exp(-2 pi i k / N), N = 2M, k = 0..M-1
inline std::vector<float2> build_vtab(int M) {
  std::vector<float2> v(M);
  for (int k = 0; k < M; k++) {
    double c, s;
    unit_root(k, 2LL * M, c, s);         // c - i*(-s): exp = c + i s (s already negative)
    // -i * (c + i s) = s - i c
    v[k] = make_float2((float) s, (float) (-c));
  }
  return v;
}

}  // namespace glb
