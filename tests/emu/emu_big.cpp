// CPU emulation of the 32-points-per-thread FFT core (fft_big.cuh, the same header the GPU compiles):
// every thread of one frame phase by phase with the exchange buffer as a plain array, checked against a
// double-precision FFT (recursive radix-2 on long double twiddles).  Catches index / twiddle / layout
// mistakes without a GPU.  Prints "M maxrel_psd maxabs_spec bad".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <complex>
#include "../../glfer_b200/csrc/tables.hpp"
#include "../../glfer_b200/csrc/fft_big.cuh"

using namespace glb;
typedef std::complex<long double> cld;

static void fft_ref(std::vector<cld> &a) {
  const size_t n = a.size();
  if (n == 1) return;
  std::vector<cld> e(n / 2), o(n / 2);
  for (size_t i = 0; i < n / 2; i++) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
  fft_ref(e);
  fft_ref(o);
  const long double pi = 3.14159265358979323846264338327950288L;
  for (size_t k = 0; k < n / 2; k++) {
    const cld w = std::polar(1.0L, -2.0L * pi * (long double) k / (long double) n) * o[k];
    a[k] = e[k] + w;
    a[k + n / 2] = e[k] - w;
  }
}

template <int M> int check(unsigned seed, bool tab = false) {
  constexpr int N = 2 * M, T = Big<M>::T;
  std::vector<float> x(N);
  srand(seed);
  for (auto &v : x) v = (float) rand() / RAND_MAX - 0.5f;
  auto roots = build_roots(M);
  auto vtab = build_vtab(M);
  std::vector<float2> tw1(Big<M>::TW1 > 0 ? Big<M>::TW1 : 1);
  for (int r = 1; r < Big<M>::R1; r++)
    for (int k = 0; k < 32; k++) tw1[(r - 1) * 32 + k] = roots[(k * r * 16) % M];     // what the kernel prologue does
  std::vector<std::vector<float2>> regs(T, std::vector<float2>(kBP));
  std::vector<float2> buf(Big<M>::BUF);
  std::vector<int> written(Big<M>::BUF, 0);
  for (int t = 0; t < T; t++)
    for (int q = 0; q < kBP; q++) regs[t][q] = make_float2(x[2 * (t + T * q)], x[2 * (t + T * q) + 1]);
  for (int t = 0; t < T; t++) big_pass0(regs[t].data());
  for (auto &b : buf) b = make_float2(NAN, NAN);
  for (int t = 0; t < T; t++) big_scatter0<M>(regs[t].data(), t, buf.data());
  for (int t = 0; t < T; t++) big_load1<M>(regs[t].data(), t, buf.data());
  for (int t = 0; t < T; t++) big_pass1<M>(regs[t].data(), t, tw1.data());
  for (auto &b : buf) b = make_float2(NAN, NAN);
  for (int t = 0; t < T; t++) big_scatter1<M>(regs[t].data(), t, buf.data());
  std::vector<BigLast> L(T);
  for (int t = 0; t < T; t++) big_load_last<M>(L[t], t, roots.data(), vtab.data());
  for (int t = 0; t < T; t++) big_load2<M>(regs[t].data(), t, buf.data());
  std::vector<float2> tw2(15 * T);
  for (int r = 1; r < 16; r++)
    for (int t = 0; t < T; t++) tw2[(r - 1) * T + t] = roots[t * r];
  for (int t = 0; t < T; t++) {
    if (tab) big_pass2_tab<M>(regs[t].data(), t, tw2.data());
    else big_pass2<M>(regs[t].data(), t, L[t]);
  }
  std::vector<double> psd(M + 1, -1.0);
  std::vector<float2> spec(M + 1);
  std::vector<int> hits(M + 1, 0);
  for (int t = 0; t < T; t++)
    big_emit<M>(regs[t].data(), t, L[t], [&](int slot, float2 a, bool conj) {
      const int bin = big_slot_bin<M>(t, slot);
      hits[bin]++;
      psd[bin] = 0.25 * (double) norm2(a);
      spec[bin] = make_float2(0.5f * a.x, conj ? -0.5f * a.y : 0.5f * a.y);
    });
  std::vector<cld> ref(N);
  for (int n = 0; n < N; n++) ref[n] = cld(x[n], 0);
  fft_ref(ref);
  double rms = 0;
  for (int k = 0; k <= M; k++) rms += (double) std::norm(ref[k]);
  rms = std::sqrt(rms / (M + 1));
  int bad = 0;
  double maxrel = 0, maxabs = 0;
  for (int k = 0; k <= M; k++) {
    if (hits[k] != 1) bad++;
    const double r = (double) std::norm(ref[k]);
    const double rel = std::fabs(psd[k] - r) / (r + 1e-6 * rms * rms);
    if (rel > maxrel) maxrel = rel;
    const double ea = std::hypot((double) spec[k].x - (double) ref[k].real(), (double) spec[k].y - (double) ref[k].imag()) / rms;
    if (ea > maxabs) maxabs = ea;
  }
  printf("%d %.3e %.3e %d\n", M, maxrel, maxabs, bad);
  return bad != 0 || !(maxrel < 1e-4) || !(maxabs < 1e-5);
}

int main() {
  int rc = 0;
  rc |= check<2048>(1);
  rc |= check<4096>(2);
  rc |= check<8192>(3);
  rc |= check<16384>(4);
  rc |= check<8192>(5, true);
  rc |= check<16384>(6, true);
  return rc;
}
