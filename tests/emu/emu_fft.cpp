// CPU emulation of the gram kernel's FFT phases (same fft_core.cuh the GPU compiles):
// runs every thread of one frame phase by phase with the smem buffer as a plain
// array, and checks |X|^2 and X against a long-double DFT.  Host-logic test: catches
// index/twiddle/swizzle mistakes without a GPU.  Prints "M maxrel_psd maxabs_spec".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../glfer_b200/csrc/tables.hpp"

using namespace glb;

template <int M, int P> struct MidPasses {
  static void run(std::vector<std::vector<float2>> &regs, float2 *buf, const float2 *tw) {
    constexpr int T = M / kPoints;
    if constexpr (P < Plan<M>::NP - 1) {
      for (int t = 0; t < T; t++) pass_load<M>(regs[t].data(), t, buf);
      for (int t = 0; t < T; t++) pass_store<M, P>(regs[t].data(), t, buf, tw);
      MidPasses<M, P + 1>::run(regs, buf, tw);
    }
  }
};

template <int M, int P> struct MidPassesRT {
  static void run(std::vector<std::vector<float2>> &regs, float2 *buf, std::vector<TwRegs> &tr) {
    constexpr int T = M / kPoints;
    if constexpr (P < Plan<M>::NP - 1) {
      for (int t = 0; t < T; t++) pass_load<M>(regs[t].data(), t, buf);
      for (int t = 0; t < T; t++) pass_compute_rt<M, P>(regs[t].data(), tr[t]);
      for (int t = 0; t < T; t++) pass_scatter<M, P>(regs[t].data(), t, buf);
      MidPassesRT<M, P + 1>::run(regs, buf, tr);
    }
  }
};

template <int M> int check(unsigned seed, bool rt = false) {
  constexpr int N = 2 * M, T = M / kPoints;
  std::vector<float> x(N);
  srand(seed);
  for (auto &v : x) v = (float) rand() / RAND_MAX - 0.5f;
  auto tw = build_twiddles<M>();
  auto vtab = build_vtab(M);
  std::vector<std::vector<float2>> regs(T, std::vector<float2>(kPoints));
  // 16-byte aligned (pass 0 uses 128-bit stores)
  std::vector<float4> buf4((BufSize<M>::value + 1) / 2 + 1);
  float2 *bufp = reinterpret_cast<float2 *>(buf4.data());
  // load phase: element q of thread t is z[t + T q]
  for (int t = 0; t < T; t++)
    for (int q = 0; q < kPoints; q++) regs[t][q] = make_float2(x[2 * (t + T * q)], x[2 * (t + T * q) + 1]);
  std::vector<TwRegs> tr(T);
  if (rt) {
    // register-twiddle variant (what the kernels run for the periodogram path)
    for (int t = 0; t < T; t++) load_tw_regs<M>(tr[t], t, tw.data(), vtab.data());
    for (int t = 0; t < T; t++) pass_compute_rt<M, 0>(regs[t].data(), tr[t]);
    for (int t = 0; t < T; t++) pass_scatter<M, 0>(regs[t].data(), t, bufp);
    MidPassesRT<M, 1>::run(regs, bufp, tr);
    for (int t = 0; t < T; t++) last_pass_rt<M>(regs[t].data(), t, bufp, tw.data(), tr[t]);
  } else {
    for (int t = 0; t < T; t++) pass_store<M, 0>(regs[t].data(), t, bufp, tw.data());
    MidPasses<M, 1>::run(regs, bufp, tw.data());
    for (int t = 0; t < T; t++) last_pass<M>(regs[t].data(), t, bufp, tw.data());
  }
  std::vector<double> psd(M + 1, -1.0);
  std::vector<float2> spec(M + 1);
  std::vector<int> hits(M + 1, 0);
  int bin_err = 0;
  auto sink = [&](int t) {
    return [&, t](int slot, float2 a, bool conj) {
      const int bin = slot_bin<M>(t, slot);
      if (slot >= slot_count<M>(t)) bin_err++;
      hits[bin]++;
      psd[bin] = 0.25 * (double) norm2(a);
      spec[bin] = make_float2(0.5f * a.x, conj ? -0.5f * a.y : 0.5f * a.y);
    };
  };
  for (int t = 0; t < T; t++) {
    if (rt) emit_bins_rt<M>(regs[t].data(), t, tr[t], sink(t));
    else emit_bins<M>(regs[t].data(), t, vtab.data(), sink(t));
  }
  double maxrel = 0, maxabs = 0, ref_rms = 0;
  std::vector<long double> xr(M + 1), xi(M + 1);
  for (int k = 0; k <= M; k++) {
    long double sr = 0, si = 0;
    for (int n = 0; n < N; n++) {
      long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double) (((long long) n * k) % N) / N;
      sr += x[n] * cosl(a);
      si += x[n] * sinl(a);
    }
    xr[k] = sr; xi[k] = si;
    ref_rms += (double) (sr * sr + si * si);
  }
  ref_rms = std::sqrt(ref_rms / (M + 1));
  int bad = bin_err;
  for (int k = 0; k <= M; k++) {
    if (hits[k] != 1) bad++;
    double ref = (double) (xr[k] * xr[k] + xi[k] * xi[k]);
    double rel = std::fabs(psd[k] - ref) / (ref + 1e-6 * ref_rms * ref_rms);
    if (rel > maxrel) maxrel = rel;
    double ea = std::hypot((double) spec[k].x - (double) xr[k], (double) spec[k].y - (double) xi[k]) / ref_rms;
    if (ea > maxabs) maxabs = ea;
  }
  printf("%d %.3e %.3e %d %s\n", M, maxrel, maxabs, bad, rt ? "rt" : "table");
  return bad != 0 || maxrel > 1e-4 || maxabs > 1e-5;
}

int main(int argc, char **argv) {
  const bool full = argc > 1;
  int rc = 0;
  rc |= check<16>(1);
  rc |= check<32>(2);
  rc |= check<64>(3);
  rc |= check<128>(4);
  rc |= check<256>(5);
  rc |= check<512>(6);
  rc |= check<1024>(7);
  rc |= check<2048>(8);
  rc |= check<16>(21, true);
  rc |= check<64>(22, true);
  rc |= check<128>(23, true);
  rc |= check<256>(24, true);
  rc |= check<512>(25, true);
  rc |= check<1024>(26, true);
  rc |= check<2048>(27, true);
  if (full) rc |= check<4096>(9);
  if (full) rc |= check<4096>(28, true);
  if (full) rc |= check<8192>(10);
  if (full) rc |= check<8192>(29, true);
  if (full) rc |= check<16384>(11);
  return rc;
}
