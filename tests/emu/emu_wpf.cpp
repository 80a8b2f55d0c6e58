// CPU emulation of the warp-per-frame FFT core (fft_wpf.cuh): every lane of one frame is run
// phase by phase with the padded tile as a plain array; checks X against a long-double DFT
// and that every bin 0..M is produced exactly once.  Prints "M maxrel_psd maxabs_spec bad".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../glfer_b200/csrc/tables.hpp"
#include "../../glfer_b200/csrc/fft_wpf.cuh"

using namespace glb;

template <int M> int check(unsigned seed) {
  constexpr int N = 2 * M, T = Wpf<M>::T;
  std::vector<float> x(N);
  srand(seed);
  for (auto &v : x) v = (float) rand() / RAND_MAX - 0.5f;
  auto roots = build_roots(M);
  auto vtab = build_vtab(M);
  std::vector<std::vector<float2>> regs(T, std::vector<float2>(kWP));
  std::vector<WpfRegs> rg(T);
  std::vector<float4> tile4(Wpf<M>::TILE / 2 + 2);
  float2 *tile = reinterpret_cast<float2 *>(tile4.data());
  for (int t = 0; t < T; t++) {
    wpf_load_regs<M>(rg[t], t, roots.data(), vtab.data());
    for (int q = 0; q < kWP; q++) regs[t][q] = make_float2(x[2 * (t + T * q)], x[2 * (t + T * q) + 1]);
  }
  for (int t = 0; t < T; t++) wpf_pass_a<M>(regs[t].data());
  for (int t = 0; t < T; t++) wpf_scatter<M>(regs[t].data(), t, tile);
  for (int t = 0; t < T; t++) wpf_gather<M>(regs[t].data(), t, tile);
  for (int t = 0; t < T; t++) wpf_pass_b<M>(regs[t].data(), t, rg[t]);
  std::vector<double> psd(M + 1, -1.0);
  std::vector<float2> spec(M + 1);
  std::vector<int> hits(M + 1, 0);
  int bad = 0;
  for (int t = 0; t < T; t++)
    wpf_emit<M>(regs[t].data(), t, rg[t], [&](int slot, int bin, float2 a, bool conj) {
      if (bin < 0 || bin > M || slot > 32) { bad++; return; }
      hits[bin]++;
      psd[bin] = 0.25 * (double) norm2(a);
      spec[bin] = make_float2(0.5f * a.x, conj ? -0.5f * a.y : 0.5f * a.y);
    });
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
  for (int k = 0; k <= M; k++) {
    if (hits[k] != 1) bad++;
    double ref = (double) (xr[k] * xr[k] + xi[k] * xi[k]);
    double rel = std::fabs(psd[k] - ref) / (ref + 1e-6 * ref_rms * ref_rms);
    if (rel > maxrel) maxrel = rel;
    double ea = std::hypot((double) spec[k].x - (double) xr[k], (double) spec[k].y - (double) xi[k]) / ref_rms;
    if (ea > maxabs) maxabs = ea;
  }
  printf("%d %.3e %.3e %d wpf\n", M, maxrel, maxabs, bad);
  return bad != 0 || maxrel > 1e-4 || maxabs > 1e-5;
}

int main() {
  int rc = 0;
  rc |= check<256>(31);
  rc |= check<512>(32);
  rc |= check<1024>(33);
  rc |= check<2048>(34);
  return rc;
}
