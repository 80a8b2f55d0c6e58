// fft_wpf.cuh -- "warp per frame" FFT core: one frame is owned by T = M/64 lanes of ONE warp,
// each lane holding 64 complex points in registers.  Two passes only:
//   pass A  radix RA = M/32 (64/RA butterflies per lane), no twiddles, results scattered
//           to a padded shared-memory tile private to the lane group;
//   pass B  radix 32, two butterflies (j, 2T - j) per lane so that Z[k] and Z[M-k] meet in
//           one lane and the real-FFT split needs no further exchange.
// Because a frame never leaves its warp there is no block barrier anywhere (only
// __syncwarp around the single exchange), half the shared-memory traffic of the three-pass
// CTA-wide core (fft_core.cuh) and ~40 % fewer instructions per frame.  Valid for
// 256 <= M <= 2048 (N = 512 .. 4096).
//
// Like fft_core.cuh everything is __host__ __device__ so tests/emu can run the index and
// twiddle arithmetic on the CPU.
#pragma once
#include "fft_core.cuh"
#include "twiddle_consts.cuh"

namespace glb {

constexpr int kWP = 64;   // complex points per lane

// composite in-register DFT of size R1*R2 over v[0], v[S], ...: n = R2 n1 + n2, k = k1 + R1 k2
template <int R1, int R2, int S> GLB_HD void dft_comp(float2 *v) {
  constexpr int R = R1 * R2;
#pragma unroll
  for (int n2 = 0; n2 < R2; n2++) Dft<R1, R2 * S>::run(v + n2 * S);       // over n1, stride R2
#pragma unroll
  for (int k1 = 1; k1 < R1; k1++)
#pragma unroll
    for (int n2 = 1; n2 < R2; n2++) v[(R2 * k1 + n2) * S] = mul_wconst<R>(v[(R2 * k1 + n2) * S], n2 * k1);
#pragma unroll
  for (int k1 = 0; k1 < R1; k1++) Dft<R2, S>::run(v + R2 * k1 * S);      // over n2, stride 1
  // position R2 k1 + k2 holds X[k1 + R1 k2]: transpose back to natural order
  float2 tmp[R];
#pragma unroll
  for (int i = 0; i < R; i++) tmp[i] = v[i * S];
#pragma unroll
  for (int k1 = 0; k1 < R1; k1++)
#pragma unroll
    for (int k2 = 0; k2 < R2; k2++) v[(k1 + R1 * k2) * S] = tmp[R2 * k1 + k2];
}
template <int S> struct Dft<32, S> { static GLB_HD void run(float2 *v) { dft_comp<4, 8, S>(v); } };
template <int S> struct Dft<64, S> { static GLB_HD void run(float2 *v) { dft_comp<8, 8, S>(v); } };

template <int M> struct Wpf {
  static constexpr int T = M / kWP;            // lanes per frame
  static constexpr int RA = M / 32;            // radix of pass A
  static constexpr int SA = kWP / RA;          // pass-A butterflies per lane
  static constexpr int ROW = RA + 2;           // padded row (float2 entries): conflict-free 128-bit stores
  static constexpr int TILE = 32 * ROW;        // float2 entries per frame
  static_assert(M >= 256 && M <= 2048, "warp-per-frame core covers N = 512..4096");
};

// frame-invariant per-lane constants kept in registers
struct WpfRegs {
  float2 wb[10];      // w^1 w^2 w^3 w^4 w^8 ... w^28, w = W_M^(tb); tb = t (lane 0 of a frame: T)
  float2 v_lo, v_hi;  // split factor bases for slots 0..15 / 16..31
  int kb_lo, kb_hi;   // bin of slot s is kb + 2T s
};

// roots[k] = exp(-2 pi i k / M) (k < M) and vtab[k] = -i exp(-2 pi i k / N) (k < M) are the
// host-built (double precision, rounded once) tables of tables.hpp
template <int M>
GLB_HD void wpf_load_regs(WpfRegs &r, int t, const float2 *roots, const float2 *vtab) {
  constexpr int T = Wpf<M>::T;
  const int tb = (t == 0) ? T : t;
  const int e[10] = {1, 2, 3, 4, 8, 12, 16, 20, 24, 28};
#pragma unroll
  for (int i = 0; i < 10; i++) r.wb[i] = roots[(tb * e[i]) & (M - 1)];
  if (t != 0) {
    r.v_lo = vtab[t];
    r.v_hi = r.v_lo;
    r.kb_lo = t;
    r.kb_hi = t;
  } else {
    // lane 0 owns the two self-paired butterflies 0 and T: slots 0..15 pair inside A
    // (k = 2T s), slots 16..31 pair inside B (k = T + 2T (s - 16)); W_64^(s-16) = i W_64^s
    r.v_lo = vtab[0];
    r.v_hi = mul_pi(vtab[T]);
    r.kb_lo = 0;
    r.kb_hi = T - 32 * T;
  }
}

// pass A: SA butterflies of radix RA over v[u + r SA] (inputs z[t + T q], q = u + r SA)
template <int M> GLB_HD void wpf_pass_a(float2 *v) {
#pragma unroll
  for (int u = 0; u < Wpf<M>::SA; u++) Dft<Wpf<M>::RA, Wpf<M>::SA>::run(v + u);
}

// butterfly j = t + u T writes its RA outputs to row j of the padded tile (128-bit stores)
template <int M> GLB_HD void wpf_scatter(const float2 *v, int t, float2 *tile) {
  constexpr int T = Wpf<M>::T, RA = Wpf<M>::RA, SA = Wpf<M>::SA, ROW = Wpf<M>::ROW;
#pragma unroll
  for (int u = 0; u < SA; u++) {
    float4 *dst = reinterpret_cast<float4 *>(tile + (t + u * T) * ROW);
#pragma unroll
    for (int r = 0; r < RA; r += 2)
      dst[r / 2] = make_float4(v[u + r * SA].x, v[u + r * SA].y, v[u + (r + 1) * SA].x, v[u + (r + 1) * SA].y);
  }
}

// pass B inputs: entry j + 2T r = row r, column j (2T = RA): v[r] for jA = t, v[32 + r] for jB = 2T - tb
template <int M> GLB_HD void wpf_gather(float2 *v, int t, const float2 *tile) {
  constexpr int T = Wpf<M>::T, ROW = Wpf<M>::ROW;
  const int jA = t, jB = (t == 0) ? T : 2 * T - t;
#pragma unroll
  for (int r = 0; r < 32; r++) {
    v[r] = tile[jA + r * ROW];
    v[32 + r] = tile[jB + r * ROW];
  }
}

// pass B: A gets W_M^(t r) (identity for lane 0), B gets conj(w^r) and a one-place cyclic shift of
// the DFT outputs (W_M^(jB r) = W_32^r conj(W_M^(tb r))).  On return v[r'] = Z[jA + 2T r'],
// v[32 + r'] = Z[jB + 2T r'].
template <int M> GLB_HD void wpf_pass_b(float2 *v, int t, const WpfRegs &rg) {
  float2 w[32];
  w[1] = rg.wb[0];
  w[2] = rg.wb[1];
  w[3] = rg.wb[2];
#pragma unroll
  for (int a = 1; a < 8; a++) {
    w[4 * a] = rg.wb[2 + a];
    w[4 * a + 1] = cmul(w[4 * a], w[1]);
    w[4 * a + 2] = cmul(w[4 * a], w[2]);
    w[4 * a + 3] = cmul(w[4 * a], w[3]);
  }
  const bool lane0 = (t == 0);
#pragma unroll
  for (int r = 1; r < 32; r++) {
    const float2 a = cmul(v[r], w[r]);
    v[r] = lane0 ? v[r] : a;
    v[32 + r] = cmulc(v[32 + r], w[r]);
  }
  Dft<32, 1>::run(v);
  Dft<32, 1>::run(v + 32);
  const float2 first = v[32];
#pragma unroll
  for (int r = 0; r < 31; r++) v[32 + r] = v[33 + r];
  v[63] = first;
}

// slot s (0..31) of a lane: the conjugate pair it evaluates and the bins it produces.
//   general lane: (A[s], B[31 - s]), k = t + 2T s
//   lane 0      : s < 16: (A[s], A[(32 - s) & 31]), k = 2T s;  s >= 16: (B[s-16], B[47-s]), k = T + 2T (s-16)
// f(slot, k, value, conj) is called for bin k (value = 2 X[k]) and bin M - k (value = 2 X[M-k]^*).
// Lane 0 additionally evaluates the self-paired A[16] (bin M/2).
template <int M, class F> GLB_HD void wpf_emit(const float2 *v, int t, const WpfRegs &rg, F &&f) {
  constexpr int T = Wpf<M>::T;
  const bool lane0 = (t == 0);
#pragma unroll
  for (int s = 0; s < 32; s++) {
    float2 zk, zm;
    if (s < 16) {
      zk = v[s];
      const float2 m0 = v[(32 - s) & 31];
      zm = lane0 ? m0 : v[32 + 31 - s];
    } else {
      zk = lane0 ? v[32 + s - 16] : v[s];
      zm = lane0 ? v[32 + 47 - s] : v[32 + 31 - s];
    }
    const float2 vb = (s < 16) ? rg.v_lo : rg.v_hi;
    const int kb = (s < 16) ? rg.kb_lo : rg.kb_hi;
    const float2 p = make_float2(zk.x + zm.x, zk.y - zm.y);
    const float2 q = make_float2(zk.x - zm.x, zk.y + zm.y);
    const float2 vq = cmul(vb, mul_wconst<64>(q, s));
    f(s, kb + 2 * T * s, cadd(p, vq), false);
    f(s, M - (kb + 2 * T * s), csub(p, vq), true);
  }
  if (lane0) {
    // A[16] pairs with itself: k = M/2, V = -i W_N^(M/2) = -1
    const float2 z = v[16];
    f(32, M / 2, make_float2(2.f * z.x, -2.f * z.y), false);
  }
}

}  // namespace glb
