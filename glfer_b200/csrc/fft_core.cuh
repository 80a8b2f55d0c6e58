// fft_core.cuh -- register/shared-memory FFT building blocks for the gram kernels.
//
// Real-input FFT of length N computed as a complex FFT of length M = N/2 over
// z[m] = x[2m] + i x[2m+1], followed by the even/odd split.  One frame is owned by
// T = M/16 threads; every thread holds P = 16 complex points in registers and the
// passes exchange data through one M-entry float2 buffer in shared memory
// (Stockham auto-sort indexing, XOR-swizzled so both the strided first-pass stores
// and the unit-stride loads are bank-conflict free).
//
// Everything here is __host__ __device__ so tests/emu (g++) can run the exact index
// arithmetic on the CPU, phase by phase, without a GPU.  This replaces (does not
// port) the reference's serial radix-2 real FFT, fft_radix2.c:75-177; the only
// contract kept is the DFT itself and the PSD layout of fft_psd, fft.c:203-217.
#pragma once

#if defined(__CUDACC__)
#define GLB_HD __host__ __device__ __forceinline__
#else
#define GLB_HD inline
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace glb {

constexpr int kPoints = 16;          // complex points per thread

// ------------------------------------------------------------------ radix plans
// The last pass is always radix 8 with two butterflies (j, M/8 - j) per thread, so
// that Z[k] and Z[M-k] meet in one thread and the real-FFT split needs no exchange.
// The first pass is radix 16 whenever M allows: every later store then lands in
// aligned runs of 16 consecutive entries (conflict free under the swizzle).
template <int M> struct Plan;
#define GLB_PLAN(M_, NP_, A, B, C, D) \
  template <> struct Plan<M_> { static constexpr int NP = NP_; static constexpr int R0 = A, R1 = B, R2 = C, R3 = D; };
GLB_PLAN(16, 2, 2, 8, 1, 1)
GLB_PLAN(32, 2, 4, 8, 1, 1)
GLB_PLAN(64, 2, 8, 8, 1, 1)
GLB_PLAN(128, 2, 16, 8, 1, 1)
GLB_PLAN(256, 3, 16, 2, 8, 1)
GLB_PLAN(512, 3, 16, 4, 8, 1)
GLB_PLAN(1024, 3, 16, 8, 8, 1)
GLB_PLAN(2048, 3, 16, 16, 8, 1)
GLB_PLAN(4096, 4, 16, 8, 4, 8)
GLB_PLAN(8192, 4, 16, 16, 4, 8)
GLB_PLAN(16384, 4, 16, 16, 8, 8)
#undef GLB_PLAN

template <int M, int P> struct PlanRadix;
template <int M> struct PlanRadix<M, 0> { static constexpr int R = Plan<M>::R0; static constexpr int Ns = 1; };
template <int M> struct PlanRadix<M, 1> { static constexpr int R = Plan<M>::R1; static constexpr int Ns = Plan<M>::R0; };
template <int M> struct PlanRadix<M, 2> { static constexpr int R = Plan<M>::R2; static constexpr int Ns = Plan<M>::R0 * Plan<M>::R1; };
template <int M> struct PlanRadix<M, 3> { static constexpr int R = Plan<M>::R3; static constexpr int Ns = Plan<M>::R0 * Plan<M>::R1 * Plan<M>::R2; };

// Offset (in float2 entries) of pass p's twiddle block inside the per-plan table.
// Block of a pass with (R, Ns): entries [(r-1) * Ns + k] = exp(-2 pi i k r / (Ns R)),
// r = 1..R-1, k = 0..Ns-1.  Pass 0 has none.
template <int M, int P> struct TwOffset {
  static constexpr int value = TwOffset<M, P - 1>::value + (P - 1 == 0 ? 0 : (PlanRadix<M, P - 1>::R - 1) * PlanRadix<M, P - 1>::Ns);
};
template <int M> struct TwOffset<M, 0> { static constexpr int value = 0; };
template <int M> struct TwTotal { static constexpr int value = TwOffset<M, Plan<M>::NP>::value; };

// ------------------------------------------------------------------ complex helpers
GLB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GLB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GLB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GLB_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
GLB_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
GLB_HD float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }   // a * (+i)

// XOR swizzle of a float2 index: low 4 bits (the 8-byte bank pair) ^= next 4 bits.
GLB_HD int swz(int p) { return p ^ ((p >> 4) & 15); }

// ------------------------------------------------------------------ in-register DFTs
// Forward transforms (kernel e^{-2 pi i nk/R}) over v[0], v[S], ..., v[(R-1)S];
// results in natural order at the same positions.  All indices are compile-time.
template <int S> GLB_HD void dft2(float2 *v) {
  float2 a = v[0], b = v[S];
  v[0] = cadd(a, b);
  v[S] = csub(a, b);
}

template <int S> GLB_HD void dft4(float2 *v) {
  float2 a = v[0], b = v[S], c = v[2 * S], d = v[3 * S];
  float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), bmd = csub(b, d);
  v[0] = cadd(apc, bpd);
  v[S] = cadd(amc, mul_mi(bmd));
  v[2 * S] = csub(apc, bpd);
  v[3 * S] = cadd(amc, mul_pi(bmd));
}

template <int S> GLB_HD void dft8(float2 *v) {
  const float h = 0.70710678118654752440f;
  float2 e[4] = {v[0], v[2 * S], v[4 * S], v[6 * S]};
  float2 o[4] = {v[S], v[3 * S], v[5 * S], v[7 * S]};
  dft4<1>(e);
  dft4<1>(o);
  float2 o1 = make_float2((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);      // * W8^1
  float2 o2 = mul_mi(o[2]);                                                    // * W8^2
  float2 o3 = make_float2((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);     // * W8^3
  v[0] = cadd(e[0], o[0]);
  v[4 * S] = csub(e[0], o[0]);
  v[S] = cadd(e[1], o1);
  v[5 * S] = csub(e[1], o1);
  v[2 * S] = cadd(e[2], o2);
  v[6 * S] = csub(e[2], o2);
  v[3 * S] = cadd(e[3], o3);
  v[7 * S] = csub(e[3], o3);
}

template <int S> GLB_HD void dft16(float2 *v) {
  // n = 4 n1 + n2, k = k1 + 4 k2
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos, sin(pi/8)
  const float h = 0.70710678118654752440f;
  float2 y[4][4];   // y[n2][k1]
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) {
    float2 t[4] = {v[(0 + n2) * S], v[(4 + n2) * S], v[(8 + n2) * S], v[(12 + n2) * S]};
    dft4<1>(t);
    y[n2][0] = t[0]; y[n2][1] = t[1]; y[n2][2] = t[2]; y[n2][3] = t[3];
  }
  // twiddle y[n2][k1] *= W16^(n2 k1), W16^e = exp(-2 pi i e / 16)
  y[1][1] = cmul(y[1][1], make_float2(c1, -s1));                                  // e = 1
  y[1][2] = make_float2((y[1][2].x + y[1][2].y) * h, (y[1][2].y - y[1][2].x) * h);  // e = 2
  y[1][3] = cmul(y[1][3], make_float2(s1, -c1));                                  // e = 3
  y[2][1] = make_float2((y[2][1].x + y[2][1].y) * h, (y[2][1].y - y[2][1].x) * h);  // e = 2
  y[2][2] = mul_mi(y[2][2]);                                                      // e = 4
  y[2][3] = make_float2((y[2][3].y - y[2][3].x) * h, -(y[2][3].x + y[2][3].y) * h); // e = 6
  y[3][1] = cmul(y[3][1], make_float2(s1, -c1));                                  // e = 3
  y[3][2] = make_float2((y[3][2].y - y[3][2].x) * h, -(y[3][2].x + y[3][2].y) * h); // e = 6
  y[3][3] = cmul(y[3][3], make_float2(-c1, s1));                                  // e = 9
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    float2 t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
    dft4<1>(t);
    v[(k1 + 0) * S] = t[0];
    v[(k1 + 4) * S] = t[1];
    v[(k1 + 8) * S] = t[2];
    v[(k1 + 12) * S] = t[3];
  }
}

template <int R, int S> struct Dft;
template <int S> struct Dft<2, S> { static GLB_HD void run(float2 *v) { dft2<S>(v); } };
template <int S> struct Dft<4, S> { static GLB_HD void run(float2 *v) { dft4<S>(v); } };
template <int S> struct Dft<8, S> { static GLB_HD void run(float2 *v) { dft8<S>(v); } };
template <int S> struct Dft<16, S> { static GLB_HD void run(float2 *v) { dft16<S>(v); } };

// ------------------------------------------------------------------ passes
// Element q of thread t in every non-final pass is entry t + T*q of the buffer.
//
// pass_store<M,P>: twiddle (P > 0), radix-R butterflies on the 16 register points
// and Stockham store: butterfly j = t + u*T (u < 16/R), k = j mod Ns, output r' goes
// to entry (j - k) * R + k + r' * Ns.
template <int M, int P>
GLB_HD void pass_store(float2 *v, int t, float2 *buf, const float2 *tw) {
  constexpr int T = M / kPoints;
  constexpr int R = PlanRadix<M, P>::R;
  constexpr int Ns = PlanRadix<M, P>::Ns;
  constexpr int S = kPoints / R;         // butterflies per thread, and register stride
#pragma unroll
  for (int u = 0; u < S; u++) {
    const int j = t + u * T;
    const int k = j & (Ns - 1);
    if (P > 0) {
      const float2 *twp = tw + TwOffset<M, P>::value + k;
#pragma unroll
      for (int r = 1; r < R; r++) v[u + r * S] = cmul(v[u + r * S], twp[(r - 1) * Ns]);
    }
    Dft<R, S>::run(v + u);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) buf[swz(base + r * Ns)] = v[u + r * S];
  }
}

template <int M>
GLB_HD void pass_load(float2 *v, int t, const float2 *buf) {
  constexpr int T = M / kPoints;
#pragma unroll
  for (int q = 0; q < kPoints; q++) v[q] = buf[swz(t + T * q)];
}

// Final pass: radix 8, Ns = M/8 = 2T.  Thread t owns butterflies jA = t and
// jB = 2T - t (thread 0: jA = 0, jB = T).  On return v[r'] = Z[jA + r' 2T] and
// v[8 + r'] = Z[jB + r' 2T].
template <int M>
GLB_HD void last_pass(float2 *v, int t, const float2 *buf, const float2 *tw) {
  constexpr int T = M / kPoints;
  constexpr int NP = Plan<M>::NP;
  constexpr int Ns = 2 * T;
  const int jA = t;
  const int jB = (t == 0) ? T : 2 * T - t;
  const float2 *twA = tw + TwOffset<M, NP - 1>::value + jA;
  const float2 *twB = tw + TwOffset<M, NP - 1>::value + jB;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    v[r] = buf[swz(jA + r * Ns)];
    v[8 + r] = buf[swz(jB + r * Ns)];
  }
#pragma unroll
  for (int r = 1; r < 8; r++) {
    v[r] = cmul(v[r], twA[(r - 1) * Ns]);
    v[8 + r] = cmul(v[8 + r], twB[(r - 1) * Ns]);
  }
  dft8<1>(v);
  dft8<1>(v + 8);
}

// Real-FFT split of one conjugate pair: given Zk = Z[k], Zm = Z[M-k] and
// vk = -i exp(-2 pi i k / N), returns 2 X[k] in a and 2 X[M-k]^* in b
// (scaled by whatever scale the input carried).
GLB_HD void split_pair(float2 zk, float2 zm, float2 vk, float2 &a, float2 &b) {
  float2 p = make_float2(zk.x + zm.x, zk.y - zm.y);
  float2 q = make_float2(zk.x - zm.x, zk.y + zm.y);
  float2 vq = cmul(vk, q);
  a = cadd(p, vq);
  b = csub(p, vq);
}

GLB_HD float norm2(float2 a) { return a.x * a.x + a.y * a.y; }

// Bin bookkeeping of the final pass.  For thread t >= 1 pair r' (0..7) is
// (Z[k], Z[M-k]) with k = t + r' 2T held in (v[r'], v[8 + 7 - r']).  Thread 0 pairs
// inside each butterfly: A: (v[r'], v[(8 - r') & 7]) k = r' 2T, r' = 0..4 (k = 0 and
// k = M/2 are self-paired); B: (v[8 + r'], v[8 + 7 - r']) k = T + r' 2T, r' = 0..3.
// Results are numbered by "slot": a thread owns slots 0..15 (thread 0: 0..16) and
// slot_bin() gives the PSD bin of a slot, so accumulators can live in registers
// across tapers.  Every bin 0..M belongs to exactly one (thread, slot).
template <int M> GLB_HD int slot_bin(int t, int slot) {
  constexpr int T = M / kPoints;
  if (t != 0) {
    const int k = t + (slot >> 1) * 2 * T;
    return (slot & 1) ? M - k : k;
  }
  if (slot < 8) {
    const int k = (slot >> 1) * 2 * T;
    return (slot & 1) ? M - k : k;
  }
  if (slot == 8) return M / 2;
  const int k = T + ((slot - 9) >> 1) * 2 * T;
  return ((slot - 9) & 1) ? M - k : k;
}
template <int M> GLB_HD int slot_count(int t) { return t != 0 ? 16 : 17; }

// f(slot, value, conjugated): value = 2 X[slot_bin(t, slot)] (conjugated when the flag
// is set), still carrying the scale folded into the taper.
template <int M, class F>
GLB_HD void emit_bins(const float2 *v, int t, const float2 *vtab, F &&f) {
  constexpr int T = M / kPoints;
  float2 a, b;
  if (t != 0) {
#pragma unroll
    for (int rp = 0; rp < 8; rp++) {
      const int k = t + rp * 2 * T;
      split_pair(v[rp], v[8 + 7 - rp], vtab[k], a, b);
      f(2 * rp, a, false);
      f(2 * rp + 1, b, true);
    }
  } else {
    split_pair(v[0], v[0], vtab[0], a, b);
    f(0, a, false);
    f(1, b, true);
#pragma unroll
    for (int rp = 1; rp < 4; rp++) {
      split_pair(v[rp], v[8 - rp], vtab[rp * 2 * T], a, b);
      f(2 * rp, a, false);
      f(2 * rp + 1, b, true);
    }
    split_pair(v[4], v[4], vtab[M / 2], a, b);
    f(8, a, false);
#pragma unroll
    for (int rp = 0; rp < 4; rp++) {
      split_pair(v[8 + rp], v[8 + 7 - rp], vtab[T + rp * 2 * T], a, b);
      f(9 + 2 * rp, a, false);
      f(10 + 2 * rp, b, true);
    }
  }
}

}  // namespace glb
