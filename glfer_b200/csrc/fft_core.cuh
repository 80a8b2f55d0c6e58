// fft_core.cuh -- register/shared-memory FFT building blocks for the gram kernels.
//
// Real-input FFT of length N computed as a complex FFT of length M = N/2 over
// z[m] = x[2m] + i x[2m+1], followed by the even/odd split.  One frame is owned by
// T = M/16 threads; every thread holds P = 16 complex points in registers and the
// passes exchange data through one M-entry float2 buffer in shared memory
// (Stockham auto-sort indexing; the buffer is padded by 2 entries per 16 so that the
// strided first-pass stores and the unit-stride loads are both bank-conflict free AND
// every address is a per-thread base plus a compile-time immediate).
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
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
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
// Complex add / subtract.  On sm_100a (device code, GLB_PACKED_ADD) they are single packed
// instructions (FADD2 on an aligned register pair): half the issue slots of the scalar form.
#if !defined(GLB_PACKED_ADD)
#define GLB_PACKED_ADD 1
#endif
#if defined(__CUDA_ARCH__) && GLB_PACKED_ADD
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long *>(&r))
      : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  float2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long *>(&r))
      : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return r;
}
// component-wise product (window / taper multiply): one FMUL2
__device__ __forceinline__ float2 emul(float2 a, float2 b) {
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long *>(&r))
      : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
  return r;
}
#else
GLB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GLB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GLB_HD float2 emul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif
GLB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
GLB_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
GLB_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
GLB_HD float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }   // a * (+i)

// Padded float2 index: 2 spare entries after every 16 (16 B alignment of even entries is
// kept, so pairs can move as one 128-bit access).  pad(a + c) = pad(a) + c + c/8 whenever
// c is a multiple of 16: strided accesses become base + immediate.
GLB_HD int pad(int p) { return p + 2 * (p >> 4); }
template <int M> struct BufSize { static constexpr int value = M + M / 8 + 2; };   // float2 entries

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
template <int M> GLB_HD int ld_index(int t, int q) {
  constexpr int T = M / kPoints;
  if constexpr (T % 16 == 0) return pad(t) + q * (T + T / 8);
  else return pad(t + T * q);
}

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
    if constexpr (P == 0 && R == 16) {
      // 16 consecutive entries 16j..16j+15 -> padded 18j..18j+15, as eight 128-bit stores
      float4 *dst = reinterpret_cast<float4 *>(buf + 18 * j);
#pragma unroll
      for (int r = 0; r < 16; r += 2) dst[r / 2] = make_float4(v[u + r * S].x, v[u + r * S].y, v[u + (r + 1) * S].x, v[u + (r + 1) * S].y);
    } else if constexpr (Ns % 16 == 0) {
      const int base = pad((j - k) * R + k);
#pragma unroll
      for (int r = 0; r < R; r++) buf[base + r * (Ns + Ns / 8)] = v[u + r * S];
    } else {
      const int base = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; r++) buf[pad(base + r * Ns)] = v[u + r * S];
    }
  }
}

template <int M>
GLB_HD void pass_load(float2 *v, int t, const float2 *buf) {
#pragma unroll
  for (int q = 0; q < kPoints; q++) v[q] = buf[ld_index<M>(t, q)];
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
  if constexpr (Ns % 16 == 0) {
    const int bA = pad(jA), bB = pad(jB);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = buf[bA + r * (Ns + Ns / 8)];
      v[8 + r] = buf[bB + r * (Ns + Ns / 8)];
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = buf[pad(jA + r * Ns)];
      v[8 + r] = buf[pad(jB + r * Ns)];
    }
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

// ------------------------------------------------------------------ register-resident twiddles
// The twiddles a thread needs do not depend on the frame, so a kernel that walks many
// frames keeps a few base powers in registers and rebuilds the rest with one complex
// multiply each (products of two correctly rounded roots: ~1.2e-7 relative, the same
// order as the FFT's own rounding).  This removes every per-frame table load:
//   mid pass, radix 16: w^1 w^2 w^3 w^4 w^8 w^12 kept, w^(4a+b) = w^(4a) w^b
//   mid pass, radix 8 : w^1..w^4 kept, w^5 w^6 w^7 = w^4 w^(1..3)
//   last pass (radix 8): butterfly A as radix 8; butterfly B = 2T - t uses
//     W_M^(jB r) = W_8^r conj(W_M^(t r)): conjugate twiddles and a one-place cyclic shift
//     of the DFT outputs (thread 0, whose B is butterfly T, multiplies by W_16^r instead)
//   split factors: V_(t + r' 2T) = V_t W_16^r'
template <int R> struct TwBases;
template <> struct TwBases<16> { static constexpr int n = 6; };
template <> struct TwBases<8> { static constexpr int n = 4; };
template <> struct TwBases<4> { static constexpr int n = 3; };
template <> struct TwBases<2> { static constexpr int n = 1; };
template <> struct TwBases<1> { static constexpr int n = 0; };

struct TwRegs {
  float2 mid[2][12];   // passes 1 and 2 (when they are not the last): [u * nbases + i]
  float2 last[4];      // w_a^1..w_a^4, w_a = W_M^t
  float2 v0;           // V_t = -i exp(-2 pi i t / N)
};

template <int M, int P>
GLB_HD void load_mid_bases(TwRegs &tr, int t, const float2 *tw) {
  if constexpr (P >= 1 && P < Plan<M>::NP - 1) {
    constexpr int T = M / kPoints;
    constexpr int R = PlanRadix<M, P>::R, Ns = PlanRadix<M, P>::Ns, S = kPoints / R, NB = TwBases<R>::n;
#pragma unroll
    for (int u = 0; u < S; u++) {
      const int k = (t + u * T) & (Ns - 1);
      const float2 *twp = tw + TwOffset<M, P>::value + k;
      if constexpr (R == 16) {
        const int e[6] = {1, 2, 3, 4, 8, 12};
#pragma unroll
        for (int i = 0; i < 6; i++) tr.mid[P - 1][u * NB + i] = twp[(e[i] - 1) * Ns];
      } else {
#pragma unroll
        for (int i = 0; i < NB; i++) tr.mid[P - 1][u * NB + i] = twp[i * Ns];
      }
    }
  }
}

template <int M>
GLB_HD void load_tw_regs(TwRegs &tr, int t, const float2 *tw, const float2 *vtab) {
  constexpr int T = M / kPoints, NP = Plan<M>::NP;
  load_mid_bases<M, 1>(tr, t, tw);
  load_mid_bases<M, 2>(tr, t, tw);
  const float2 *twA = tw + TwOffset<M, NP - 1>::value + t;
#pragma unroll
  for (int i = 0; i < 4; i++) tr.last[i] = twA[i * 2 * T];
  tr.v0 = vtab[t];
}

// v[u + r S] *= w^r for r = 1..R-1 from the kept bases
template <int R, int S>
GLB_HD void apply_tw_bases(float2 *v, const float2 *b) {
  if constexpr (R == 16) {
    const float2 w1 = b[0], w2 = b[1], w3 = b[2], w4 = b[3], w8 = b[4], w12 = b[5];
    v[1 * S] = cmul(v[1 * S], w1);
    v[2 * S] = cmul(v[2 * S], w2);
    v[3 * S] = cmul(v[3 * S], w3);
    v[4 * S] = cmul(v[4 * S], w4);
    v[5 * S] = cmul(v[5 * S], cmul(w4, w1));
    v[6 * S] = cmul(v[6 * S], cmul(w4, w2));
    v[7 * S] = cmul(v[7 * S], cmul(w4, w3));
    v[8 * S] = cmul(v[8 * S], w8);
    v[9 * S] = cmul(v[9 * S], cmul(w8, w1));
    v[10 * S] = cmul(v[10 * S], cmul(w8, w2));
    v[11 * S] = cmul(v[11 * S], cmul(w8, w3));
    v[12 * S] = cmul(v[12 * S], w12);
    v[13 * S] = cmul(v[13 * S], cmul(w12, w1));
    v[14 * S] = cmul(v[14 * S], cmul(w12, w2));
    v[15 * S] = cmul(v[15 * S], cmul(w12, w3));
  } else if constexpr (R == 8) {
    const float2 w1 = b[0], w2 = b[1], w3 = b[2], w4 = b[3];
    v[1 * S] = cmul(v[1 * S], w1);
    v[2 * S] = cmul(v[2 * S], w2);
    v[3 * S] = cmul(v[3 * S], w3);
    v[4 * S] = cmul(v[4 * S], w4);
    v[5 * S] = cmul(v[5 * S], cmul(w4, w1));
    v[6 * S] = cmul(v[6 * S], cmul(w4, w2));
    v[7 * S] = cmul(v[7 * S], cmul(w4, w3));
  } else if constexpr (R == 4) {
    v[1 * S] = cmul(v[1 * S], b[0]);
    v[2 * S] = cmul(v[2 * S], b[1]);
    v[3 * S] = cmul(v[3 * S], b[2]);
  } else if constexpr (R == 2) {
    v[1 * S] = cmul(v[1 * S], b[0]);
  }
}

// mid pass with register twiddles, split in two so the block barrier can sit between the
// arithmetic and the stores (warps then wait with their butterflies already done)
template <int M, int P>
GLB_HD void pass_compute_rt(float2 *v, const TwRegs &tr) {
  constexpr int R = PlanRadix<M, P>::R, S = kPoints / R, NB = TwBases<R>::n;
#pragma unroll
  for (int u = 0; u < S; u++) {
    if constexpr (P > 0) apply_tw_bases<R, S>(v + u, &tr.mid[P - 1][u * NB]);
    Dft<R, S>::run(v + u);
  }
}

template <int M, int P>
GLB_HD void pass_scatter(const float2 *v, int t, float2 *buf) {
  constexpr int T = M / kPoints;
  constexpr int R = PlanRadix<M, P>::R, Ns = PlanRadix<M, P>::Ns, S = kPoints / R;
#pragma unroll
  for (int u = 0; u < S; u++) {
    const int j = t + u * T;
    const int k = j & (Ns - 1);
    if constexpr (P == 0 && R == 16) {
      float4 *dst = reinterpret_cast<float4 *>(buf + 18 * j);
#pragma unroll
      for (int r = 0; r < 16; r += 2) dst[r / 2] = make_float4(v[u + r * S].x, v[u + r * S].y, v[u + (r + 1) * S].x, v[u + (r + 1) * S].y);
    } else if constexpr (Ns % 16 == 0) {
      const int base = pad((j - k) * R + k);
#pragma unroll
      for (int r = 0; r < R; r++) buf[base + r * (Ns + Ns / 8)] = v[u + r * S];
    } else {
      const int base = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; r++) buf[pad(base + r * Ns)] = v[u + r * S];
    }
  }
}

// last pass with register twiddles; on return v[r'] = Z[jA + r' 2T], v[8 + r'] = Z[jB + r' 2T]
template <int M>
GLB_HD void last_pass_rt(float2 *v, int t, const float2 *buf, const float2 *tw, const TwRegs &tr) {
  constexpr int T = M / kPoints, NP = Plan<M>::NP, Ns = 2 * T;
  const int jA = t;
  const int jB = (t == 0) ? T : 2 * T - t;
  if constexpr (Ns % 16 == 0) {
    const int bA = pad(jA), bB = pad(jB);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = buf[bA + r * (Ns + Ns / 8)];
      v[8 + r] = buf[bB + r * (Ns + Ns / 8)];
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = buf[pad(jA + r * Ns)];
      v[8 + r] = buf[pad(jB + r * Ns)];
    }
  }
  const float2 w1 = tr.last[0], w2 = tr.last[1], w3 = tr.last[2], w4 = tr.last[3];
  const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  const float2 w[8] = {make_float2(1.f, 0.f), w1, w2, w3, w4, w5, w6, w7};
#pragma unroll
  for (int r = 1; r < 8; r++) v[r] = cmul(v[r], w[r]);
  dft8<1>(v);
  if (t != 0) {
    // B: x_r conj(w_a^r), DFT8, outputs shifted by one place
    float2 y[8];
    y[0] = v[8];
#pragma unroll
    for (int r = 1; r < 8; r++) y[r] = cmulc(v[8 + r], w[r]);
    dft8<1>(y);
#pragma unroll
    for (int r = 0; r < 8; r++) v[8 + r] = y[(r + 1) & 7];
  } else {
    // thread 0: butterfly T, twiddles W_M^(T r) = W_16^r from the table
    const float2 *twB = tw + TwOffset<M, NP - 1>::value + T;
#pragma unroll
    for (int r = 1; r < 8; r++) v[8 + r] = cmul(v[8 + r], twB[(r - 1) * Ns]);
    dft8<1>(v + 8);
  }
}

GLB_HD float2 w16_mul(float2 q, int e) {
  // q * exp(-2 pi i e / 16), e = 0..7, constants folded after unrolling
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  switch (e) {
    case 0: return q;
    case 1: return cmul(q, make_float2(c1, -s1));
    case 2: return make_float2((q.x + q.y) * h, (q.y - q.x) * h);
    case 3: return cmul(q, make_float2(s1, -c1));
    case 4: return mul_mi(q);
    case 5: return cmul(q, make_float2(-s1, -c1));
    case 6: return make_float2((q.y - q.x) * h, -(q.x + q.y) * h);
    default: return cmul(q, make_float2(-c1, -s1));
  }
}

// split with V_k = V_t W_16^rp rebuilt from the kept V_t
GLB_HD void split_pair_rt(float2 zk, float2 zm, float2 v0, int rp, float2 &a, float2 &b) {
  float2 p = make_float2(zk.x + zm.x, zk.y - zm.y);
  float2 q = make_float2(zk.x - zm.x, zk.y + zm.y);
  float2 vq = cmul(v0, w16_mul(q, rp));
  a = cadd(p, vq);
  b = csub(p, vq);
}

template <int M, class F>
GLB_HD void emit_bins_rt(const float2 *v, int t, const float2 *vtab, const TwRegs &tr, F &&f) {
  constexpr int T = M / kPoints;
  float2 a, b;
  if (t != 0) {
#pragma unroll
    for (int rp = 0; rp < 8; rp++) {
      split_pair_rt(v[rp], v[8 + 7 - rp], tr.v0, rp, a, b);
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
