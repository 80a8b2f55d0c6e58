// fft_core.cuh -- register/shared-memory FFT building blocks for the gram kernels.
//
// Real-input FFT of length N computed as a complex FFT of length M = N/2 over
// z[m] = x[2m] + i x[2m+1], followed by the even/odd split.  One frame is owned by
// T = M/16 threads; every thread holds P = 16 complex points in registers and the
// passes exchange data through one M-entry float2 buffer in shared memory
// (Stockham auto-sort indexing; the buffer is padded by 1 entry per 16 so that the
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

// Checked build (-DGLB_CHECKED=1, glfer_b200.build.build_variant("checked", ...)): every shared-memory
// exchange index, every bulk-copy source range and every row store is bounds-checked on the device and
// traps on a violation.  compute-sanitizer is closed on the GPU pool this was developed on; the checked
// library run over every kernel family (tools/sanitize_driver.py) is the memory-safety evidence instead.
#ifndef GLB_CHECKED
#define GLB_CHECKED 0
#endif
#if GLB_CHECKED && defined(__CUDA_ARCH__)
#define GLB_CHECK(cond) do { if (!(cond)) { printf("GLB_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define GLB_CHECK(cond) do { } while (0)
#endif

namespace glb {

constexpr int kPoints = 16;          // complex points per thread
// index of an exchange-buffer access, checked against the buffer size in the checked build
template <int LIMIT> GLB_HD int chk(int idx) {
  GLB_CHECK(idx >= 0 && idx < LIMIT);
  return idx;
}

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
// A complex value is one float2 = one aligned 64-bit register pair (re, im).  On sm_100a the
// FP32 pipe has packed two-lane instructions (FADD2 / FMUL2 / FFMA2: PTX add/mul/fma.rn.f32x2)
// whose SASS operands can be the pair as it is, the pair with its halves swapped, one 32-bit
// register broadcast to both lanes, or a broadcast immediate, each optionally negated -- ptxas
// folds swp(), bc() and neg() below into those operand forms.  With them every step of a
// complex FFT is a packed instruction on the natural (re, im) pair:
//   a + b, a - b                      one FADD2
//   a -/+ i b                         one FFMA2: swp(b) * (1,-1) + a
//   a * w, w kept as (wr, (-wi, wi))  two: FMUL2 swp(a) * q, FFMA2 a * bc(wr) + t
// i.e. half the issue slots of the scalar form at the same FP32 lane throughput.
#if !defined(GLB_PACKED)
#define GLB_PACKED 1
#endif
#if defined(__CUDA_ARCH__) && GLB_PACKED
#define GLB_U64(x) (*reinterpret_cast<unsigned long long *>(&(x)))
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(GLB_U64(r)) : "l"(GLB_U64(a)), "l"(GLB_U64(b)));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(GLB_U64(r)) : "l"(GLB_U64(a)), "l"(GLB_U64(b)));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(GLB_U64(r)) : "l"(GLB_U64(a)), "l"(GLB_U64(b)));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(GLB_U64(r)) : "l"(GLB_U64(a)), "l"(GLB_U64(b)), "l"(GLB_U64(c)));
  return r;
}
#else
GLB_HD float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GLB_HD float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GLB_HD float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#if defined(__CUDA_ARCH__)
GLB_HD float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
GLB_HD float2 fma2(float2 a, float2 b, float2 c) { return make_float2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)); }
#endif
#endif
GLB_HD float2 swp(float2 a) { return make_float2(a.y, a.x); }       // operand form .LO_HI
GLB_HD float2 bc(float s) { return make_float2(s, s); }             // operand form .F32 / immediate
GLB_HD float2 neg(float2 a) { return make_float2(-a.x, -a.y); }     // operand negation
GLB_HD float2 pm() { return make_float2(1.f, -1.f); }
GLB_HD float2 mp() { return make_float2(-1.f, 1.f); }

GLB_HD float2 cadd(float2 a, float2 b) { return add2(a, b); }
GLB_HD float2 csub(float2 a, float2 b) { return sub2(a, b); }
GLB_HD float2 emul(float2 a, float2 b) { return mul2(a, b); }        // component-wise (window / taper)
GLB_HD float2 add_mi(float2 a, float2 b) { return fma2(swp(b), pm(), a); }   // a - i b
GLB_HD float2 add_pi(float2 a, float2 b) { return fma2(swp(b), mp(), a); }   // a + i b
GLB_HD float2 mul_mi(float2 a) { return mul2(swp(a), pm()); }        // a * (-i)
GLB_HD float2 mul_pi(float2 a) { return mul2(swp(a), mp()); }        // a * (+i)
// a * w and a * conj(w) with w as a plain (re, im) pair: three packed instructions
GLB_HD float2 cmul(float2 a, float2 w) { return fma2(mul2(swp(a), bc(w.y)), mp(), mul2(a, bc(w.x))); }
GLB_HD float2 cmulc(float2 a, float2 w) { return fma2(mul2(swp(a), bc(w.y)), pm(), mul2(a, bc(w.x))); }
// a * w with a compile-time constant w = (WR, WI): two packed instructions
#define GLB_CMULK(a, WR, WI) fma2((a), bc(WR), mul2(swp(a), make_float2(-(WI), (WI))))

// A twiddle kept in registers across frames: wr and the pair (-wi, wi); a * w and a * conj(w)
// are then two packed instructions each.
struct Tw3 {
  float r;
  float2 q;
};
GLB_HD Tw3 make_tw3(float2 w) {
  Tw3 t;
  t.r = w.x;
  t.q = make_float2(-w.y, w.y);
  return t;
}
GLB_HD float2 cmul(float2 a, const Tw3 &w) { return fma2(a, bc(w.r), mul2(swp(a), w.q)); }
GLB_HD float2 cmulc(float2 a, const Tw3 &w) { return fma2(a, bc(w.r), neg(mul2(swp(a), w.q))); }

// Padded float2 index: 1 spare entry after every 16, so that 64-bit accesses of 16 lanes that
// are 16 entries apart (first-pass stores) or 1 entry apart (loads) hit 16 different banks.
// pad(a + c) = pad(a) + c + c/16 whenever c is a multiple of 16: strided accesses become a
// per-thread base plus a compile-time immediate.  (128-bit stores with a pad of 2 need four
// register moves each to line the pairs up: slower.)
GLB_HD int pad(int p) { return p + (p >> 4); }
template <int M> struct BufSize { static constexpr int value = M + M / 8 + 2; };   // float2 entries

// Layout of the exchange that feeds the final radix-8 pass (plans with a mid pass): entry
// j + r 2T (butterfly j, input r) lives at 9 j + r.  The uniform stride of 9 entries keeps the
// ascending loads of butterfly t, the descending loads of butterfly 2T - t (whatever their
// alignment) and the stores of the pass before all bank-conflict free with 64-bit accesses.
#if !defined(GLB_LAST9)
#define GLB_LAST9 1
#endif
template <int M> struct LastLayout { static constexpr bool value = GLB_LAST9 && Plan<M>::NP >= 3; };
GLB_HD int last_phys(int j, int r) { return 9 * j + r; }

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
  float2 apc = add2(a, c), amc = sub2(a, c), bpd = add2(b, d), bmd = sub2(b, d);
  v[0] = add2(apc, bpd);
  v[S] = add_mi(amc, bmd);
  v[2 * S] = sub2(apc, bpd);
  v[3 * S] = add_pi(amc, bmd);
}

// 26 packed instructions: the W8 factors are folded into the last butterflies
template <int S> GLB_HD void dft8(float2 *v) {
  const float h = 0.70710678118654752440f;
  float2 e[4] = {v[0], v[2 * S], v[4 * S], v[6 * S]};
  float2 o[4] = {v[S], v[3 * S], v[5 * S], v[7 * S]};
  dft4<1>(e);
  dft4<1>(o);
  const float2 s1 = fma2(swp(o[1]), pm(), o[1]);        // (x + y, y - x)   = o1 W8^1 / h
  const float2 s3 = fma2(swp(o[3]), pm(), neg(o[3]));   // (y - x, -x - y)  = o3 W8^3 / h
  v[0] = add2(e[0], o[0]);
  v[4 * S] = sub2(e[0], o[0]);
  v[S] = fma2(s1, bc(h), e[1]);
  v[5 * S] = fma2(s1, bc(-h), e[1]);
  v[2 * S] = add_mi(e[2], o[2]);
  v[6 * S] = add_pi(e[2], o[2]);
  v[3 * S] = fma2(s3, bc(h), e[3]);
  v[7 * S] = fma2(s3, bc(-h), e[3]);
}

// 81 packed instructions
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
  y[1][1] = GLB_CMULK(y[1][1], c1, -s1);     // e = 1
  y[1][2] = GLB_CMULK(y[1][2], h, -h);       // e = 2
  y[1][3] = GLB_CMULK(y[1][3], s1, -c1);     // e = 3
  y[2][1] = GLB_CMULK(y[2][1], h, -h);       // e = 2
  y[2][2] = mul_mi(y[2][2]);                 // e = 4
  y[2][3] = GLB_CMULK(y[2][3], -h, -h);      // e = 6
  y[3][1] = GLB_CMULK(y[3][1], s1, -c1);     // e = 3
  y[3][2] = GLB_CMULK(y[3][2], -h, -h);      // e = 6
  y[3][3] = GLB_CMULK(y[3][3], -c1, s1);     // e = 9
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
  if constexpr (T % 16 == 0) return pad(t) + q * (T + T / 16);
  else return pad(t + T * q);
}

// Where the outputs of pass P go (Stockham): butterfly j = t + u T, k = j mod Ns, output r' is
// entry (j - k) R + k + r' Ns of the next pass; in buffer terms a per-thread base plus r' times a
// compile-time stride, for each of the three layouts (first pass, pass feeding the last, other).
template <int M, int P> GLB_HD int scatter_base(int t, int u) {
  constexpr int T = M / kPoints;
  constexpr int R = PlanRadix<M, P>::R, Ns = PlanRadix<M, P>::Ns;
  const int j = t + u * T;
  const int k = j & (Ns - 1);
  if constexpr (P == 0 && R == 16) return 17 * j;
  else if constexpr (LastLayout<M>::value && P == Plan<M>::NP - 2) return last_phys(k, (j - k) / Ns);
  else if constexpr (Ns % 16 == 0) return pad((j - k) * R + k);
  else return (j - k) * R + k;                      // unpadded natural index: use scatter_index()
}
template <int M, int P> GLB_HD int scatter_index(int base, int r) {
  constexpr int R = PlanRadix<M, P>::R, Ns = PlanRadix<M, P>::Ns;
  if constexpr (P == 0 && R == 16) return base + r;
  else if constexpr (LastLayout<M>::value && P == Plan<M>::NP - 2) return base + r * 9 * Ns;
  else if constexpr (Ns % 16 == 0) return base + r * (Ns + Ns / 16);
  else return pad(base + r * Ns);
}

template <int M>
GLB_HD void pass_load(float2 *v, int t, const float2 *buf) {
#pragma unroll
  for (int q = 0; q < kPoints; q++) v[q] = buf[chk<BufSize<M>::value>(ld_index<M>(t, q))];
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
  if constexpr (LastLayout<M>::value) {
    const int bA = last_phys(jA, 0), bB = last_phys(jB, 0);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = buf[bA + r];
      v[8 + r] = buf[bB + r];
    }
  } else if constexpr (Ns % 16 == 0) {
    const int bA = pad(jA), bB = pad(jB);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      v[r] = buf[bA + r * (Ns + Ns / 16)];
      v[8 + r] = buf[bB + r * (Ns + Ns / 16)];
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
  float2 p = fma2(zm, pm(), zk);      // (zk.x + zm.x, zk.y - zm.y)
  float2 q = fma2(zm, mp(), zk);      // (zk.x - zm.x, zk.y + zm.y)
  float2 vq = cmul(q, vk);
  a = add2(p, vq);
  b = sub2(p, vq);
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

// first bin of the upper four pairs of thread t (see "bins of the final pass" below)
template <int M> GLB_HD int khi(int t) {
  constexpr int T = M / kPoints;
  return t != 0 ? t + 8 * T : T;
}

struct TwRegs {
  Tw3 mid[2][12];   // passes 1 and 2 (when they are not the last): [u * nbases + i]
  Tw3 last[4];      // w_a^1..w_a^4, w_a = W_M^t
  Tw3 v0;           // V_t = -i exp(-2 pi i t / N): split factor of the pairs rp < 4 (times W_16^rp)
  Tw3 v0hi;         // V_khi(t): pairs rp >= 4 (times W_16^(rp-4))
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
        for (int i = 0; i < 6; i++) tr.mid[P - 1][u * NB + i] = make_tw3(twp[(e[i] - 1) * Ns]);
      } else {
#pragma unroll
        for (int i = 0; i < NB; i++) tr.mid[P - 1][u * NB + i] = make_tw3(twp[i * Ns]);
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
  for (int i = 0; i < 4; i++) tr.last[i] = make_tw3(twA[i * 2 * T]);
  tr.v0 = make_tw3(vtab[t]);
  tr.v0hi = make_tw3(vtab[khi<M>(t)]);
}

// the last pass's bases and the two split factors only (table-twiddle plans load them per
// transform: 6 loads instead of the 14 twiddles + 8 split factors of the plain table path)
template <int M>
GLB_HD void load_last_regs(TwRegs &tr, int t, const float2 *tw, const float2 *vtab) {
  constexpr int T = M / kPoints, NP = Plan<M>::NP;
  const float2 *twA = tw + TwOffset<M, NP - 1>::value + t;
#pragma unroll
  for (int i = 0; i < 4; i++) tr.last[i] = make_tw3(twA[i * 2 * T]);
  tr.v0 = make_tw3(vtab[t]);
  tr.v0hi = make_tw3(vtab[khi<M>(t)]);
}

// v[u + r S] *= w^r for r = 1..R-1 from the kept bases; w^(4a+b) is applied as two successive
// multiplications (w^(4a) then w^b): no derived twiddle is ever formed
template <int R, int S>
GLB_HD void apply_tw_bases(float2 *v, const Tw3 *b) {
  if constexpr (R == 16) {
    const Tw3 w1 = b[0], w2 = b[1], w3 = b[2], w4 = b[3], w8 = b[4], w12 = b[5];
#pragma unroll
    for (int l = 0; l < 4; l++) {
      v[(4 + l) * S] = cmul(v[(4 + l) * S], w4);
      v[(8 + l) * S] = cmul(v[(8 + l) * S], w8);
      v[(12 + l) * S] = cmul(v[(12 + l) * S], w12);
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      v[(4 * a + 1) * S] = cmul(v[(4 * a + 1) * S], w1);
      v[(4 * a + 2) * S] = cmul(v[(4 * a + 2) * S], w2);
      v[(4 * a + 3) * S] = cmul(v[(4 * a + 3) * S], w3);
    }
  } else if constexpr (R == 8) {
    const Tw3 w1 = b[0], w2 = b[1], w3 = b[2], w4 = b[3];
#pragma unroll
    for (int l = 0; l < 4; l++) v[(4 + l) * S] = cmul(v[(4 + l) * S], w4);
#pragma unroll
    for (int a = 0; a < 2; a++) {
      v[(4 * a + 1) * S] = cmul(v[(4 * a + 1) * S], w1);
      v[(4 * a + 2) * S] = cmul(v[(4 * a + 2) * S], w2);
      v[(4 * a + 3) * S] = cmul(v[(4 * a + 3) * S], w3);
    }
  } else if constexpr (R == 4) {
    v[1 * S] = cmul(v[1 * S], b[0]);
    v[2 * S] = cmul(v[2 * S], b[1]);
    v[3 * S] = cmul(v[3 * S], b[2]);
  } else if constexpr (R == 2) {
    v[1 * S] = cmul(v[1 * S], b[0]);
  }
}

// pass_store<M,P>: twiddle (P > 0), radix-R butterflies on the 16 register points and Stockham store
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
    if constexpr (P > 0) {
      // only the base powers are loaded (6 of 15 for radix 16, 4 of 7 for radix 8); the others are
      // applied as two successive multiplications, as with register twiddles
      const float2 *twp = tw + TwOffset<M, P>::value + k;
      constexpr int NB = TwBases<R>::n;
      Tw3 b[NB > 0 ? NB : 1];
      if constexpr (R == 16) {
        const int e[6] = {1, 2, 3, 4, 8, 12};
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = make_tw3(twp[(e[i] - 1) * Ns]);
      } else {
#pragma unroll
        for (int i = 0; i < NB; i++) b[i] = make_tw3(twp[i * Ns]);
      }
      apply_tw_bases<R, S>(v + u, b);
    }
    Dft<R, S>::run(v + u);
    const int base = scatter_base<M, P>(t, u);
#pragma unroll
    for (int r = 0; r < R; r++) buf[chk<BufSize<M>::value>(scatter_index<M, P>(base, r))] = v[u + r * S];
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
  constexpr int R = PlanRadix<M, P>::R, S = kPoints / R;
#pragma unroll
  for (int u = 0; u < S; u++) {
    const int base = scatter_base<M, P>(t, u);
#pragma unroll
    for (int r = 0; r < R; r++) buf[chk<BufSize<M>::value>(scatter_index<M, P>(base, r))] = v[u + r * S];
  }
}

// Two frames at once: the buffer holds float4 entries (frame A's value, frame B's value) at the
// same entry indices, so every exchange is a 128-bit access (quarter-warps of 8 lanes: the
// layouts above stay conflict free, 17 j, 9 k and consecutive entries all being distinct mod 8).
template <int M, int P>
GLB_HD void pass_scatter2(const float2 *va, const float2 *vb, int t, float4 *buf) {
  constexpr int R = PlanRadix<M, P>::R, S = kPoints / R;
#pragma unroll
  for (int u = 0; u < S; u++) {
    const int base = scatter_base<M, P>(t, u);
#pragma unroll
    for (int r = 0; r < R; r++)
      buf[chk<BufSize<M>::value>(scatter_index<M, P>(base, r))] = make_float4(va[u + r * S].x, va[u + r * S].y, vb[u + r * S].x, vb[u + r * S].y);
  }
}
template <int M>
GLB_HD void pass_load2(float2 *va, float2 *vb, int t, const float4 *buf) {
#pragma unroll
  for (int q = 0; q < kPoints; q++) {
    const float4 e = buf[chk<BufSize<M>::value>(ld_index<M>(t, q))];
    va[q] = make_float2(e.x, e.y);
    vb[q] = make_float2(e.z, e.w);
  }
}

GLB_HD float2 w16_mul(float2 q, int e) {
  // q * exp(-2 pi i e / 16), e = 0..7, constants folded after unrolling
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  switch (e) {
    case 0: return q;
    case 1: return GLB_CMULK(q, c1, -s1);
    case 2: return GLB_CMULK(q, h, -h);
    case 3: return GLB_CMULK(q, s1, -c1);
    case 4: return mul_mi(q);
    case 5: return GLB_CMULK(q, -s1, -c1);
    case 6: return GLB_CMULK(q, -h, -h);
    default: return GLB_CMULK(q, -c1, -s1);
  }
}
GLB_HD float2 w16_mulc(float2 q, int e) {
  // q * exp(+2 pi i e / 16)
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  switch (e) {
    case 0: return q;
    case 1: return GLB_CMULK(q, c1, s1);
    case 2: return GLB_CMULK(q, h, h);
    case 3: return GLB_CMULK(q, s1, c1);
    case 4: return mul_pi(q);
    case 5: return GLB_CMULK(q, -s1, c1);
    case 6: return GLB_CMULK(q, -h, h);
    default: return GLB_CMULK(q, -c1, s1);
  }
}

// butterflies of the last pass on loaded inputs (v[0..7] butterfly A = t, v[8..15] butterfly B)
template <int M>
GLB_HD void last_pass_compute_rt(float2 *v, int t, const TwRegs &tr) {
  const Tw3 w1 = tr.last[0], w2 = tr.last[1], w3 = tr.last[2], w4 = tr.last[3];
  // A: x_r w_a^r, w_a^(4+b) applied as w_a^4 then w_a^b
#pragma unroll
  for (int l = 0; l < 4; l++) v[4 + l] = cmul(v[4 + l], w4);
#pragma unroll
  for (int a = 0; a < 2; a++) {
    v[4 * a + 1] = cmul(v[4 * a + 1], w1);
    v[4 * a + 2] = cmul(v[4 * a + 2], w2);
    v[4 * a + 3] = cmul(v[4 * a + 3], w3);
  }
  dft8<1>(v);
  // B = butterfly 2T - t: x_r conj(w_a^r), DFT8, outputs shifted by one place (the W_8^r factor).
  // Thread 0 holds butterfly T instead (twiddles W_16^r, w_a = 1): pre-multiplying its inputs by
  // conj(W_16^r) = W_16^r W_8^-r lets it share the code below, shift included.
  float2 y[8];
#pragma unroll
  for (int r = 0; r < 8; r++) y[r] = v[8 + r];
  if (t == 0) {
#pragma unroll
    for (int r = 1; r < 8; r++) y[r] = w16_mulc(y[r], r);
  }
#pragma unroll
  for (int l = 0; l < 4; l++) y[4 + l] = cmulc(y[4 + l], w4);
#pragma unroll
  for (int a = 0; a < 2; a++) {
    y[4 * a + 1] = cmulc(y[4 * a + 1], w1);
    y[4 * a + 2] = cmulc(y[4 * a + 2], w2);
    y[4 * a + 3] = cmulc(y[4 * a + 3], w3);
  }
  dft8<1>(y);
#pragma unroll
  for (int r = 0; r < 8; r++) v[8 + r] = y[(r + 1) & 7];
}

// buffer index of input r of butterfly j of the last pass
template <int M> GLB_HD int last_index(int j, int r) {
  constexpr int T = M / kPoints, Ns = 2 * T;
  if constexpr (LastLayout<M>::value) return last_phys(j, r);
  else if constexpr (Ns % 16 == 0) return pad(j) + r * (Ns + Ns / 16);
  else return pad(j + r * Ns);
}
template <int M>
GLB_HD void last_pass_load(float2 *v, int t, const float2 *buf) {
  constexpr int T = M / kPoints;
  const int jA = t;
  const int jB = (t == 0) ? T : 2 * T - t;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    v[r] = buf[chk<BufSize<M>::value>(last_index<M>(jA, r))];
    v[8 + r] = buf[chk<BufSize<M>::value>(last_index<M>(jB, r))];
  }
}
template <int M>
GLB_HD void last_pass_load2(float2 *va, float2 *vb, int t, const float4 *buf) {
  constexpr int T = M / kPoints;
  const int jA = t;
  const int jB = (t == 0) ? T : 2 * T - t;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const float4 a = buf[chk<BufSize<M>::value>(last_index<M>(jA, r))], b = buf[chk<BufSize<M>::value>(last_index<M>(jB, r))];
    va[r] = make_float2(a.x, a.y);
    vb[r] = make_float2(a.z, a.w);
    va[8 + r] = make_float2(b.x, b.y);
    vb[8 + r] = make_float2(b.z, b.w);
  }
}

// last pass with register twiddles; on return v[r'] = Z[jA + r' 2T], v[8 + r'] = Z[jB + r' 2T]
template <int M>
GLB_HD void last_pass_rt(float2 *v, int t, const float2 *buf, const float2 *tw, const TwRegs &tr) {
  last_pass_load<M>(v, t, buf);
  last_pass_compute_rt<M>(v, t, tr);
}


// ------------------------------------------------------------------ bins of the final pass
// After the last pass thread t >= 1 holds butterflies A = t and B = 2T - t: pair rp (0..7) is
// (Z[k], Z[M-k]) with k = t + rp 2T in (v[rp], v[15 - rp]).  Thread 0 holds the two self-paired
// butterflies 0 and T; thread0_fixup() re-orders its registers so that the same 8 (v[rp],
// v[15 - rp]) pairs are (Z[k], Z[M-k]) for k = rp 2T (rp < 4) and k = T + (rp - 4) 2T (rp >= 4),
// and returns the one value left over, Z[M/2].  All threads then run the same split code; only
// the base index of the upper four pairs differs (khi).  Results are numbered by "slot": slot
// 2 rp is bin k, slot 2 rp + 1 is bin M - k, slot 16 (thread 0 only) is bin M/2, so accumulators
// can live in registers across tapers.  Every bin 0..M belongs to exactly one (thread, slot).
template <int M> GLB_HD int slot_bin(int t, int slot) {
  constexpr int T = M / kPoints;
  if (slot == 16) return M / 2;
  const int rp = slot >> 1;
  const int k = rp < 4 ? t + rp * 2 * T : khi<M>(t) + (rp - 4) * 2 * T;
  return (slot & 1) ? M - k : k;
}
template <int M> GLB_HD int slot_count(int t) { return t != 0 ? 16 : 17; }

GLB_HD float2 thread0_fixup(float2 *v) {
  const float2 a4 = v[4];
  const float2 a5 = v[5], a6 = v[6], a7 = v[7], a0 = v[0];
#pragma unroll
  for (int r = 0; r < 8; r++) v[4 + r] = v[8 + r];      // B0..B7 -> v[4..11]
  v[12] = a5;
  v[13] = a6;
  v[14] = a7;
  v[15] = a0;
  return a4;
}

// split with V_k = vb W_16^e rebuilt from a kept base factor
GLB_HD void split_pair_rt(float2 zk, float2 zm, const Tw3 &vb, int e, float2 &a, float2 &b) {
  float2 p = fma2(zm, pm(), zk);      // (zk.x + zm.x, zk.y - zm.y)
  float2 q = fma2(zm, mp(), zk);      // (zk.x - zm.x, zk.y + zm.y)
  float2 vq = cmul(w16_mul(q, e), vb);
  a = add2(p, vq);
  b = sub2(p, vq);
}

// f(slot, value, conjugated): value = 2 X[slot_bin(t, slot)] (conjugated when the flag is set),
// still carrying the scale folded into the taper.
// T0 = false: the caller knows that t != 0 (warps without thread 0 then carry no trace of the
// re-ordering: the compiler needs ~17 register moves around it)
template <int M, bool T0 = true, class F>
GLB_HD void emit_bins_rt(float2 *v, int t, const TwRegs &tr, F &&f) {
  float2 a, b;
  if (T0 && t == 0) {
    const float2 z = thread0_fixup(v);                  // X[M/2] = conj(Z[M/2])
    f(16, make_float2(2.f * z.x, -2.f * z.y), false);
  }
#pragma unroll
  for (int rp = 0; rp < 8; rp++) {
    if (rp < 4) split_pair_rt(v[rp], v[15 - rp], tr.v0, rp, a, b);
    else split_pair_rt(v[rp], v[15 - rp], tr.v0hi, rp - 4, a, b);
    f(2 * rp, a, false);
    f(2 * rp + 1, b, true);
  }
}

template <int M, bool T0 = true, class F>
GLB_HD void emit_bins(float2 *v, int t, const float2 *vtab, F &&f) {
  constexpr int T = M / kPoints;
  float2 a, b;
  if (T0 && t == 0) {
    const float2 z = thread0_fixup(v);
    f(16, make_float2(2.f * z.x, -2.f * z.y), false);
  }
  const int kh = khi<M>(t);
#pragma unroll
  for (int rp = 0; rp < 8; rp++) {
    const int k = rp < 4 ? t + rp * 2 * T : kh + (rp - 4) * 2 * T;
    split_pair(v[rp], v[15 - rp], vtab[k], a, b);
    f(2 * rp, a, false);
    f(2 * rp + 1, b, true);
  }
}

}  // namespace glb
