// fft_big.cuh -- the 32-points-per-thread FFT core for the big frames (N = 16384, 32768).
//
// Same construction as fft_core.cuh (real FFT of length N as a complex FFT of length M = N/2 over
// z[m] = x[2m] + i x[2m+1], Stockham passes through one shared-memory buffer, last pass arranged so
// that Z[k] and Z[M-k] meet in one thread) with twice the points per thread:
//   T = M/32 threads per frame, 32 complex points each, THREE passes (32, M/512, 16) instead of the
//   four the 16-point core needs at these sizes, i.e. two exchanges instead of three -- the
//   shared-memory pipe is what bounds the 16-point core here;
//   256 threads per frame at N = 16384, so two CTAs (two frames) fit one SM.
// Layouts (float2 entries; all accesses 64-bit, per-thread base + compile-time immediates):
//   exchange 0   logical e -> e + (e >> 5): butterfly t of the radix-32 pass stores its 32 outputs at
//                33 t + r (lanes 33 entries apart: conflict free), pass 1 loads t + T q (unit stride);
//   exchange 1   input r of butterfly j of the last (radix-16) pass lives at 17 j + r: the ascending
//                loads of butterfly t, the descending loads of butterfly 2T - t and the stores of the
//                pass before are all 17 entries apart between lanes: conflict free whatever the alignment.
// Everything is __host__ __device__ so tests/emu/emu_big.cpp runs the exact index and twiddle
// arithmetic on the CPU.  This replaces (does not port) fft_radix2.c:75-177; the only contract kept
// is the DFT and the PSD layout of fft_psd (fft.c:203-217).
#pragma once
#include "fft_core.cuh"
#include "twiddle_consts.cuh"

namespace glb {

constexpr int kBP = 32;              // complex points per thread

template <int M> struct Big {
  static_assert(M == 2048 || M == 4096 || M == 8192 || M == 16384, "32-point core: N = 4096 .. 32768");
  static constexpr int T = M / kBP;              // threads per frame
  static constexpr int R1 = M / (32 * 16);       // radix of the middle pass (4, 8, 16, 32)
  static constexpr int S1 = kBP / R1;            // its butterflies per thread
  static constexpr int BUF = 17 * (M / 16) + 16; // float2 entries (the second layout is the larger one)
  static constexpr int TW1 = (R1 - 1) * 32;      // middle-pass twiddles: [(r - 1) * 32 + k] = exp(-2 pi i k r / (32 R1))
};

// a * w and a * conj(w) for a w that is a compile-time constant after unrolling: two packed instructions
GLB_HD float2 cmulk(float2 a, float2 w) { return fma2(a, bc(w.x), mul2(swp(a), make_float2(-w.y, w.y))); }
GLB_HD float2 cmulkc(float2 a, float2 w) { return fma2(a, bc(w.x), mul2(swp(a), make_float2(w.y, -w.y))); }
// q * exp(-2 pi i e / 32)
GLB_HD float2 mul_w32(float2 q, int e) {
  e &= 31;
  if (e == 0) return q;
  if (e == 8) return mul_mi(q);
  if (e == 16) return neg(q);
  if (e == 24) return mul_pi(q);
  return cmulk(q, wconst32(e));
}

// 32-point DFT over v[0], v[S], ..., v[31 S], natural order in place: n = 8 n1 + n2, k = k1 + 4 k2
template <int S> GLB_HD void dft32(float2 *v) {
  float2 y[32];
#pragma unroll
  for (int n2 = 0; n2 < 8; n2++) {
    float2 t[4] = {v[n2 * S], v[(8 + n2) * S], v[(16 + n2) * S], v[(24 + n2) * S]};
    dft4<1>(t);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) y[8 * k1 + n2] = mul_w32(t[k1], n2 * k1);
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    dft8<1>(y + 8 * k1);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) v[(k1 + 4 * k2) * S] = y[8 * k1 + k2];
  }
}
// (fft_wpf.cuh owns Dft<32, S>, with three-instruction constant multiplies: this core has its own dispatch)
template <int R, int S> struct BigDft { static GLB_HD void run(float2 *v) { Dft<R, S>::run(v); } };
template <int S> struct BigDft<32, S> { static GLB_HD void run(float2 *v) { dft32<S>(v); } };

// ---------------------------------------------------------------- passes
// pass 0: radix 32 over v[q] = z[t + T q], outputs to exchange 0
GLB_HD void big_pass0(float2 *v) { dft32<1>(v); }

template <int M> GLB_HD void big_scatter0(const float2 *v, int t, float2 *buf) {
#pragma unroll
  for (int r = 0; r < kBP; r++) buf[chk<Big<M>::BUF>(33 * t + r)] = v[r];
}

template <int M> GLB_HD void big_load1(float2 *v, int t, const float2 *buf) {
  constexpr int T = Big<M>::T;
  const int base = t + (t >> 5);
#pragma unroll
  for (int q = 0; q < kBP; q++) v[q] = buf[chk<Big<M>::BUF>(base + q * (T + T / 32))];
}

// pass 1: S1 butterflies of radix R1; butterfly u works on v[u + r S1]; twiddle exp(-2 pi i k r / (32 R1)),
// k = t mod 32 (the same for every butterfly of the thread), from a table the lanes of a warp read
// at unit stride
template <int M> GLB_HD void big_pass1(float2 *v, int t, const float2 *tw1) {
  constexpr int R1 = Big<M>::R1, S1 = Big<M>::S1;
  const float2 *tk = tw1 + (t & 31);
#pragma unroll
  for (int r = 1; r < R1; r++) {
    const Tw3 w = make_tw3(tk[(r - 1) * 32]);
#pragma unroll
    for (int u = 0; u < S1; u++) v[u + r * S1] = cmul(v[u + r * S1], w);
  }
#pragma unroll
  for (int u = 0; u < S1; u++) BigDft<R1, S1>::run(v + u);
}

// output r' of butterfly j = t + u T is input (t >> 5) + u R1/2 of last-pass butterfly (t & 31) + 32 r'
template <int M> GLB_HD void big_scatter1(const float2 *v, int t, float2 *buf) {
  constexpr int R1 = Big<M>::R1, S1 = Big<M>::S1;
  const int base = 17 * (t & 31) + (t >> 5);
#pragma unroll
  for (int u = 0; u < S1; u++)
#pragma unroll
    for (int r = 0; r < R1; r++) buf[chk<Big<M>::BUF>(base + 17 * 32 * r + u * (R1 / 2))] = v[u + r * S1];
}

// last pass, radix 16: butterfly A = t in v[0..15], butterfly B = 2T - t (thread 0: T) in v[16..31]
template <int M> GLB_HD void big_load2(float2 *v, int t, const float2 *buf) {
  constexpr int T = Big<M>::T;
  const int jB = (t == 0) ? T : 2 * T - t;
#pragma unroll
  for (int r = 0; r < 16; r++) {
    v[r] = buf[chk<Big<M>::BUF>(17 * t + r)];
    v[16 + r] = buf[chk<Big<M>::BUF>(17 * jB + r)];
  }
}

// bases w^1 w^2 w^3 w^4 w^8 w^12 of w = W_M^t and the two split factors, loaded per transform
struct BigLast {
  Tw3 w[6];
  Tw3 v0, v0hi;     // V_t, V_khi(t): V_k = -i exp(-2 pi i k / N)
};
template <int M> GLB_HD int big_khi(int t) { return t != 0 ? t + 16 * Big<M>::T : Big<M>::T; }

template <int M> GLB_HD void big_load_last(BigLast &L, int t, const float2 *roots, const float2 *vtab, bool bases = true) {
  const int e[6] = {1, 2, 3, 4, 8, 12};
  if (bases) {
#pragma unroll
    for (int i = 0; i < 6; i++) L.w[i] = make_tw3(roots[t * e[i]]);
  }
  L.v0 = make_tw3(vtab[t]);
  L.v0hi = make_tw3(vtab[big_khi<M>(t)]);
}

// On return v[r'] = Z[t + r' 2T] and v[16 + r'] = Z[jB + r' 2T].
//   A: x_r w^r (w^(4a+b) applied as w^(4a) then w^b), DFT16.
//   B: W_M^(jB r) = W_16^r conj(w^r): conjugate twiddles, DFT16, outputs shifted by one place.  Thread 0
//      holds butterfly T (twiddle W_32^r, w = 1): pre-multiplying its inputs by conj(W_32^r) =
//      W_32^r W_16^-r lets it run the same code, shift included.
template <int M> GLB_HD void big_pass2(float2 *v, int t, const BigLast &L) {
  apply_tw_bases<16, 1>(v, L.w);
  dft16<1>(v);
  float2 y[16];
#pragma unroll
  for (int r = 0; r < 16; r++) y[r] = v[16 + r];
  if (t == 0) {
#pragma unroll
    for (int r = 1; r < 16; r++) y[r] = mul_w32(y[r], 32 - r);
  }
  const Tw3 w1 = L.w[0], w2 = L.w[1], w3 = L.w[2], w4 = L.w[3], w8 = L.w[4], w12 = L.w[5];
#pragma unroll
  for (int l = 0; l < 4; l++) {
    y[4 + l] = cmulc(y[4 + l], w4);
    y[8 + l] = cmulc(y[8 + l], w8);
    y[12 + l] = cmulc(y[12 + l], w12);
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    y[4 * a + 1] = cmulc(y[4 * a + 1], w1);
    y[4 * a + 2] = cmulc(y[4 * a + 2], w2);
    y[4 * a + 3] = cmulc(y[4 * a + 3], w3);
  }
  dft16<1>(y);
#pragma unroll
  for (int r = 0; r < 16; r++) v[16 + r] = y[(r + 1) & 15];
}

// The same pass with every power w^r = W_M^(t r), r = 1..15, read from a table [15][T] (shared memory, built
// once per CTA from the M-th roots: each entry rounded once, and 18 complex multiplications fewer per thread
// than rebuilding the powers from six bases); one load serves butterfly A (w^r) and butterfly B (conj w^r).
template <int M> GLB_HD void big_pass2_tab(float2 *v, int t, const float2 *tab) {
  constexpr int T = Big<M>::T;
  if (t == 0) {
#pragma unroll
    for (int r = 1; r < 16; r++) v[16 + r] = mul_w32(v[16 + r], 32 - r);
  }
#pragma unroll
  for (int r = 1; r < 16; r++) {
    const Tw3 w = make_tw3(tab[(r - 1) * T + t]);
    v[r] = cmul(v[r], w);
    v[16 + r] = cmulc(v[16 + r], w);
  }
  dft16<1>(v);
  float2 y[16];
#pragma unroll
  for (int r = 0; r < 16; r++) y[r] = v[16 + r];
  dft16<1>(y);
#pragma unroll
  for (int r = 0; r < 16; r++) v[16 + r] = y[(r + 1) & 15];
}

// ---------------------------------------------------------------- bins
// Thread t >= 1: pair rp (0..15) is (Z[k], Z[M-k]) = (v[rp], v[31 - rp]) with k = t + rp 2T.  Thread 0 holds
// the self-paired butterflies 0 and T; big_thread0_fixup() re-orders its registers so that the same 16
// (v[rp], v[31 - rp]) pairs are (Z[k], Z[M-k]) with k = rp 2T for rp < 8 and k = T + (rp - 8) 2T for rp >= 8,
// and returns the value left over, Z[M/2].  Slot 2 rp is bin k, slot 2 rp + 1 is bin M - k, slot 32 (thread 0
// only) is bin M/2: every bin 0..M belongs to exactly one (thread, slot).
template <int M> GLB_HD int big_slot_bin(int t, int slot) {
  constexpr int T = Big<M>::T;
  if (slot == 32) return M / 2;
  const int rp = slot >> 1;
  const int k = rp < 8 ? t + rp * 2 * T : big_khi<M>(t) + (rp - 8) * 2 * T;
  return (slot & 1) ? M - k : k;
}

GLB_HD float2 big_thread0_fixup(float2 *v) {
  // in: v[0..15] = A0..A15 (butterfly 0), v[16..31] = B0..B15 (butterfly T)
  float2 a[16], b[16];
#pragma unroll
  for (int r = 0; r < 16; r++) {
    a[r] = v[r];
    b[r] = v[16 + r];
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    v[8 + r] = b[r];                                    // rp = 8 + r: Z[T + r 2T]
    v[16 + r] = b[8 + r];                               // its partner v[31 - rp] = B[15 - r]
  }
#pragma unroll
  for (int r = 1; r < 8; r++) v[31 - r] = a[16 - r];    // rp = r: partner of A[r] is A[16 - r]
  v[31] = a[0];                                         // Z[0] pairs with itself (bins 0 and M)
  return a[8];
}

// f(slot, value, conjugated): value = 2 X[big_slot_bin(t, slot)] (its conjugate when the flag is set), still
// carrying the scale folded into the taper
template <int M, bool T0 = true, class F>
GLB_HD void big_emit(float2 *v, int t, const BigLast &L, F &&f) {
  if (T0 && t == 0) {
    const float2 z = big_thread0_fixup(v);              // X[M/2] = conj(Z[M/2])
    f(32, make_float2(2.f * z.x, -2.f * z.y), false);
  }
#pragma unroll
  for (int rp = 0; rp < 16; rp++) {
    const float2 zk = v[rp], zm = v[31 - rp];
    const float2 p = fma2(zm, pm(), zk);                // (zk.x + zm.x, zk.y - zm.y)
    const float2 q = fma2(zm, mp(), zk);                // (zk.x - zm.x, zk.y + zm.y)
    // V_k = V_base W_32^rp (rp < 8) or V_khi W_32^(rp - 8)
    const float2 vq = cmul(mul_w32(q, rp < 8 ? rp : rp - 8), rp < 8 ? L.v0 : L.v0hi);
    f(2 * rp, add2(p, vq), false);
    f(2 * rp + 1, sub2(p, vq), true);
  }
}

}  // namespace glb
