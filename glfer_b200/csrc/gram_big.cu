// gram_big.cu -- the spectrogram kernel for the big frames (N = 16384, 32768) on the 32-points-per-thread
// FFT core (fft_big.cuh): overlapped-frame gather -> block-mean removal -> taper -> real FFT in three
// passes (two shared-memory exchanges) -> |X|^2 -> [10 log10] -> one PSD row per frame.
//
// Why a second kernel family: at these sizes the 16-point core needs four passes, 512 threads and 74 KB of
// exchange buffer per frame, so with its 64 KB TMA ring only ONE frame fits an SM and the shared-memory
// pipe (three exchanges) bounds it (N = 16384: 34 % of the HBM roofline, round 1).  Here a frame is 256
// threads x 32 points (N = 16384), 72 KB of shared memory and no ring: a frame's samples land by ONE TMA
// bulk copy in the exchange buffer itself while it is idle (between the last-pass loads of the previous
// frame and the first-pass stores of this one); its older hop blocks were fetched by this very CTA one
// frame earlier and come from L2, so every sample still crosses HBM once.  Two CTAs (two frames in
// flight) fit an SM, with a third fewer shared-memory wavefronts per frame.
//
// Geometry: hop = N / NBLK, NBLK = 1, 2, 4 (0 %, 50 %, 75 % overlap), periodogram or multitaper, no RA9MB /
// limiter; everything else stays on the 16-point kernels (gram_common.cuh).
#include "gram_common.cuh"
#include "fft_big.cuh"

int g_big_pair = 1;                  // 0: never the two-groups-per-CTA variant (experiments)

template <int M> struct BigGeo {
  static constexpr int T = Big<M>::T, N = 2 * M, NW = T / 32;
  static constexpr size_t BUF_BYTES = (size_t) Big<M>::BUF * sizeof(float2);
  static constexpr size_t TW_BYTES = (size_t) Big<M>::TW1 * sizeof(float2);
  static constexpr int MINB = 512 / T;            // CTAs per SM at 128 registers per thread
  static constexpr size_t RED_BYTES = (size_t) 4 * NW * sizeof(float);
  static constexpr size_t ACC_BYTES = (size_t) 33 * T * sizeof(float);         // multitaper: the row being summed
  static constexpr size_t TW2_BYTES = (size_t) 15 * T * sizeof(float2);       // last-pass powers W_M^(t r), r = 1..15
  // the last-pass table is used wherever it fits beside the rest (not at N = 32768 multitaper)
#ifndef GLB_BIG_TW2
#define GLB_BIG_TW2 1
#endif
  // N = 32768 multitaper (one 512-thread CTA per SM): the row being summed lives in TENSOR MEMORY, not in shared
  // memory -- 33 words per thread written and read back with tcgen05.st / tcgen05.ld (32x32b: a thread owns a
  // lane, its words are columns) -- which takes 66 accesses per thread and taper off the shared-memory pipe and
  // frees the 68 KB that the last-pass table (61 KB) needs
#ifndef GLB_BIG_TMEM_ACC
#define GLB_BIG_TMEM_ACC 1
#endif
  static constexpr bool TACC = GLB_BIG_TMEM_ACC && (T == 512);
  static constexpr int TACC_COLS = 256;           // 4 column groups (warp / 4) of 40 columns, a power of two allocated
  static __host__ __device__ constexpr size_t acc_bytes(bool multi) { return (multi && !TACC) ? ACC_BYTES : 0; }
  static __host__ __device__ constexpr bool tw2(bool multi) { return GLB_BIG_TW2 && BUF_BYTES + TW_BYTES + RED_BYTES + 16 + acc_bytes(multi) + TW2_BYTES <= (size_t) 227 * 1024 / MINB - 1024; }
  static __host__ __device__ constexpr size_t smem(bool multi) { return BUF_BYTES + TW_BYTES + RED_BYTES + 16 + acc_bytes(multi) + (tw2(multi) ? TW2_BYTES : 0); }
  // PAIR: one CTA of 2 T threads = two frame groups (own exchange buffer, reduction scratch and mbarrier each,
  // own named barrier) that share the twiddle tables and ONE copy of the first half of the taper in shared memory
  static constexpr size_t GROUP_BYTES = BUF_BYTES + RED_BYTES + 16;
  static constexpr size_t HALF_TAPER_BYTES = (size_t) (M / 2) * sizeof(float2);
  static constexpr size_t PAIR_SMEM = 2 * GROUP_BYTES + TW_BYTES + TW2_BYTES + HALF_TAPER_BYTES;
  static constexpr bool PAIR_FITS = GLB_BIG_TW2 && T == 256 && PAIR_SMEM <= (size_t) 227 * 1024;   // N = 16384
};

// the per-warp partial sums of a block, added as a tree from vector loads (not a chain of dependent loads)
template <int NW> __device__ __forceinline__ float big_sum_warps(const float *r) {
  float a[NW];
  if constexpr (NW % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NW / 4; i++) {
      const float4 x = reinterpret_cast<const float4 *>(r)[i];
      a[4 * i] = x.x; a[4 * i + 1] = x.y; a[4 * i + 2] = x.z; a[4 * i + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NW; i++) a[i] = r[i];
  }
#pragma unroll
  for (int w = NW / 2; w >= 1; w /= 2) {
#pragma unroll
    for (int i = 0; i < w; i++) a[i] += a[i + w];
  }
  return a[0];
}

// tensor memory as per-thread scratch: 16 words of this thread's lane from / to consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t addr, const float *v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
                 "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float *v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                 "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t addr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t addr) {
  float v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// MULTI: Thomson multitaper (mtm_do, mtm.c:189-220): the frame goes through the transform once per taper -- its
// samples land again by TMA from L2 each time, the block means are formed once -- and the eigenspectra
// (1 / lambda_k folded into the tapers) are summed in a shared-memory row [33][T], so that the register
// budget of the periodogram kernel (two frames per SM at N = 16384) holds for the multitaper one too.
// LEV: 8-bit display levels (levels.cuh) beside / instead of the float rows, pixel i = bin M - i.
// PAIR (periodogram with a symmetric taper, w[i] == w[N - 1 - i] bit for bit -- every window of fft.c:37-82
// is): two frame groups per CTA, one per half of the threads, synchronised by a named barrier each; the first
// half of the taper lives in shared memory once per SM and both halves of a frame are read from it (the
// second half mirrored), instead of 64 KB per frame through an L1 that 2 x 104 KB of shared memory leave too
// small for it (11 % hit rate).
template <int M, int NBLK, bool MULTI, bool LEV, bool PAIR>
__global__ void __launch_bounds__(Big<M>::T * (PAIR ? 2 : 1), PAIR ? 1 : BigGeo<M>::MINB) gram_big_kernel(const KParams p) {
  using G = BigGeo<M>;
  constexpr int T = G::T, N = G::N, NW = G::NW, QB = kBP / NBLK;     // QB registers (float2) per hop block
  constexpr int HOP = N / NBLK;
  static_assert(!PAIR || (!MULTI && G::PAIR_FITS), "pair variant: periodogram only");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int g = PAIR ? (int) threadIdx.x / T : 0;
  const int t = PAIR ? (int) threadIdx.x % T : (int) threadIdx.x;
  // PAIR: [group 0: buf | red | mbar][group 1: ...][tw1 | tw2 | half taper]; else buf | tw1 | red | mbar | acc | tw2
  unsigned char *grp = smem_raw + (PAIR ? (size_t) g * G::GROUP_BYTES : 0);
  unsigned char *shr = smem_raw + 2 * G::GROUP_BYTES;
  float2 *buf = reinterpret_cast<float2 *>(grp);
  float2 *tw1 = reinterpret_cast<float2 *>(PAIR ? shr : smem_raw + G::BUF_BYTES);
  float *red = reinterpret_cast<float *>(grp + G::BUF_BYTES + (PAIR ? 0 : G::TW_BYTES));
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(grp + G::BUF_BYTES + (PAIR ? 0 : G::TW_BYTES) + G::RED_BYTES);
  float *acc = reinterpret_cast<float *>(smem_raw + G::BUF_BYTES + G::TW_BYTES + G::RED_BYTES + 16) + threadIdx.x;   // [slot][T]
  constexpr bool TW2 = PAIR || G::tw2(MULTI);
  float2 *tw2 = reinterpret_cast<float2 *>(PAIR ? shr + G::TW_BYTES : smem_raw + G::BUF_BYTES + G::TW_BYTES + G::RED_BYTES + 16 + G::acc_bytes(MULTI));
  constexpr bool TACC = MULTI && G::TACC;           // the row being summed over the tapers lives in tensor memory
  const float2 *htap = reinterpret_cast<const float2 *>(shr + G::TW_BYTES + G::TW2_BYTES);
  // (barrier numbers as immediates: with the number in a register ptxas reserves all 16 barriers)
  auto gsync = [&]() {
    if constexpr (PAIR) {
      if (g == 0) asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
      else asm volatile("bar.sync 2, %0;" ::"n"(T) : "memory");
    } else {
      __syncthreads();
    }
  };

  // middle-pass twiddles exp(-2 pi i k r / (32 R1)) = roots[16 k r], one copy per CTA
  for (int i = threadIdx.x; i < Big<M>::TW1; i += blockDim.x) tw1[i] = p.roots[(i & 31) * ((i >> 5) + 1) * 16];
  if constexpr (TW2) {
    if (g == 0) {
#pragma unroll
      for (int r = 1; r < 16; r++) tw2[(r - 1) * T + t] = p.roots[t * r];       // t r < 15 M / 32
    }
  }
  if constexpr (PAIR) {
    float2 *ht = reinterpret_cast<float2 *>(shr + G::TW_BYTES + G::TW2_BYTES);
    const float2 *src = reinterpret_cast<const float2 *>(p.tapers);
    for (int i = threadIdx.x; i < M / 2; i += 2 * T) ht[i] = src[i];
  }
  uint32_t tacc = 0;                                // this thread's first word in tensor memory: (lane << 16) | column
  uint32_t *tslot = reinterpret_cast<uint32_t *>(mbar + 1);          // (the second half of the mbarrier's 16 bytes)
  if constexpr (TACC) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t) __cvta_generic_to_shared(tslot)), "n"(G::TACC_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if constexpr (TACC) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w reaches lanes 32 (w mod 4) .. + 31 (32x32b shape); the four warps that share a lane quarter take 40 columns each
    const int w = threadIdx.x >> 5;
    tacc = *tslot + ((uint32_t) ((w & 3) * 32) << 16) + (uint32_t) ((w >> 2) * 40);
  }

  const long long fb = (long long) (blockIdx.x * (PAIR ? 2 : 1) + g) * p.frames_per_group;
  const int nact = (int) ((p.nframes - fb < p.frames_per_group) ? p.nframes - fb : p.frames_per_group);
  const bool sub = p.fused_mean != 0;
  const bool db = p.rows_db != 0;
  const int ntap = MULTI ? p.ntapers : 1;
  // (Carrying only a frame counter and forming these 64-bit values where they are used removes most of the PAIR
  // variant's spills -- 16 -> 4 bytes -- and measured 0.8 % slower.)
  float *row_ptr = p.rows + fb * p.row_stride;                   // (never dereferenced when p.rows is null)
  unsigned char *lev_ptr = p.levels + fb * p.lev_stride;
  long long s0 = (p.first_frame + fb) * (long long) HOP - (N - HOP);   // stream index of the frame's first sample

  // The exchange buffer is idle from the last-pass loads of one frame to the first-pass stores of the next:
  // that is where the NEXT frame's samples land, by one TMA bulk copy (cp.async.bulk + mbarrier) issued as
  // soon as the buffer is free and completed behind the last pass, the split and the row stores.  No
  // staging memory of its own, no global-load instructions for samples, no L1 pollution (the taper stays).
  // A frame qualifies when it lies entirely inside the staged span, 16-byte aligned; the few others
  // (zero history at the stream start, unaligned origins) are read with checked scalar loads.
  auto bulk_ok = [&](long long fs0) {
    const long long r = fs0 - p.origin;
    return fs0 >= 0 && r >= 0 && r + N <= p.count && (r & 3) == 0;
  };
  unsigned phase = 0;
  if (t == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (nact > 0 && bulk_ok(s0)) {
      GLB_CHECK_SRC(p, p.samples + (s0 - p.origin), N);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(mbar, N * 4u);
      tma_load_1d(buf, p.samples + (s0 - p.origin), N * 4u, mbar);
    }
  }
  __syncthreads();                     // (the last CTA-wide barrier: from here on the groups run on their own)

  // PAIR: the two groups of a CTA start together and take the same time per frame; left alone they stay in step,
  // both in their butterflies (FP32 pipe) or both in their exchanges (shared-memory pipe) at once.  A frame is
  // FP-heavy at both ends (first pass; last pass + split) and exchange-heavy in the middle, so the groups are
  // held half a frame apart by a handshake on two named barriers: a group starts a frame only when the OTHER one
  // has passed the first exchange of its current frame (bar.arrive there, bar.sync here: 256 + 256 threads).
  // Both groups run the same number of iterations (idle ones past their last frame) so that the counts match.
  // (Measured at N = 16384, 50 % overlap: free-running 1.58 ms; group 1 started 7 250 ... 8 500 cycles late 1.35 ms, any
  // other delay 1.56 ms -- small offsets decay back into step; this handshake 1.36 ms at every overlap.  Arriving later
  // in the frame -- after pass 1, after the second exchange -- measured 1 % and 7 % slower.)
  // (one-sided -- only group 1 waiting for group 0 -- measured 0.6 % slower)
  auto hs_arrive = [&]() {
    if constexpr (PAIR) {
      if (g == 0) asm volatile("bar.arrive 3, %0;" ::"n"(2 * T) : "memory");
      else asm volatile("bar.arrive 4, %0;" ::"n"(2 * T) : "memory");
    }
  };
  auto hs_wait = [&]() {
    if constexpr (PAIR) {
      if (g == 0) asm volatile("bar.sync 4, %0;" ::"n"(2 * T) : "memory");
      else asm volatile("bar.sync 3, %0;" ::"n"(2 * T) : "memory");
    }
  };
  // (PAIR: the trip count is read from the parameter bank, not carried in a register: the kernel is at its 128)
  for (int it = 0; it < (PAIR ? p.frames_per_group : nact); ++it, s0 += HOP, row_ptr += p.row_stride, lev_ptr += p.lev_stride) {
   if (g == 1 || it > 0) hs_wait();    // group 1: group 0 is past the first exchange of frame `it`; group 0: group 1 of frame `it - 1`
   if (it >= nact) {                   // (PAIR only: an idle iteration keeps the handshake counts equal)
     hs_arrive();
     continue;
   }
   float bs[NBLK];
   for (int j = 0; j < ntap; ++j) {
    const float2 *w2 = reinterpret_cast<const float2 *>(p.tapers + (size_t) j * N) + t;
    float2 v[kBP];
    if (bulk_ok(s0)) {
      mbar_wait(mbar, phase);
      phase ^= 1;
#pragma unroll
      for (int q = 0; q < kBP; q++) v[q] = buf[t + T * q];
    } else {
      // edge frame: zero history before the stream start (fft.c:103-108), nothing past the staged span
#pragma unroll
      for (int q = 0; q < kBP; q++) {
        float y[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const long long s = s0 + 2 * (t + T * q) + e;
          const long long r = s - p.origin;
          y[e] = (s >= 0 && r >= 0 && r < p.count) ? __ldg(p.samples + r) : 0.f;
        }
        v[q] = make_float2(y[0], y[1]);
      }
    }
    if (sub && j == 0) {
      // block means (prepare_audio, fft.c:86-96): block b = registers [b QB, (b + 1) QB).  The summation tree
      // of a block does not depend on its position in the frame, so its mean is the same bits in every
      // frame (and time shard) it appears in; zero history sums to a zero mean.  (Carrying the older blocks'
      // means from frame to frame instead of summing them again saved nothing and cost live registers.)
#pragma unroll
      for (int b = 0; b < NBLK; b++) {
        float a[QB];
#pragma unroll
        for (int q = 0; q < QB; q++) a[q] = v[b * QB + q].x + v[b * QB + q].y;
#pragma unroll
        for (int w = QB / 2; w >= 1; w /= 2) {
#pragma unroll
          for (int q = 0; q < w; q++) a[q] += a[q + w];
        }
        float s = a[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((t & 31) == 0) red[b * NW + (t >> 5)] = s;
      }
    }
    gsync();                   // (A) block sums visible; everyone has taken its samples out of the buffer
    if (sub) {
      if (j == 0) {
#pragma unroll
        for (int b = 0; b < NBLK; b++) bs[b] = big_sum_warps<NW>(red + b * NW) * p.inv_hop_mean;
      }
#pragma unroll
      for (int q = 0; q < kBP; q++) v[q] = sub2(v[q], bc(bs[q / QB]));
    }
    if constexpr (PAIR) {
      // pair i of the taper for i < M / 2; pair i >= M / 2 is pair M - 1 - i with its two weights exchanged
#pragma unroll
      for (int q = 0; q < kBP / 2; q++) v[q] = mul2(v[q], htap[t + T * q]);
#pragma unroll
      for (int q = kBP / 2; q < kBP; q++) {
        const float2 w = htap[M - 1 - t - T * q];
        v[q] = mul2(v[q], make_float2(w.y, w.x));
      }
    } else {
#pragma unroll
      for (int q = 0; q < kBP; q++) v[q] = mul2(v[q], ld_taper(w2 + T * q));
    }
    big_pass0(v);
    big_scatter0<M>(v, t, buf);
    gsync();
    hs_arrive();               // (later points of the frame measured slower: after pass 1 +1 %, after the second exchange +7 %)
    big_load1<M>(v, t, buf);
    big_pass1<M>(v, t, tw1);
    gsync();                   // every thread has read before anyone overwrites
    big_scatter1<M>(v, t, buf);
    gsync();
    BigLast L;
    big_load_last<M>(L, t, p.roots, p.vtab, !TW2);
    big_load2<M>(v, t, buf);
    gsync();                   // (E) the last-pass loads are done: the buffer is free for the next frame
    {
      // what lands next: this frame again for its next taper (from L2), else the next frame
      const bool again = j + 1 < ntap;
      const long long sn = again ? s0 : s0 + HOP;
      if (t == 0 && (again || it + 1 < nact) && bulk_ok(sn)) {
        GLB_CHECK_SRC(p, p.samples + (sn - p.origin), N);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mbar, N * 4u);
        tma_load_1d(buf, p.samples + (sn - p.origin), N * 4u, mbar);
      }
    }
    if constexpr (TW2) big_pass2_tab<M>(v, t, tw2);
    else big_pass2<M>(v, t, L);
    float yv[33];
    yv[32] = 1.f;                      // only thread 0 has a 33rd bin
    auto sink = [&](int slot, float2 a, bool) { yv[slot] = norm2(a); };
    if (t < 32) big_emit<M, true>(v, t, L, sink);      // warp-uniform: only warp 0 pays for thread 0's re-ordering
    else big_emit<M, false>(v, t, L, sink);
    if constexpr (MULTI) {
      // sum over the tapers in this thread's column of the shared-memory row (own entries only: no hazard)
      if constexpr (TACC) {
        if (j > 0) {
          float a[33];
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");     // the previous taper's stores
          tmem_ld16(tacc, a);
          tmem_ld16(tacc + 16, a + 16);
          a[32] = tmem_ld1(tacc + 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int sl = 0; sl < 33; sl++) yv[sl] += a[sl];
        }
        if (j + 1 < ntap) {
          tmem_st16(tacc, yv);
          tmem_st16(tacc + 16, yv + 16);
          tmem_st1(tacc + 32, yv[32]);
          continue;
        }
      } else {
        if (j > 0) {
#pragma unroll
          for (int sl = 0; sl < 33; sl++) yv[sl] += acc[sl * T];
        }
        if (j + 1 < ntap) {
#pragma unroll
          for (int sl = 0; sl < 33; sl++) acc[sl * T] = yv[sl];
          continue;
        }
      }
    }
    // the row leaves the registers: slot 2 rp is bin k, slot 2 rp + 1 is bin M - k, k = t + rp 2T (thread 0:
    // rp 2T, then T + (rp - 8) 2T); consecutive threads store consecutive bins
    const int kh = big_khi<M>(t) - 16 * T;
    if constexpr (LEV) {
      unsigned char *pa = lev_ptr + (M - t), *pb = lev_ptr + t, *pah = lev_ptr + (M - kh), *pbh = lev_ptr + kh;
#pragma unroll
      for (int rp = 0; rp < 16; rp++) {
        *((rp < 8 ? pa : pah) - rp * 2 * T) = map_level(yv[2 * rp], p.lm);
        *((rp < 8 ? pb : pbh) + rp * 2 * T) = map_level(yv[2 * rp + 1], p.lm);
      }
      if (t == 0) lev_ptr[M - M / 2] = map_level(yv[32], p.lm);
      if (p.rows == nullptr) continue;
    }
    if (db) {
#pragma unroll
      for (int s = 0; s < 33; s++) yv[s] = 10.f * log10f(yv[s]);
    }
    float *ra = row_ptr + t, *rb = row_ptr + (M - t), *rah = row_ptr + kh, *rbh = row_ptr + (M - kh);
    GLB_CHECK_ROW(p, row_ptr);
    GLB_CHECK_ROW(p, row_ptr + M);
    GLB_CHECK(kh + 30 * T <= M && M - kh - 30 * T >= 0);
#pragma unroll
    for (int rp = 0; rp < 16; rp++) {
      st_row((rp < 8 ? ra : rah) + rp * 2 * T, yv[2 * rp]);
      st_row((rp < 8 ? rb : rbh) - rp * 2 * T, yv[2 * rp + 1]);
    }
    if (t == 0) st_row(row_ptr + M / 2, yv[32]);
   }
  }
  if (g == 0 && (PAIR ? p.frames_per_group : nact) > 0) hs_wait();   // (PAIR: group 1's last arrival)
  if constexpr (TACC) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tslot), "n"(G::TACC_COLS) : "memory");
  }
}

template <int M, int NBLK, bool MULTI, bool LEV, bool PAIR = false>
static int launch_big(const KParams &kp, int groups_hint, cudaStream_t st) {
  using G = BigGeo<M>;
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  auto kern = gram_big_kernel<M, NBLK, MULTI, LEV, PAIR>;
  constexpr size_t smem = PAIR ? G::PAIR_SMEM : G::smem(MULTI);
  constexpr int GPC = PAIR ? 2 : 1;               // frame groups per CTA
  static thread_local int occ_cache[64];
  int &occ = occ_cache[dev & 63];
  if (occ == 0) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::T * GPC, smem));
    if (occ < 1) occ = 1;
  }
  // resident grid: every frame group walks a contiguous run of frames (its older hop blocks stay in L2)
  long long groups = groups_hint > 0 ? groups_hint : (long long) sms * occ * GPC;
  if (groups > kp.nframes) groups = kp.nframes;
  if (groups < 1) groups = 1;
  KParams k = kp;
  k.frames_per_group = (int) ((kp.nframes + groups - 1) / groups);
  const long long ngroups = (kp.nframes + k.frames_per_group - 1) / k.frames_per_group;
  const long long ctas = (ngroups + GPC - 1) / GPC;
  kern<<<(unsigned) ctas, G::T * GPC, smem, st>>>(k);
  CU(cudaGetLastError());
  g_launches++;
  g_last_family = 5;
  return GLB_OK;
}

template <int M, bool MULTI, bool LEV>
static int launch_big_ml(const KParams &kp, int groups_hint, cudaStream_t st) {
  const int n = 2 * M;
  if constexpr (!MULTI && !LEV && BigGeo<M>::PAIR_FITS) {
    if (kp.taper_sym && g_big_pair) {
      if (kp.hop == n) return launch_big<M, 1, false, false, true>(kp, groups_hint, st);
      if (kp.hop == n / 2) return launch_big<M, 2, false, false, true>(kp, groups_hint, st);
      if (kp.hop == n / 4) return launch_big<M, 4, false, false, true>(kp, groups_hint, st);
    }
  }
  if (kp.hop == n) return launch_big<M, 1, MULTI, LEV>(kp, groups_hint, st);
  if (kp.hop == n / 2) return launch_big<M, 2, MULTI, LEV>(kp, groups_hint, st);
  if (kp.hop == n / 4) return launch_big<M, 4, MULTI, LEV>(kp, groups_hint, st);
  return -1;
}

template <int M, bool WITH_LEV>
static int launch_big_m(const KParams &kp, bool multi, int groups_hint, cudaStream_t st) {
  if (kp.levels != nullptr) {
    if constexpr (WITH_LEV) return multi ? launch_big_ml<M, true, true>(kp, groups_hint, st) : launch_big_ml<M, false, true>(kp, groups_hint, st);
    else return -1;
  }
  return multi ? launch_big_ml<M, true, false>(kp, groups_hint, st) : launch_big_ml<M, false, false>(kp, groups_hint, st);
}

// -1: this launch is not one the big-frame kernel serves (the caller goes on to the 16-point families).
// small_too: also N = 4096 / 8192 (kernel preference 5: experiments; the 16-point ring kernel is faster there)
int glb_gram_big(int m, const KParams &kp, bool multi, int groups_hint, cudaStream_t st, bool small_too) {
  if ((kp.rows == nullptr && kp.levels == nullptr) || kp.spectrum != nullptr || kp.means != nullptr) return -1;
  if (kp.ra9mb_a > 0.f || kp.limiter != 0 || kp.zero_hist || kp.av_on) return -1;
  // pairs of samples are read as one 64-bit word: even offsets, 8-byte aligned base
  if ((kp.hop & 1) || (kp.origin & 1) || (reinterpret_cast<uintptr_t>(kp.samples) & 7)) return -1;
  switch (m) {
    case 2048: return small_too ? launch_big_m<2048, false>(kp, multi, groups_hint, st) : -1;
    case 4096: return small_too ? launch_big_m<4096, false>(kp, multi, groups_hint, st) : -1;
    case 8192: return launch_big_m<8192, true>(kp, multi, groups_hint, st);
    case 16384: return launch_big_m<16384, true>(kp, multi, groups_hint, st);
    default: return -1;
  }
}
