// gram_kernels.cu -- sm_100a kernels of the spectrogram hot path + the C-ABI shim.
//
// Kernels (all hand-written, no library FFT):
//   gram_kernel<M>        overlapped-frame gather -> [block-mean removal] -> [RA9MB] ->
//                         taper multiply -> [limiter] -> real FFT (N = 2M, in registers +
//                         swizzled shared memory) -> |X|^2 -> [sum over K' tapers with the
//                         1/lambda weights folded into the tapers] -> [10 log10] -> one PSD
//                         row per frame straight to HBM.  Replaces prepare_audio + fft_do +
//                         fft_psd (fft.c:66-226) and the taper loop of mtm_do (mtm.c:189-220).
//   block_means_kernel    mean of every hop block (prepare_audio, fft.c:86-96).
//   avg_kernel            sliding per-bin frame averaging, three normalisations
//                         (update_avg_*, avg.c:108-298).
//   peak_carry_kernel     the carried *peakbin of avg.c:129-133.
//   pcm*_to_float_kernel  WAV sample conversion (wav_fmt.c:105-116).
//
// No tensor cores: the path is an FFT at ~9 flop/B executed, bound by HBM and the
// FP32/shared-memory pipes, not a dense contraction.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <atomic>
#include <vector>
#include <algorithm>

#include "fft_core.cuh"
#include "fft_wpf.cuh"
#include "tables.hpp"
#include "../../include/glb_shim.h"

using namespace glb;

// ------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
static int g_force_generic = 0;     // tests: run the general kernel where the ring kernel would be chosen

extern "C" const char *glb_last_error(void) { return g_err; }
extern "C" void glb_set_error(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg ? msg : ""); }
extern "C" unsigned long long glb_kernel_launches(void) { return g_launches.load(); }
extern "C" void glb_force_generic_kernel(int on) { g_force_generic = on; }

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      snprintf(g_err, sizeof g_err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
               __FILE__, __LINE__);                                                           \
      return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? GLB_ENODEV      \
             : (e_ == cudaErrorMemoryAllocation ? GLB_ENOMEM : GLB_ECUDA);                    \
    }                                                                                         \
  } while (0)

// ------------------------------------------------------------------------- plumbing
extern "C" int glb_device_count(int *count) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof g_err, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    *count = 0;
    return GLB_ENODEV;
  }
  *count = c;
  return GLB_OK;
}
extern "C" int glb_set_device(int dev) { CU(cudaSetDevice(dev)); return GLB_OK; }
extern "C" int glb_sm_count(int dev, int *sms) {
  CU(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return GLB_OK;
}
extern "C" int glb_malloc(void **p, size_t bytes) { CU(cudaMalloc(p, bytes ? bytes : 1)); return GLB_OK; }
extern "C" int glb_free(void *p) { if (p) CU(cudaFree(p)); return GLB_OK; }
extern "C" int glb_memset(void *p, int v, size_t bytes, void *s) {
  CU(cudaMemsetAsync(p, v, bytes, (cudaStream_t) s));
  return GLB_OK;
}
extern "C" int glb_host_alloc(void **p, size_t bytes) { CU(cudaMallocHost(p, bytes ? bytes : 1)); return GLB_OK; }
extern "C" int glb_host_free(void *p) { if (p) CU(cudaFreeHost(p)); return GLB_OK; }
static int copy_(void *d, const void *s, size_t n, cudaMemcpyKind k, void *stream) {
  if (n == 0) return GLB_OK;
  if (stream) CU(cudaMemcpyAsync(d, s, n, k, (cudaStream_t) stream));
  else CU(cudaMemcpy(d, s, n, k));
  return GLB_OK;
}
extern "C" int glb_memcpy_h2d(void *d, const void *s, size_t n, void *st) { return copy_(d, s, n, cudaMemcpyHostToDevice, st); }
extern "C" int glb_memcpy_d2h(void *d, const void *s, size_t n, void *st) { return copy_(d, s, n, cudaMemcpyDeviceToHost, st); }
extern "C" int glb_memcpy_d2d(void *d, const void *s, size_t n, void *st) { return copy_(d, s, n, cudaMemcpyDeviceToDevice, st); }
extern "C" int glb_stream_create(void **s) {
  cudaStream_t st;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *s = st;
  return GLB_OK;
}
extern "C" int glb_stream_destroy(void *s) { if (s) CU(cudaStreamDestroy((cudaStream_t) s)); return GLB_OK; }
extern "C" int glb_stream_sync(void *s) { CU(cudaStreamSynchronize((cudaStream_t) s)); return GLB_OK; }
extern "C" int glb_stream_wait_event(void *s, void *e) { CU(cudaStreamWaitEvent((cudaStream_t) s, (cudaEvent_t) e, 0)); return GLB_OK; }
extern "C" int glb_device_sync(void) { CU(cudaDeviceSynchronize()); return GLB_OK; }
extern "C" int glb_event_create(void **e) { cudaEvent_t ev; CU(cudaEventCreate(&ev)); *e = ev; return GLB_OK; }
extern "C" int glb_event_destroy(void *e) { if (e) CU(cudaEventDestroy((cudaEvent_t) e)); return GLB_OK; }
extern "C" int glb_event_record(void *e, void *s) { CU(cudaEventRecord((cudaEvent_t) e, (cudaStream_t) s)); return GLB_OK; }
extern "C" int glb_event_sync(void *e) { CU(cudaEventSynchronize((cudaEvent_t) e)); return GLB_OK; }
extern "C" int glb_event_elapsed_ms(void *a, void *b, float *ms) {
  CU(cudaEventElapsedTime(ms, (cudaEvent_t) a, (cudaEvent_t) b));
  return GLB_OK;
}

// ------------------------------------------------------------------------- tables
struct GramTables {
  int n;
  float2 *tw;
  float2 *vtab;
  float2 *roots;     // exp(-2 pi i k / M), k < M (warp-per-frame kernel)
};

template <int M> static std::vector<float2> tw_for() { return build_twiddles<M>(); }

static bool host_twiddles(int m, std::vector<float2> &tw) {
  switch (m) {
    case 16: tw = tw_for<16>(); return true;
    case 32: tw = tw_for<32>(); return true;
    case 64: tw = tw_for<64>(); return true;
    case 128: tw = tw_for<128>(); return true;
    case 256: tw = tw_for<256>(); return true;
    case 512: tw = tw_for<512>(); return true;
    case 1024: tw = tw_for<1024>(); return true;
    case 2048: tw = tw_for<2048>(); return true;
    case 4096: tw = tw_for<4096>(); return true;
    case 8192: tw = tw_for<8192>(); return true;
    case 16384: tw = tw_for<16384>(); return true;
    default: return false;
  }
}

extern "C" int glb_fft_supported(int n) {
  return n >= 32 && n <= 32768 && (n & (n - 1)) == 0;
}

extern "C" int glb_tables_create(int n, void **out) {
  if (!glb_fft_supported(n)) {
    snprintf(g_err, sizeof g_err, "FFT size %d unsupported (power of two, 32..32768)", n);
    return GLB_EINVAL;
  }
  std::vector<float2> tw;
  host_twiddles(n / 2, tw);
  std::vector<float2> vt = build_vtab(n / 2);
  GramTables *t = new GramTables();
  t->n = n;
  t->tw = nullptr;
  t->vtab = nullptr;
  t->roots = nullptr;
  std::vector<float2> rt = build_roots(n / 2);
  CU(cudaMalloc(&t->roots, rt.size() * sizeof(float2)));
  CU(cudaMemcpy(t->roots, rt.data(), rt.size() * sizeof(float2), cudaMemcpyHostToDevice));
  CU(cudaMalloc(&t->tw, tw.size() * sizeof(float2)));
  CU(cudaMalloc(&t->vtab, vt.size() * sizeof(float2)));
  CU(cudaMemcpy(t->tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(t->vtab, vt.data(), vt.size() * sizeof(float2), cudaMemcpyHostToDevice));
  *out = t;
  return GLB_OK;
}

extern "C" int glb_tables_destroy(void *tp) {
  GramTables *t = (GramTables *) tp;
  if (!t) return GLB_OK;
  cudaFree(t->tw);
  cudaFree(t->vtab);
  cudaFree(t->roots);
  delete t;
  return GLB_OK;
}

// ------------------------------------------------------------------------- gram kernel
struct KParams {
  const float *samples;
  long long origin, count;
  const float *tapers;
  int ntapers;
  const float *means;         // pre-computed block means (general geometry), or nullptr
  long long means_first_block;
  int fused_mean;             // 1: block means are computed inside the kernel (regular geometry)
  int qs;                     // regular geometry: hop = 2T << qs
  float inv_hop_mean;         // 1 / hop
  int hop, n_ov, cblk;        // cblk = ceil(n_ov / hop)
  float inv_hop;
  float ra9mb_a;
  int limiter;
  float lim_scale;            // taper_scale^0.9
  float spec_scale;           // 1 / (2 taper_scale)
  long long first_frame, nframes;
  int frames_per_group;
  float *rows;
  long long row_stride;
  int rows_db;
  float2 *spectrum;
  const float2 *tw, *vtab, *roots;
};

#ifndef GLB_REG_TARGET
#define GLB_REG_TARGET 80
#endif
#ifndef GLB_STREAM_STORE
#define GLB_STREAM_STORE 1
#endif

template <int M> struct Geo {
  static constexpr int N = 2 * M;
  static constexpr int T = M / kPoints;
  static constexpr int G = (T >= 128) ? 1 : 128 / T;   // frame groups per CTA
  static constexpr int THREADS = G * T;
  static constexpr int NW = (T + 31) / 32;             // warps per group
  // frames are staged in shared memory by TMA bulk copies, one frame ahead, when the
  // staging buffer still leaves room for >= 2 CTAs per SM
  static constexpr bool STAGE = (M <= 4096);           // (used by the multitaper variant only)
  // twiddles kept in registers across frames instead of per-frame table loads
  static constexpr bool RT = (M <= 2048);
  static constexpr size_t BUF_BYTES = (size_t) BufSize<M>::value * sizeof(float2);      // multiple of 16
  static constexpr size_t RED_BYTES = 16 * NW * sizeof(float) + 16 * sizeof(float);     // partial sums + means
  // per frame group: FFT buffer | [staging buffer] | reduction scratch | mbarrier
  static __host__ __device__ constexpr size_t stage_bytes(bool multi) { return (STAGE && multi) ? (size_t) N * sizeof(float) : 0; }
  static __host__ __device__ constexpr size_t group_bytes(bool multi) { return ((BUF_BYTES + stage_bytes(multi) + RED_BYTES + 16 + 15) / 16) * 16; }
  static __host__ __device__ constexpr size_t smem_bytes(bool multi) { return (size_t) G * group_bytes(multi); }
  // CTAs per SM the register allocation is tuned for (~GLB_REG_TARGET registers per thread)
  static constexpr int MINB_ = 65536 / (THREADS * (THREADS >= 512 ? 64 : GLB_REG_TARGET));
  static constexpr int MINB = MINB_ < 1 ? 1 : (MINB_ > 16 ? 16 : MINB_);
};

// Barrier over the frame groups of a CTA.  (Tried and dropped, B200: mapping a 4-warp group onto
// one SM sub-partition with its own named barrier -- no gain over the CTA-wide barrier.)
template <int M> __device__ __forceinline__ void group_sync(int) { __syncthreads(); }

// ---- mbarrier / TMA bulk-copy helpers (1-D cp.async.bulk global -> shared) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float2 ldg2(const float2 *p) { return __ldg(p); }
// The taper table is the one global array every frame re-reads; with ~28 KB of L1 left beside the
// shared-memory carve-out it stays resident only if its lines are the last to go and the row
// stores do not allocate.
#ifndef GLB_L1_POLICY
#define GLB_L1_POLICY 1
#endif
__device__ __forceinline__ float2 ld_taper(const float2 *p) {
#if GLB_L1_POLICY
  float2 r;
  asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
#else
  return __ldg(p);
#endif
}

// frame f can be fetched by one bulk copy: entirely inside the staged samples, no zero
// history, 16-byte aligned
__device__ __forceinline__ bool frame_bulk_ok(const KParams &p, long long f, int n) {
  const long long s0 = f * (long long) p.hop - p.n_ov;
  const long long rel = s0 - p.origin;
  return (s0 >= 0) && (rel >= 0) && (rel + n <= p.count) && ((rel & 3) == 0);
}

// Raw samples of one frame: x[q] = (y[2m], y[2m+1]), m = t + T q; zeros before the stream.
template <int M>
__device__ __forceinline__ void load_raw(float2 (&x)[kPoints], int t, const KParams &p, long long f, const float *stage,
                                         bool staged) {
  constexpr int T = M / kPoints, N = 2 * M;
  if (staged) {
    const float2 *s2 = reinterpret_cast<const float2 *>(stage);
#pragma unroll
    for (int q = 0; q < kPoints; q++) x[q] = s2[t + T * q];
    return;
  }
  const long long s0 = f * (long long) p.hop - p.n_ov;
  const long long rel = s0 - p.origin;
  if ((s0 >= 0) && (rel >= 0) && (rel + N <= p.count) && ((rel & 1) == 0)) {
    const float2 *src = reinterpret_cast<const float2 *>(p.samples + rel);
#pragma unroll
    for (int q = 0; q < kPoints; q++) x[q] = ldg2(src + t + T * q);
    return;
  }
#pragma unroll
  for (int q = 0; q < kPoints; q++) {
    float y[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const long long s = s0 + 2 * (t + T * q) + e;
      const long long r = s - p.origin;
      y[e] = (s >= 0 && r >= 0 && r < p.count) ? __ldg(p.samples + r) : 0.f;
    }
    x[q] = make_float2(y[0], y[1]);
  }
}

// Block-mean removal inside the kernel (prepare_audio, fft.c:86-96) for the regular
// geometry hop = 2T << QS, n_ov a multiple of hop: the frame is NB = 16 >> QS whole hop
// blocks, block b = registers q with (q >> QS) == b.  Every thread sums its share of each
// block, warps reduce by shuffle, the group combines through shared memory.  A block gets
// the same summation tree in every frame it appears in, so its mean is bit-identical
// across frames (and across time shards).  Zero history sums to a zero mean.
template <int M, int QS>
__device__ __forceinline__ void remove_block_means(float2 (&x)[kPoints], int t, float *red, float inv_hop, int g) {
  constexpr int T = M / kPoints, NW = (T + 31) / 32, NB = kPoints >> QS;
  constexpr int W = T < 32 ? T : 32;      // lanes of a warp that belong to this group
  float bs[NB];
#pragma unroll
  for (int b = 0; b < NB; b++) {
    float s = 0.f;
#pragma unroll
    for (int q = b << QS; q < (b + 1) << QS; q++) s += x[q].x + x[q].y;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    bs[b] = s;
  }
  if (NW > 1) {
    const int w = t >> 5;
    if ((t & 31) == 0) {
#pragma unroll
      for (int b = 0; b < NB; b++) red[b * NW + w] = bs[b];
    }
    group_sync<M>(g);
#pragma unroll
    for (int b = 0; b < NB; b++) {
      float s = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NW; w2++) s += red[b * NW + w2];
      bs[b] = s;
    }
  }
#pragma unroll
  for (int q = 0; q < kPoints; q++) {
    const float m = bs[q >> QS] * inv_hop;
    x[q].x -= m;
    x[q].y -= m;
  }
}

template <int M>
__device__ __forceinline__ void remove_block_means_qs(float2 (&x)[kPoints], int t, float *red, const KParams &p, int g) {
  switch (p.qs) {
    case 4: remove_block_means<M, 4>(x, t, red, p.inv_hop_mean, g); break;
    case 3: remove_block_means<M, 3>(x, t, red, p.inv_hop_mean, g); break;
    case 2: remove_block_means<M, 2>(x, t, red, p.inv_hop_mean, g); break;
    case 1: remove_block_means<M, 1>(x, t, red, p.inv_hop_mean, g); break;
    default: remove_block_means<M, 0>(x, t, red, p.inv_hop_mean, g); break;
  }
}

// pre-computed block means, any geometry: block of frame sample i is
// f - cblk + floor((i + cblk*hop - n_ov) / hop); samples before the stream keep 0
template <int M>
__device__ __forceinline__ void remove_table_means(float2 (&x)[kPoints], int t, const KParams &p, long long f) {
  constexpr int T = M / kPoints;
  const long long s0 = f * (long long) p.hop - p.n_ov;
  const float *mu = p.means + (f - p.cblk - p.means_first_block);
  const float off = (float) (p.cblk * p.hop - p.n_ov) + 0.5f;
#pragma unroll
  for (int q = 0; q < kPoints; q++) {
    const int i = 2 * (t + T * q);
    if (s0 + i >= 0) x[q].x -= __ldg(mu + (int) (((float) i + off) * p.inv_hop));
    if (s0 + i + 1 >= 0) x[q].y -= __ldg(mu + (int) (((float) (i + 1) + off) * p.inv_hop));
  }
}

// RA9MB -> taper -> limiter (fft.c:127-156); the common case is the bare multiply
template <int M, bool PLAIN>
__device__ __forceinline__ void apply_taper(float2 (&v)[kPoints], const float2 (&x)[kPoints], int t, const KParams &p,
                                            const float *tap) {
  constexpr int T = M / kPoints;
  const float2 *w2 = reinterpret_cast<const float2 *>(tap);
  if (PLAIN || (p.ra9mb_a <= 0.f && p.limiter == 0)) {
#pragma unroll
    for (int q = 0; q < kPoints; q++) {
      const float2 w = ld_taper(w2 + t + T * q);
      v[q] = mul2(x[q], w);
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < kPoints; q++) {
    const float2 w = ldg2(w2 + t + T * q);
    float y[2] = {x[q].x, x[q].y};
    const float ww[2] = {w.x, w.y};
#pragma unroll
    for (int e = 0; e < 2; e++) {
      if (p.ra9mb_a > 0.f) y[e] = y[e] / (p.ra9mb_a + y[e] * y[e]);
      y[e] *= ww[e];
      if (p.limiter == 1) {
        const float m = p.lim_scale * powf(fabsf(y[e]), 0.1f);
        y[e] = (y[e] > 0.f) ? m : -m;
      }
    }
    v[q] = make_float2(y[0], y[1]);
  }
}

template <int M, int P, bool RT> struct MidPasses {
  static __device__ __forceinline__ void run(float2 (&v)[kPoints], int t, float2 *buf, const float2 *tw, const TwRegs &tr, int g) {
    if constexpr (P < Plan<M>::NP - 1) {
      pass_load<M>(v, t, buf);
      if constexpr (RT) {
        pass_compute_rt<M, P>(v, tr);
        group_sync<M>(g);               // every thread has read before anyone overwrites
        pass_scatter<M, P>(v, t, buf);
      } else {
        group_sync<M>(g);
        pass_store<M, P>(v, t, buf, tw);
      }
      group_sync<M>(g);
      MidPasses<M, P + 1, RT>::run(v, t, buf, tw, tr, g);
    }
  }
};

// MULTI: multitaper (K' tapers per frame, frame staged in shared memory by TMA one frame
// ahead and re-read per taper).  PLAIN: no RA9MB / limiter code in the kernel.
template <int M, bool MULTI, bool PLAIN>
__global__ void __launch_bounds__(Geo<M>::THREADS, Geo<M>::MINB) gram_kernel(const KParams p) {
  using GeoM = Geo<M>;
  constexpr int T = GeoM::T, G = GeoM::G, N = GeoM::N;
  constexpr bool STAGE = GeoM::STAGE && MULTI;
  constexpr bool RT = GeoM::RT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int g = threadIdx.x / T;
  const int t = threadIdx.x % T;
  unsigned char *gbase = smem_raw + (size_t) g * GeoM::group_bytes(MULTI);
  float2 *buf = reinterpret_cast<float2 *>(gbase);
  float *stage = reinterpret_cast<float *>(gbase + GeoM::BUF_BYTES);
  float *red = reinterpret_cast<float *>(gbase + GeoM::BUF_BYTES + GeoM::stage_bytes(MULTI));
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(gbase + GeoM::BUF_BYTES + GeoM::stage_bytes(MULTI) + GeoM::RED_BYTES);
  const long long gid = (long long) blockIdx.x * G + g;
  const long long fb = gid * p.frames_per_group;
  unsigned phase = 0;

  TwRegs tr;
  if constexpr (RT) load_tw_regs<M>(tr, t, p.tw, p.vtab);

  if (STAGE) {
    if (t == 0) {
      mbar_init(mbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0 && fb < p.nframes && frame_bulk_ok(p, p.first_frame + fb, N)) {
      const long long f0 = p.first_frame + fb;
      mbar_expect_tx(mbar, N * 4);
      tma_load_1d(stage, p.samples + (f0 * (long long) p.hop - p.n_ov - p.origin), N * 4, mbar);
    }
  }

  for (int it = 0; it < p.frames_per_group; ++it) {
    const long long fl = fb + it;
    const bool active = fl < p.nframes;
    const long long f = p.first_frame + fl;
    const bool staged = STAGE && active && frame_bulk_ok(p, f, N);
    const bool next_there = (it + 1 < p.frames_per_group) && (fl + 1 < p.nframes);
    const bool next_staged = STAGE && next_there && frame_bulk_ok(p, f + 1, N);
    if (staged) {
      mbar_wait(mbar, phase);
      phase ^= 1;
    }
    if (!STAGE && next_there) {
      // pull the next frame's new hop block towards L2 while this frame is computed
      const long long nb = (f + 1) * (long long) p.hop - p.origin;       // first new sample, buffer index
      for (int i = t * 32; i < p.hop; i += T * 32)
        if (nb + i >= 0 && nb + i < p.count) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.samples + nb + i));
    }
    float acc[17];
    if (MULTI) {
#pragma unroll
      for (int i = 0; i < 17; i++) acc[i] = 0.f;
    }
    const int ntap = MULTI ? p.ntapers : 1;
    for (int j = 0; j < ntap; ++j) {
      float2 v[kPoints];
      {
        float2 x[kPoints];
        // Control flow around the warp shuffles / barriers of the mean removal depends only
        // on j (uniform over the CTA), never on a group's own frame: groups of one warp or
        // CTA can sit on different kinds of frames.
        if (MULTI && STAGE && j > 0) {
          // multitaper: tapers after the first read the cleaned copy kept in the staging buffer
          const float2 *s2 = reinterpret_cast<const float2 *>(stage);
#pragma unroll
          for (int q = 0; q < kPoints; q++) x[q] = s2[t + T * q];
        } else {
          if (active) {
            load_raw<M>(x, t, p, f, stage, staged);
          } else {
#pragma unroll
            for (int q = 0; q < kPoints; q++) x[q] = make_float2(0.f, 0.f);
          }
          if (p.fused_mean) remove_block_means_qs<M>(x, t, red, p, g);
          else if (p.means != nullptr && active) remove_table_means<M>(x, t, p, f);
          if (MULTI && STAGE && ntap > 1) {
            float2 *s2 = reinterpret_cast<float2 *>(stage);
#pragma unroll
            for (int q = 0; q < kPoints; q++) s2[t + T * q] = x[q];     // own elements only: no hazard
          }
        }
        apply_taper<M, PLAIN>(v, x, t, p, p.tapers + (size_t) j * N);
      }
      if constexpr (RT) {
        pass_compute_rt<M, 0>(v, tr);
        group_sync<M>(g);              // (A) the previous transform's last pass has been read by all
        pass_scatter<M, 0>(v, t, buf);
      } else {
        group_sync<M>(g);
        pass_store<M, 0>(v, t, buf, p.tw);
      }
      if (STAGE && next_staged && j == ntap - 1 && t == 0) {
        // past barrier (A) everyone is done reading the staging buffer: fetch the next frame
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mbar, N * 4);
        tma_load_1d(stage, p.samples + ((f + 1) * (long long) p.hop - p.n_ov - p.origin), N * 4, mbar);
      }
      group_sync<M>(g);
      MidPasses<M, 1, RT>::run(v, t, buf, p.tw, tr, g);
      auto sink_multi = [&](int slot, float2 a, bool) { acc[slot] += norm2(a); };
      float *row = (!MULTI && active && p.rows) ? p.rows + fl * p.row_stride : nullptr;
      float2 *sp = (!MULTI && active && p.spectrum) ? p.spectrum + fl * (long long) (M + 1) : nullptr;
      const bool db = p.rows_db != 0;
      const float ss = p.spec_scale;
      auto sink_single = [&](int slot, float2 a, bool cj) {
        const int bin = slot_bin<M>(t, slot);
        if (row) {
          float y = norm2(a);
          if (db) y = 10.f * log10f(y);
          row[bin] = y;
        }
        if (sp) sp[bin] = make_float2(a.x * ss, cj ? -a.y * ss : a.y * ss);
      };
      if constexpr (RT) {
        last_pass_rt<M>(v, t, buf, p.tw, tr);
        if (MULTI) emit_bins_rt<M>(v, t, tr, sink_multi);
        else emit_bins_rt<M>(v, t, tr, sink_single);
      } else {
        last_pass<M>(v, t, buf, p.tw);
        if (MULTI) emit_bins<M>(v, t, p.vtab, sink_multi);
        else emit_bins<M>(v, t, p.vtab, sink_single);
      }
      // no barrier here: (A) of the next transform orders these reads before its stores
    }
    if (MULTI && active && p.rows) {
      float *row = p.rows + fl * p.row_stride;
      const bool db = p.rows_db != 0;
#pragma unroll
      for (int slot = 0; slot < 17; slot++) {
        if (slot < 16 || t == 0) {
          float y = acc[slot];
          if (db) y = 10.f * log10f(y);
          row[slot_bin<M>(t, slot)] = y;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------- ring kernel
// The fast path for the regular geometry (hop = 2T << QS, N - hop a multiple of hop: 0, 50,
// 75, 87.5, 93.75 % overlap; no RA9MB / limiter).  A frame is NB = 16 >> QS whole hop blocks.
// Each frame group keeps the last NB + 1 blocks of its run in a shared-memory ring that is
// filled by TMA bulk copies (cp.async.bulk + mbarrier), always one block ahead of the FFT,
// so every sample crosses HBM -> SM exactly once, whatever the overlap, and the DRAM
// latency hides behind the previous frame's transform.  Block means (sub_mean) are computed
// once per block when it lands and kept beside the ring.
#ifndef GLB_MEAN_AHEAD
#define GLB_MEAN_AHEAD 0  // 1: the next block's mean is formed before the last barrier of the current transform
                          //    (measured 4 % slower: the block has to land 40 % of a frame earlier)
#endif
#ifndef GLB_RING_EXTRA
#define GLB_RING_EXTRA 0  // 1: one spare slot, the next block is requested at the top of a frame;
#endif                    // 0: NB slots, requested after barrier (A) into the oldest block's slot
struct RingLayout {
  int slots;            // NB + GLB_RING_EXTRA
  size_t ring_off, red_off, mu_off, mbar_off, group_bytes;
};

template <int M>
__host__ __device__ inline RingLayout ring_layout(int hop, int nb) {
  RingLayout L;
  L.slots = nb + GLB_RING_EXTRA;
  L.ring_off = Geo<M>::BUF_BYTES;
  L.red_off = L.ring_off + (size_t) L.slots * hop * sizeof(float);
  L.mu_off = L.red_off + (size_t) 18 * Geo<M>::NW * sizeof(float);
  L.mbar_off = ((L.mu_off + 18 * sizeof(float) + 7) / 8) * 8;
  L.group_bytes = ((L.mbar_off + 18 * sizeof(unsigned long long) + 15) / 16) * 16;
  return L;
}

// mean of one ring block: every thread sums its (1 << qs) float2 entries, warps reduce by
// shuffle, the group combines through `red` (one barrier).  The summation tree of a block is
// the same wherever the block sits in a frame: its mean is bit-identical in every frame, group
// and time shard.  Must be called by all threads of the CTA (contains a block barrier).
template <int M>
__device__ __forceinline__ float ring_block_partial(const float *blk, int qs, int t, float *red) {
  constexpr int T = M / kPoints, NW = (T + 31) / 32, W = T < 32 ? T : 32;
  const float2 *b2 = reinterpret_cast<const float2 *>(blk);
  float s = 0.f;
  for (int i = 0; i < (1 << qs); i++) {
    const float2 a = b2[t + T * i];
    s += a.x + a.y;
  }
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (NW > 1 && (t & 31) == 0) red[t >> 5] = s;
  return s;
}
// after a block barrier that follows ring_block_partial()
template <int M>
__device__ __forceinline__ float ring_block_total(float s_warp, const float *red, float inv_hop) {
  constexpr int T = M / kPoints, NW = (T + 31) / 32;
  float s = s_warp;
  if (NW > 1) {
    s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; w++) s += red[w];
  }
  return s * inv_hop;
}
template <int M>
__device__ __forceinline__ float ring_block_mean(const float *blk, int qs, int t, float *red, float inv_hop) {
  constexpr int T = M / kPoints, NW = (T + 31) / 32;
  const float s = ring_block_partial<M>(blk, qs, t, red);
  if (NW > 1) __syncthreads();
  return ring_block_total<M>(s, red, inv_hop);
}

// Top of a frame: the NB blocks of the frame from the ring into registers, block means removed.
// `new_mean`: the newest block has just landed and its mean is not known yet; it is summed from
// the registers just loaded (same order as ring_block_partial, so the mean of a block is the
// same bits wherever it is formed), combined across the warps of the group through `red` with
// one block barrier, and left in mu_new / mu[slot_newest].  The loads of the other blocks are
// in flight across that barrier.
template <int M, int QS>
__device__ __forceinline__ void ring_fetch(float2 (&x)[kPoints], int t, const float *ring, int hop, int slot_oldest,
                                           int slots, float *mu, float &mu_new, bool sub, bool new_mean, float *red,
                                           float inv_hop) {
  constexpr int T = M / kPoints, NB = kPoints >> QS, NW = (T + 31) / 32, W = T < 32 ? T : 32;
  int sidx = slot_oldest;
  int slot_of[NB];
#pragma unroll
  for (int b = 0; b < NB; b++) {
    const float2 *bp = reinterpret_cast<const float2 *>(ring + (size_t) sidx * hop);
    slot_of[b] = sidx;
#pragma unroll
    for (int i = 0; i < (1 << QS); i++) x[(b << QS) + i] = bp[t + T * i];
    sidx = (sidx + 1 == slots) ? 0 : sidx + 1;
  }
  if (!sub) return;
  if (new_mean) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < (1 << QS); i++) {
      const float2 a = x[((NB - 1) << QS) + i];
      s += a.x + a.y;
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (NW > 1) {
      if ((t & 31) == 0) red[t >> 5] = s;
      __syncthreads();
    }
    mu_new = ring_block_total<M>(s, red, inv_hop);
    if (t == 0) mu[slot_of[NB - 1]] = mu_new;
  }
#pragma unroll
  for (int b = 0; b < NB; b++) {
    const float m = (b == NB - 1) ? mu_new : mu[slot_of[b]];
#pragma unroll
    for (int i = 0; i < (1 << QS); i++) x[(b << QS) + i] = sub2(x[(b << QS) + i], bc(m));
  }
}

// One PSD row out of the registers (slot numbering of fft_core.cuh): slot 2 rp is bin k, slot
// 2 rp + 1 is bin M - k, k = t + rp 2T except for the upper four pairs of thread 0; base pointers
// plus immediates.  Thread 0 also owns bin M/2.
__device__ __forceinline__ void st_row(float *p, float y) {
#if GLB_L1_POLICY
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(y) : "memory");
#elif GLB_STREAM_STORE
  __stcs(p, y);
#else
  *p = y;
#endif
}
template <int M>
__device__ __forceinline__ void store_row(float *row, int t, const float (&yv)[17]) {
  constexpr int T = M / kPoints;
  const int kh = khi<M>(t) - 8 * T;
  float *ra = row + t, *rb = row + (M - t), *rah = row + kh, *rbh = row + (M - kh);
#pragma unroll
  for (int rp = 0; rp < 8; rp++) {
    st_row((rp < 4 ? ra : rah) + rp * 2 * T, yv[2 * rp]);
    st_row((rp < 4 ? rb : rbh) - rp * 2 * T, yv[2 * rp + 1]);
  }
  if (t == 0) st_row(row + M / 2, yv[16]);
}

// mid passes of the ring kernel; `hook` runs before the last block barrier ahead of the final pass
template <int M, int P, bool RT> struct RingMidPasses {
  template <class H>
  static __device__ __forceinline__ void run(float2 (&v)[kPoints], int t, float2 *buf, const float2 *tw, const TwRegs &tr, H &&hook) {
    if constexpr (P < Plan<M>::NP - 1) {
      pass_load<M>(v, t, buf);
      if constexpr (RT) {
        pass_compute_rt<M, P>(v, tr);
        __syncthreads();                // every thread has read before anyone overwrites
        pass_scatter<M, P>(v, t, buf);
      } else {
        __syncthreads();
        pass_store<M, P>(v, t, buf, tw);
      }
      if constexpr (P == Plan<M>::NP - 2) hook();
      __syncthreads();
      RingMidPasses<M, P + 1, RT>::run(v, t, buf, tw, tr, hook);
    }
  }
};

template <int M, bool MULTI>
__global__ void __launch_bounds__(Geo<M>::THREADS, Geo<M>::MINB) gram_ring_kernel(const KParams p) {
  using GeoM = Geo<M>;
  constexpr int T = GeoM::T, G = GeoM::G, N = GeoM::N;
  constexpr bool RT = GeoM::RT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int g = threadIdx.x / T;
  const int t = threadIdx.x % T;
  const int hop = p.hop, qs = p.qs, nb = kPoints >> qs;          // nb blocks per frame
  const RingLayout L = ring_layout<M>(hop, nb);
  const int slots = L.slots;
  unsigned char *gbase = smem_raw + (size_t) g * L.group_bytes;
  float2 *buf = reinterpret_cast<float2 *>(gbase);
  float *ring = reinterpret_cast<float *>(gbase + L.ring_off);
  float *red = reinterpret_cast<float *>(gbase + L.red_off);
  float *mu = reinterpret_cast<float *>(gbase + L.mu_off);
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(gbase + L.mbar_off);
  const long long gid = (long long) blockIdx.x * G + g;
  const long long fb = gid * p.frames_per_group;
  const bool group_active = fb < p.nframes;
  const long long f_first = p.first_frame + fb;
  const long long b0 = f_first - (nb - 1);                       // oldest block of the first frame
  const bool sub = p.fused_mean != 0;
  const unsigned blk_bytes = (unsigned) hop * 4u;
  unsigned phase_bits = 0;                                       // one parity bit per ring slot

  TwRegs tr;
  if constexpr (RT) load_tw_regs<M>(tr, t, p.tw, p.vtab);

  if (t == 0) {
    for (int sl = 0; sl < slots; sl++) mbar_init(&mbar[sl], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // prologue: the nb blocks of the first frame (zeros before the stream start, fft.c:103-108)
  for (int lb = 0; lb < nb; lb++) {
    const long long blk = b0 + lb;
    if (group_active) {
      if (blk < 0) {
        float2 *z = reinterpret_cast<float2 *>(ring + (size_t) lb * hop);
        for (int i = 0; i < (1 << qs); i++) z[t + T * i] = make_float2(0.f, 0.f);
      } else if (t == 0) {
        mbar_expect_tx(&mbar[lb], blk_bytes);
        tma_load_1d(ring + (size_t) lb * hop, p.samples + (blk * hop - p.origin), blk_bytes, &mbar[lb]);
      }
    }
  }
  float mu_new = 0.f;
  for (int lb = 0; lb < nb; lb++) {
    const long long blk = b0 + lb;
    if (group_active && blk >= 0) {
      mbar_wait(&mbar[lb], 0);
      phase_bits ^= 1u << lb;
    }
    if (sub) {
      __syncthreads();                                           // zero fill visible to all (uniform: groups differ in blk)
      const float m = ring_block_mean<M>(ring + (size_t) lb * hop, qs, t, red + lb * GeoM::NW, p.inv_hop_mean);
      mu_new = (blk < 0) ? 0.f : m;
      if (t == 0) mu[lb] = mu_new;
    }
  }
  __syncthreads();                                               // zero fill and mu[] visible

  int slot_new = nb - 1;                                         // slot of the newest block of frame `it`
  for (int it = 0; it < p.frames_per_group; ++it) {
    const long long fl = fb + it;
    const bool active = fl < p.nframes;
    const long long f = p.first_frame + fl;
    const bool next_there = (it + 1 < p.frames_per_group) && (fl + 1 < p.nframes);
    const int slot_next = (slot_new + 1 == slots) ? 0 : slot_new + 1;
#if !GLB_MEAN_AHEAD
    if (it > 0 && active) {
      // block f was requested one frame ago
      mbar_wait(&mbar[slot_new], (phase_bits >> slot_new) & 1u);
      phase_bits ^= 1u << slot_new;
    }
#endif
    auto request_next = [&]() {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&mbar[slot_next], blk_bytes);
      tma_load_1d(ring + (size_t) slot_next * hop, p.samples + ((f + 1) * (long long) hop - p.origin), blk_bytes, &mbar[slot_next]);
    };
    // spare slot: slot_next held block f - nb, whose last readers passed a block barrier in frame f - 1
    if (GLB_RING_EXTRA && next_there && t == 0) request_next();
    // slot of block f - nb + 1: next to the spare slot, or (tight ring) the slot block f + 1 will take
    const int slot_oldest = GLB_RING_EXTRA ? ((slot_next + 1 == slots) ? 0 : slot_next + 1) : slot_next;
    float acc[17];
    if (MULTI) {
#pragma unroll
      for (int i = 0; i < 17; i++) acc[i] = 0.f;
    }
    const int ntap = MULTI ? p.ntapers : 1;
    float mu_next = 0.f;
    for (int j = 0; j < ntap; ++j) {
      float2 v[kPoints];
      {
        float2 x[kPoints];
        // (mean-ahead variant: the mean is already in mu_new)
        const bool nm = !GLB_MEAN_AHEAD && it > 0 && j == 0;
        float *redn = red + slot_new * GeoM::NW;
        switch (qs) {
          case 4: ring_fetch<M, 4>(x, t, ring, hop, slot_oldest, slots, mu, mu_new, sub, nm, redn, p.inv_hop_mean); break;
          case 3: ring_fetch<M, 3>(x, t, ring, hop, slot_oldest, slots, mu, mu_new, sub, nm, redn, p.inv_hop_mean); break;
          case 2: ring_fetch<M, 2>(x, t, ring, hop, slot_oldest, slots, mu, mu_new, sub, nm, redn, p.inv_hop_mean); break;
          case 1: ring_fetch<M, 1>(x, t, ring, hop, slot_oldest, slots, mu, mu_new, sub, nm, redn, p.inv_hop_mean); break;
          default: ring_fetch<M, 0>(x, t, ring, hop, slot_oldest, slots, mu, mu_new, sub, nm, redn, p.inv_hop_mean); break;
        }
        apply_taper<M, true>(v, x, t, p, p.tapers + (size_t) j * N);
      }
      // Block f + 1 was requested after barrier (A) of this frame's last taper; before the last
      // block barrier of the transform every thread waits for it and leaves its share of the
      // block sum in `red`, so the mean is there after that barrier: the next frame starts
      // without a barrier of its own and without waiting for DRAM.
      const bool last_tap = (j == ntap - 1);
      float s_warp = 0.f;
      auto land_next = [&]() {
        if (GLB_MEAN_AHEAD && last_tap && next_there) {
          mbar_wait(&mbar[slot_next], (phase_bits >> slot_next) & 1u);
          phase_bits ^= 1u << slot_next;
          if (sub) s_warp = ring_block_partial<M>(ring + (size_t) slot_next * hop, qs, t, red);
        }
      };
      if constexpr (RT) {
        pass_compute_rt<M, 0>(v, tr);
        __syncthreads();               // (A) the previous transform's last pass has been read by all
        pass_scatter<M, 0>(v, t, buf);
      } else {
        __syncthreads();
        pass_store<M, 0>(v, t, buf, p.tw);
      }
      // tight ring: past (A) of the last taper nobody reads the oldest block any more
      if (!GLB_RING_EXTRA && next_there && last_tap && t == 0) request_next();
      if constexpr (Plan<M>::NP == 2) land_next();
      __syncthreads();
      RingMidPasses<M, 1, RT>::run(v, t, buf, p.tw, tr, land_next);
      if (GLB_MEAN_AHEAD && last_tap && next_there && sub) {
        mu_next = ring_block_total<M>(s_warp, red, p.inv_hop_mean);
        if (t == 0) mu[slot_next] = mu_next;
      }
      float *row = p.rows + fl * p.row_stride;
      const bool db = p.rows_db != 0;
      float yv[17];
      yv[16] = 1.f;                     // only thread 0 has a 17th bin
      auto sink_multi = [&](int slot, float2 a, bool) { acc[slot] += norm2(a); };
      auto sink_single = [&](int slot, float2 a, bool) { yv[slot] = norm2(a); };
      const bool w0 = t < 32;           // warp-uniform: only the warp that holds thread 0 pays for its re-ordering
      if constexpr (RT) {
        last_pass_rt<M>(v, t, buf, p.tw, tr);
        if (MULTI) {
          if (w0) emit_bins_rt<M, true>(v, t, tr, sink_multi);
          else emit_bins_rt<M, false>(v, t, tr, sink_multi);
        } else {
          if (w0) emit_bins_rt<M, true>(v, t, tr, sink_single);
          else emit_bins_rt<M, false>(v, t, tr, sink_single);
        }
      } else {
        last_pass<M>(v, t, buf, p.tw);
        if (MULTI) {
          if (w0) emit_bins<M, true>(v, t, p.vtab, sink_multi);
          else emit_bins<M, false>(v, t, p.vtab, sink_multi);
        } else {
          if (w0) emit_bins<M, true>(v, t, p.vtab, sink_single);
          else emit_bins<M, false>(v, t, p.vtab, sink_single);
        }
      }
      if (!MULTI) {
        // the row leaves the registers here: one store per bin, streaming (written once, never
        // re-read by this kernel); the dB conversion is a single uniform branch per frame
        if (db) {
#pragma unroll
          for (int slot = 0; slot < 17; slot++) yv[slot] = 10.f * log10f(yv[slot]);
        }
        if (active) store_row<M>(row, t, yv);
      }
    }
    if (MULTI && active) {
      float *row = p.rows + fl * p.row_stride;
      const bool db = p.rows_db != 0;
#pragma unroll
      for (int slot = 0; slot < 17; slot++) {
        if (slot < 16 || t == 0) {
          float y = acc[slot];
          if (db) y = 10.f * log10f(y);
          row[slot_bin<M>(t, slot)] = y;
        }
      }
    }
    slot_new = slot_next;
    if (GLB_MEAN_AHEAD) mu_new = mu_next;
  }
}

// ------------------------------------------------------------------------- warp-per-frame kernel
// Periodogram fast path for N = 512..4096: a frame never leaves its warp (fft_wpf.cuh), so the
// kernel has no block barrier at all.  Warps walk contiguous runs of INTERIOR frames (no zero
// history, fully inside the staged samples, 8-byte aligned); the few edge frames of a
// recording are launched through the general kernel by the host code below.
#ifndef GLB_WPF_MINB
#define GLB_WPF_MINB 5
#endif
constexpr int kWpfWarps = 2;          // warps per CTA (independent of each other)

// block-mean removal inside one lane group: hop = 2T << QW; block b = registers q with (q >> QW) == b
template <int M, int QW>
__device__ __forceinline__ void wpf_remove_means(float2 (&x)[kWP], float inv_hop) {
  constexpr int T = Wpf<M>::T, NB = kWP >> QW;
#pragma unroll
  for (int b = 0; b < NB; b++) {
    float s = 0.f;
#pragma unroll
    for (int q = b << QW; q < (b + 1) << QW; q++) s += x[q].x + x[q].y;
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float m = s * inv_hop;
#pragma unroll
    for (int q = b << QW; q < (b + 1) << QW; q++) {
      x[q].x -= m;
      x[q].y -= m;
    }
  }
}

template <int M>
__global__ void __launch_bounds__(32 * kWpfWarps, GLB_WPF_MINB) gram_wpf_kernel(const KParams p) {
  constexpr int T = Wpf<M>::T, N = 2 * M, FPW = 32 / T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / T, t = lane % T;
  float2 *tile = reinterpret_cast<float2 *>(smem_raw) + (size_t) (warp * FPW + g) * (Wpf<M>::TILE + 2);
  const long long gid = ((long long) blockIdx.x * kWpfWarps + warp) * FPW + g;
  const long long fb = gid * p.frames_per_group;

  WpfRegs rg;
  wpf_load_regs<M>(rg, t, p.roots, p.vtab);
  const float2 *w2 = reinterpret_cast<const float2 *>(p.tapers) + t;
  const bool db = p.rows_db != 0;

  for (int it = 0; it < p.frames_per_group; ++it) {
    const long long fl = fb + it;
    const bool active = fl < p.nframes;
    // inactive tail iterations recompute the group's last frame without storing (keeps the
    // warp converged for the shuffles and __syncwarp below)
    const long long fe = active ? fl : (p.nframes - 1);
    const long long f = p.first_frame + fe;
    const float2 *src = reinterpret_cast<const float2 *>(p.samples + (f * (long long) p.hop - p.n_ov - p.origin)) + t;
    float2 v[kWP];
#pragma unroll
    for (int q = 0; q < kWP; q++) v[q] = ldg2(src + T * q);
    if (fl + 1 < p.nframes && it + 1 < p.frames_per_group) {
      // pull the next frame's new hop block towards L2 behind this frame's arithmetic
      const float *nx = p.samples + ((f + 1) * (long long) p.hop - p.origin);
      for (int i = t * 32; i < p.hop; i += T * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + i));
    }
    if (p.fused_mean) {
      switch (p.qs) {
        case 6: wpf_remove_means<M, 6>(v, p.inv_hop_mean); break;
        case 5: wpf_remove_means<M, 5>(v, p.inv_hop_mean); break;
        case 4: wpf_remove_means<M, 4>(v, p.inv_hop_mean); break;
        case 3: wpf_remove_means<M, 3>(v, p.inv_hop_mean); break;
        default: wpf_remove_means<M, 2>(v, p.inv_hop_mean); break;
      }
    }
#pragma unroll
    for (int q = 0; q < kWP; q++) {
      const float2 w = ldg2(w2 + T * q);
      v[q].x *= w.x;
      v[q].y *= w.y;
    }
    wpf_pass_a<M>(v);
    __syncwarp();                      // the previous frame's gathers are done
    wpf_scatter<M>(v, t, tile);
    __syncwarp();
    wpf_gather<M>(v, t, tile);
    wpf_pass_b<M>(v, t, rg);
    float *row = p.rows + fe * p.row_stride;
    wpf_emit<M>(v, t, rg, [&](int, int bin, float2 a, bool) {
      float y = norm2(a);
      if (db) y = 10.f * log10f(y);
      if (active) row[bin] = y;
    });
  }
}

template <int M>
static int launch_wpf(const KParams &kp, cudaStream_t st) {
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  constexpr int FPW = 32 / Wpf<M>::T;
  const size_t smem = (size_t) kWpfWarps * FPW * (Wpf<M>::TILE + 2) * sizeof(float2);
  static thread_local int occ_cache[64];
  int &occ = occ_cache[dev & 63];
  if (occ == 0) {
    CU(cudaFuncSetAttribute(gram_wpf_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gram_wpf_kernel<M>, 32 * kWpfWarps, smem));
    if (occ < 1) occ = 1;
  }
  long long groups = (long long) sms * occ * kWpfWarps * FPW;
  if (groups > kp.nframes) groups = kp.nframes;
  if (groups < 1) groups = 1;
  KParams k = kp;
  k.frames_per_group = (int) ((kp.nframes + groups - 1) / groups);
  const long long used = (kp.nframes + k.frames_per_group - 1) / k.frames_per_group;
  const int ctas = (int) ((used + kWpfWarps * FPW - 1) / (kWpfWarps * FPW));
  gram_wpf_kernel<M><<<ctas, 32 * kWpfWarps, smem, st>>>(k);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

// which kernel family serves a launch (tests / experiments can pin one)
static int g_kernel_pref = 0;        // 0 auto, 1 general, 2 ring, 3 warp-per-frame
extern "C" void glb_set_kernel_preference(int pref) { g_kernel_pref = pref; }

template <int M>
static int launch_gram_m(const KParams &kp, bool multi, int groups_hint, cudaStream_t st, int allow) {
  using GeoM = Geo<M>;
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const bool plain = multi || (kp.ra9mb_a <= 0.f && kp.limiter == 0);
  // ---- fastest path (N <= 4096 periodograms): warp-per-frame kernel on the interior frames
  if constexpr (M >= 256 && M <= 2048) {
    constexpr int unitw = 2 * Wpf<M>::T;
    int qw = -1;
    for (int s2 = 2; s2 <= 6; s2++)
      if (kp.hop == (unitw << s2)) qw = s2;
    const bool mean_ok = !kp.fused_mean || (qw >= 0 && (kp.n_ov % kp.hop) == 0);
    const bool pref_ok = (allow & 4) != 0;
    if (pref_ok && !multi && plain && kp.rows != nullptr && kp.spectrum == nullptr && kp.means == nullptr && mean_ok &&
        (kp.hop % 2) == 0 && ((kp.n_ov + kp.origin) % 2) == 0 && ((reinterpret_cast<uintptr_t>(kp.samples) & 7) == 0)) {
      // interior frames: f*hop - n_ov >= max(origin, 0) and the frame ends inside the staged samples
      const long long lo_s = kp.origin > 0 ? kp.origin : 0;
      long long f_lo = (lo_s + kp.n_ov + kp.hop - 1) / kp.hop;
      long long f_hi = (kp.origin + kp.count - kp.hop) / kp.hop + 1;       // exclusive: (f+1)*hop <= origin+count
      if (f_lo < kp.first_frame) f_lo = kp.first_frame;
      if (f_hi > kp.first_frame + kp.nframes) f_hi = kp.first_frame + kp.nframes;
      if (f_hi > f_lo) {
        KParams kw = kp;
        kw.qs = qw;
        kw.first_frame = f_lo;
        kw.nframes = f_hi - f_lo;
        kw.rows = kp.rows + (f_lo - kp.first_frame) * kp.row_stride;
        int rc = launch_wpf<M>(kw, st);
        if (rc != GLB_OK) return rc;
        // edge frames before / after the interior run go through the kernels below
        KParams ke = kp;
        if (f_lo > kp.first_frame) {
          ke.nframes = f_lo - kp.first_frame;
          rc = launch_gram_m<M>(ke, multi, groups_hint, st, 1);
          if (rc != GLB_OK) return rc;
        }
        if (f_hi < kp.first_frame + kp.nframes) {
          ke = kp;
          ke.first_frame = f_hi;
          ke.nframes = kp.first_frame + kp.nframes - f_hi;
          ke.rows = kp.rows + (f_hi - kp.first_frame) * kp.row_stride;
          rc = launch_gram_m<M>(ke, multi, groups_hint, st, 1);
          if (rc != GLB_OK) return rc;
        }
        return GLB_OK;
      }
    }
  }
  // ---- fast path: regular geometry, rows only, 16-byte aligned blocks -> TMA ring kernel
  {
    const int unit = 2 * GeoM::T;
    int qs = -1;
    for (int s2 = 0; s2 <= 4; s2++)
      if (kp.hop == (unit << s2)) qs = s2;
    const bool regular = qs >= 0 && (kp.n_ov % kp.hop) == 0 && (kp.hop % 4) == 0 && (kp.origin % 4) == 0 &&
                         ((reinterpret_cast<uintptr_t>(kp.samples) & 15) == 0);
    if (regular && plain && kp.rows != nullptr && kp.spectrum == nullptr && kp.means == nullptr && (allow & 2) != 0) {
      const int nb = kPoints >> qs;
      const RingLayout L = ring_layout<M>(kp.hop, nb);
      const size_t smem = (size_t) GeoM::G * L.group_bytes;
      if (smem <= 227 * 1024) {
        auto rk = multi ? gram_ring_kernel<M, true> : gram_ring_kernel<M, false>;
        static thread_local int occ_ring[2][5][64];
        int &occ = occ_ring[multi ? 1 : 0][qs][dev & 63];
        if (occ == 0) {
          // opt in to the device maximum once: the ring size (hence the launch's smem) varies with the overlap
          CU(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rk, GeoM::THREADS, smem));
          if (occ < 1) occ = -1;
        }
        // big frames: the ring must not cost more residency than it saves in traffic
        int occ_generic_bound = (int) ((227 * 1024) / GeoM::smem_bytes(false));
        const bool worth = occ >= 2 || (occ >= 1 && occ_generic_bound <= 1) || GeoM::THREADS >= 1024 || g_kernel_pref == 2;
        if (occ >= 1 && worth) {
          long long groups = groups_hint > 0 ? groups_hint : (long long) sms * occ * GeoM::G;
          if (groups > kp.nframes) groups = kp.nframes;
          if (groups < 1) groups = 1;
          KParams k = kp;
          k.qs = qs;
          k.frames_per_group = (int) ((kp.nframes + groups - 1) / groups);
          long long used = (kp.nframes + k.frames_per_group - 1) / k.frames_per_group;
          int ctas = (int) ((used + GeoM::G - 1) / GeoM::G);
          rk<<<ctas, GeoM::THREADS, smem, st>>>(k);
          CU(cudaGetLastError());
          g_launches++;
          return GLB_OK;
        }
      }
    }
  }
  const int variant = multi ? 2 : (plain ? 1 : 0);
  auto kern = multi ? gram_kernel<M, true, true> : (plain ? gram_kernel<M, false, true> : gram_kernel<M, false, false>);
  // the staging buffer is only carved out for the multitaper variant
  const size_t smem = GeoM::smem_bytes(multi);
  // per (device, variant): opt in to the dynamic shared memory once, cache the occupancy
  static thread_local int occ_cache[3][64];
  int &occ = occ_cache[variant][dev & 63];
  if (occ == 0) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GeoM::THREADS, smem));
    if (occ < 1) occ = 1;
  }
  // resident grid: every CTA slot of the chip holds G frame groups, each walking a
  // contiguous run of frames (keeps the overlapped samples of consecutive frames in L1/L2)
  long long groups = groups_hint > 0 ? groups_hint : (long long) sms * occ * GeoM::G;
  if (groups > kp.nframes) groups = kp.nframes;
  if (groups < 1) groups = 1;
  KParams k = kp;
  k.frames_per_group = (int) ((kp.nframes + groups - 1) / groups);
  long long used = (kp.nframes + k.frames_per_group - 1) / k.frames_per_group;
  int ctas = (int) ((used + GeoM::G - 1) / GeoM::G);
  // regular geometry for the fused block means: hop = 2T << qs, n_ov a multiple of hop
  k.fused_mean = 0;
  k.qs = 0;
  if (kp.fused_mean) {
    const int unit = 2 * GeoM::T;
    int qs = -1;
    for (int s2 = 0; s2 <= 4; s2++)
      if (kp.hop == (unit << s2)) qs = s2;
    if (qs >= 0 && (k.n_ov % kp.hop) == 0) {
      k.fused_mean = 1;
      k.qs = qs;
    } else {
      glb_set_error("glb_launch_gram: fused block means need hop = (N/16) << s, s = 0..4");
      return GLB_EINVAL;
    }
  }
  kern<<<ctas, GeoM::THREADS, smem, st>>>(k);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

extern "C" int glb_gram_fused_mean_ok(int n, int hop) {
  if (!glb_fft_supported(n) || hop < 1 || hop > n) return 0;
  const int unit = n / 16;            // 2T
  if ((n - hop) % hop != 0) return 0;
  for (int s = 0; s <= 4; s++)
    if (hop == (unit << s)) return 1;
  return 0;
}

extern "C" int glb_launch_gram(const glb_gram_args *a, void *stream) {
  if (!a || !a->tables || !glb_fft_supported(a->n) || a->hop < 1 || a->hop > a->n || a->ntapers < 1) {
    glb_set_error("glb_launch_gram: invalid arguments");
    return GLB_EINVAL;
  }
  if (a->nframes <= 0) return GLB_OK;
  const GramTables *tb = (const GramTables *) a->tables;
  if (tb->n != a->n) {
    glb_set_error("glb_launch_gram: tables were built for another FFT size");
    return GLB_EINVAL;
  }
  KParams k;
  memset(&k, 0, sizeof k);
  k.samples = a->samples;
  k.origin = a->origin;
  k.count = a->count;
  k.tapers = a->tapers;
  k.ntapers = a->ntapers;
  k.means = a->block_means;
  k.means_first_block = a->means_first_block;
  k.fused_mean = a->fused_mean;
  k.hop = a->hop;
  k.n_ov = a->n - a->hop;
  k.cblk = (k.n_ov + a->hop - 1) / a->hop;
  k.inv_hop = (float) (1.0 / (double) a->hop);
  k.inv_hop_mean = (float) (1.0 / (double) a->hop);
  k.ra9mb_a = a->ra9mb_a;
  k.limiter = a->limiter;
  k.lim_scale = (float) pow((double) a->taper_scale, 0.9);
  k.spec_scale = (float) (1.0 / (2.0 * (double) a->taper_scale));
  k.first_frame = a->first_frame;
  k.nframes = a->nframes;
  k.rows = a->rows;
  k.row_stride = a->row_stride;
  k.rows_db = a->rows_db;
  k.spectrum = (float2 *) a->spectrum;
  k.tw = tb->tw;
  k.vtab = tb->vtab;
  k.roots = tb->roots;
  const bool multi = a->ntapers > 1;
  if (multi && a->spectrum) {
    glb_set_error("glb_launch_gram: spectrum output is only defined for one taper");
    return GLB_EINVAL;
  }
  if (a->fused_mean && a->block_means) {
    glb_set_error("glb_launch_gram: fused_mean and block_means are exclusive");
    return GLB_EINVAL;
  }
  cudaStream_t st = (cudaStream_t) stream;
  // kernel families this launch may use: 1 general, 2 TMA ring, 4 warp-per-frame
  // (automatic = ring then general: measured on B200 the ring kernel is the fastest family at
  // N = 4096, 0.726 ms vs 0.746 ms for warp-per-frame, whose 67 KB of straight-line code per
  // frame stalls on instruction fetch; warp-per-frame stays selectable for experiments)
  int allow = 3;
  if (g_force_generic || g_kernel_pref == 1) allow = 1;
  else if (g_kernel_pref == 3) allow = 7;
  switch (a->n / 2) {
    case 16: return launch_gram_m<16>(k, multi, a->groups_hint, st, allow);
    case 32: return launch_gram_m<32>(k, multi, a->groups_hint, st, allow);
    case 64: return launch_gram_m<64>(k, multi, a->groups_hint, st, allow);
    case 128: return launch_gram_m<128>(k, multi, a->groups_hint, st, allow);
    case 256: return launch_gram_m<256>(k, multi, a->groups_hint, st, allow);
    case 512: return launch_gram_m<512>(k, multi, a->groups_hint, st, allow);
    case 1024: return launch_gram_m<1024>(k, multi, a->groups_hint, st, allow);
    case 2048: return launch_gram_m<2048>(k, multi, a->groups_hint, st, allow);
    case 4096: return launch_gram_m<4096>(k, multi, a->groups_hint, st, allow);
    case 8192: return launch_gram_m<8192>(k, multi, a->groups_hint, st, allow);
    case 16384: return launch_gram_m<16384>(k, multi, a->groups_hint, st, allow);
  }
  return GLB_EINVAL;
}

// ------------------------------------------------------------------------- block means
// One warp per hop block; lanes stride the block, float partial sums, shuffle tree.
__global__ void block_means_kernel(const float *__restrict__ samples, long long origin, long long count,
                                   int hop, long long first_block, long long nblocks, float *__restrict__ means) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
  for (long long b = warp; b < nblocks; b += nwarps) {
    const long long base = (first_block + b) * hop - origin;
    float s = 0.f;
    for (int i = lane; i < hop; i += 32) {
      const long long r = base + i;
      s += (r >= 0 && r < count) ? __ldg(samples + r) : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) means[b] = s / (float) hop;
  }
}

extern "C" int glb_launch_block_means(const float *samples, long long origin, long long count, int hop,
                                      long long first_block, long long nblocks, float *means, void *stream) {
  if (nblocks <= 0) return GLB_OK;
  if (hop < 1) { glb_set_error("block_means: hop < 1"); return GLB_EINVAL; }
  const int threads = 256;
  long long ctas = (nblocks * 32 + threads - 1) / threads;
  if (ctas > 148 * 16) ctas = 148 * 16;
  block_means_kernel<<<(int) ctas, threads, 0, (cudaStream_t) stream>>>(samples, origin, count, hop, first_block, nblocks, means);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

// ------------------------------------------------------------------------- PCM ingest
__global__ void pcm16_to_float_kernel(const short *__restrict__ in, float *__restrict__ out, long long n) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (float) in[i] / 32768.f;
}
__global__ void pcm8_to_float_kernel(const unsigned char *__restrict__ in, float *__restrict__ out, long long n) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = ((float) in[i] - 128.f) / 128.f;
}
extern "C" int glb_launch_pcm16_to_float(const short *pcm, float *out, long long count, void *stream) {
  if (count <= 0) return GLB_OK;
  long long ctas = std::min<long long>((count + 255) / 256, 148 * 32);
  pcm16_to_float_kernel<<<(int) ctas, 256, 0, (cudaStream_t) stream>>>(pcm, out, count);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}
extern "C" int glb_launch_pcm8_to_float(const unsigned char *pcm, float *out, long long count, void *stream) {
  if (count <= 0) return GLB_OK;
  long long ctas = std::min<long long>((count + 255) / 256, 148 * 32);
  pcm8_to_float_kernel<<<(int) ctas, 256, 0, (cudaStream_t) stream>>>(pcm, out, count);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

// ------------------------------------------------------------------------- averaging
// One CTA walks a chunk of consecutive frames; threads stride the band
// [minbin, maxbin).  Per bin the window sum cum = sum of the last min(f+1, depth) PSD
// values is held in double in shared memory (the reference keeps double cum[],
// avg.c:119,123) and slid frame to frame; the chunk's first frame sums its window
// directly from the rows, so chunks are independent.
struct AvgRed {
  double mx; int arg; double sum; double mn;
};

template <typename OutT>
__global__ void __launch_bounds__(256) avg_kernel(const glb_avg_args a, int chunk) {
  extern __shared__ double cum[];                    // band doubles
  __shared__ double s_mx[8], s_sum[8], s_mn[8], s_var[8];
  __shared__ int s_arg[8], s_cnt[8];
  __shared__ double b_mx, b_sum, b_mn, b_var;
  __shared__ int b_arg, b_cnt;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int band = a.maxbin - a.minbin;
  const long long c0 = (long long) blockIdx.x * chunk;
  const long long c1 = (c0 + chunk < a.nframes) ? c0 + chunk : a.nframes;
  OutT *out_base = (OutT *) a.avg_rows;
  auto psd_row = [&](long long g2) -> const float * {
    const long long r = a.psd_ring_rows > 0 ? g2 % a.psd_ring_rows : g2 - a.psd_first_frame;
    return a.psd + r * a.psd_stride;
  };
  int carried = a.peakbin_init;
  bool carry_known = (blockIdx.x == 0);

  for (long long fl = c0; fl < c1; ++fl) {
    const long long f = a.first_frame + fl;          // frames since alloc_avg
    const float *row = psd_row(f);
    const long long eff = (f + 1 < a.depth) ? f + 1 : a.depth;   // effdepth after this update
    double mx = -1.0, sum = 0.0, mn = 1.0;
    int arg = -1;
    for (int i = tid; i < band; i += 256) {
      const int b = a.minbin + i;
      double c;
      if (fl == c0) {
        c = 0.0;
        for (long long g2 = f - eff + 1; g2 <= f; ++g2) c += (double) psd_row(g2)[b];
      } else {
        c = cum[i] + (double) row[b];
        if (f - a.depth >= 0) c -= (double) psd_row(f - a.depth)[b];
      }
      cum[i] = c;
      if (arg < 0 || c > mx) { mx = c; arg = b; }     // first maximum within this thread's bins
      sum += c;
      if (c < mn) mn = c;
    }
    // block reduction: max with the lowest bin on ties, sum, min
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
      if (oarg >= 0 && (arg < 0 || omx > mx || (omx == mx && oarg < arg))) { mx = omx; arg = oarg; }
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const double omn = __shfl_xor_sync(0xffffffffu, mn, o);
      if (omn < mn) mn = omn;
    }
    if (lane == 0) { s_mx[wid] = mx; s_arg[wid] = arg; s_sum[wid] = sum; s_mn[wid] = mn; }
    __syncthreads();
    if (tid == 0) {
      double m = s_mx[0], sm = s_sum[0], mi = s_mn[0];
      int ar = s_arg[0];
      for (int w = 1; w < 8; w++) {
        if (s_arg[w] >= 0 && (ar < 0 || s_mx[w] > m || (s_mx[w] == m && s_arg[w] < ar))) { m = s_mx[w]; ar = s_arg[w]; }
        sm += s_sum[w];
        if (s_mn[w] < mi) mi = s_mn[w];
      }
      // `double max = psd[minbin]; if (cum > max) { max = cum; *peakbin = index; }` avg.c:111,129-133
      const double m0 = (double) row[a.minbin];
      if (ar >= 0 && m > m0) { b_mx = m; b_arg = ar; } else { b_mx = m0; b_arg = -1; }
      b_sum = sm;
      b_mn = mi;
    }
    __syncthreads();
    const double vmax = b_mx, vsum = b_sum, vmin = b_mn;
    const int cand = b_arg;
    if (cand >= 0) { carried = cand; carry_known = true; }
    const int pk = carried;
    double avgspec = 0.0, retv = 0.0;
    if (a.mode == 2) {
      retv = (vsum - vmax) / ((double) (band - 1) * (double) (eff + 1));
    } else {
      avgspec = (vsum - vmax) / (double) (band - 1);
      retv = vmax / avgspec;
    }
    // output row
    OutT *orow = out_base + fl * a.out_stride;
    double var = 0.0;
    int cnt = 0;
    for (int b = tid; b < a.nbins; b += 256) {
      double y = 1e-15;
      if (b >= a.minbin && b < a.maxbin) {
        const double c = cum[b - a.minbin];
        if (a.mode == 2) {
          y = c / (double) (eff + 1);
        } else if (a.mode == 3) {
          y = a.max0 ? (c - vmin) / (vmax - vmin) : c / avgspec;
        } else {
          if (c - avgspec > 0) {
            y = a.max0 ? (c - avgspec) / (vmax - avgspec) : c / avgspec;
            if (b != pk) { const double r = c / avgspec; var += r * r; cnt++; }
          } else {
            y = 1e-15;
          }
        }
      }
      if (sizeof(OutT) == 4 && a.rows_db) y = 10.0 * log10(y);
      orow[b] = (OutT) y;
    }
    if (a.mode == 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        var += __shfl_xor_sync(0xffffffffu, var, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      }
      if (lane == 0) { s_var[wid] = var; s_cnt[wid] = cnt; }
      __syncthreads();
      if (tid == 0) {
        double v = 0; int c = 0;
        for (int w = 0; w < 8; w++) { v += s_var[w]; c += s_cnt[w]; }
        b_var = v; b_cnt = c;
      }
      __syncthreads();
    }
    if (tid == 0) {
      if (a.ret) a.ret[fl] = retv;
      if (a.peak_cand) a.peak_cand[fl] = cand;
      if (a.variance) a.variance[fl] = (a.mode == 1) ? b_var / (double) b_cnt : 0.0;
      if (a.mode == 1 && !carry_known && a.unresolved) atomicAdd(a.unresolved, 1);
    }
    __syncthreads();
  }
}

// Direct form for small depths: one warp per frame, no block barrier, no state carried
// between frames.  The window sum of every band bin is re-read from the PSD rows (depth loads
// per bin, L2-resident: consecutive frames share depth-1 of them); pass 1 reduces max / first
// argmax / sum / min over the band with shuffles, pass 2 recomputes the sums and writes the
// normalised row.  Frames are fully independent, so the grid is sized by the frame count.
template <typename OutT>
__global__ void __launch_bounds__(256) avg_direct_kernel(const glb_avg_args a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
  const int band = a.maxbin - a.minbin;
  OutT *out_base = (OutT *) a.avg_rows;
  auto psd_row = [&](long long g2) -> const float * {
    const long long r = a.psd_ring_rows > 0 ? g2 % a.psd_ring_rows : g2 - a.psd_first_frame;
    return a.psd + r * a.psd_stride;
  };
  for (long long fl = warp; fl < a.nframes; fl += nwarps) {
    const long long f = a.first_frame + fl;
    const long long eff = (f + 1 < a.depth) ? f + 1 : a.depth;
    const long long g0 = f - eff + 1;
    auto window_sum = [&](int b) {
      double c = 0.0;
      for (long long g2 = g0; g2 <= f; ++g2) c += (double) psd_row(g2)[b];
      return c;
    };
    double mx = -1.0, sum = 0.0, mn = 1.0;
    int arg = -1;
    for (int i = lane; i < band; i += 32) {
      const int b = a.minbin + i;
      const double c = window_sum(b);
      if (arg < 0 || c > mx) { mx = c; arg = b; }
      sum += c;
      if (c < mn) mn = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
      if (oarg >= 0 && (arg < 0 || omx > mx || (omx == mx && oarg < arg))) { mx = omx; arg = oarg; }
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const double omn = __shfl_xor_sync(0xffffffffu, mn, o);
      if (omn < mn) mn = omn;
    }
    const double m0 = (double) psd_row(f)[a.minbin];            // `double max = psd[minbin]` (avg.c:111)
    const int cand = (arg >= 0 && mx > m0) ? arg : -1;
    const double vmax = (cand >= 0) ? mx : m0;
    // *peakbin after this frame: the new candidate, else the caller's value -- known only for
    // the very first frame of the call; otherwise the frame is flagged for the in-order kernel
    const bool carry_known = (cand >= 0) || (fl == 0);
    const int pk = (cand >= 0) ? cand : a.peakbin_init;
    double avgspec = 0.0, retv;
    if (a.mode == 2) {
      retv = (sum - vmax) / ((double) (band - 1) * (double) (eff + 1));
    } else {
      avgspec = (sum - vmax) / (double) (band - 1);
      retv = vmax / avgspec;
    }
    OutT *orow = out_base + fl * a.out_stride;
    double var = 0.0;
    int cnt = 0;
    // outside the band the row is the constant 1e-15 (avg.c:152-153): plain streaming stores
    const bool out_db = sizeof(OutT) == 4 && a.rows_db;
    const OutT fill = (OutT) (out_db ? -150.0 : 1e-15);
    auto fill_range = [&](int lo, int hi) {
      if (sizeof(OutT) == 4) {
        // rows have an odd stride: align to 16 bytes per row, then 128-bit stores
        float *base = reinterpret_cast<float *>(orow);
        int head = (int) (((16 - (reinterpret_cast<uintptr_t>(base + lo) & 15)) & 15) >> 2);
        if (head > hi - lo) head = hi - lo;
        if (lane < head) base[lo + lane] = (float) fill;
        const int body = (hi - lo - head) >> 2;
        float4 *b4 = reinterpret_cast<float4 *>(base + lo + head);
        const float4 f4 = make_float4((float) fill, (float) fill, (float) fill, (float) fill);
        for (int i = lane; i < body; i += 32) b4[i] = f4;
        const int done = lo + head + 4 * body;
        if (done + lane < hi) base[done + lane] = (float) fill;
      } else {
        for (int b = lo + lane; b < hi; b += 32) orow[b] = fill;
      }
    };
    fill_range(0, a.minbin < a.nbins ? a.minbin : a.nbins);
    if (a.maxbin < a.nbins) fill_range(a.maxbin, a.nbins);
    for (int b = a.minbin + lane; b < a.maxbin && b < a.nbins; b += 32) {
      double y;
      const double c = window_sum(b);
      if (a.mode == 2) {
        y = c / (double) (eff + 1);
      } else if (a.mode == 3) {
        y = a.max0 ? (c - mn) / (vmax - mn) : c / avgspec;
      } else if (c - avgspec > 0) {
        y = a.max0 ? (c - avgspec) / (vmax - avgspec) : c / avgspec;
        if (b != pk) { const double r = c / avgspec; var += r * r; cnt++; }
      } else {
        y = 1e-15;
      }
      if (out_db) y = 10.0 * log10(y);
      orow[b] = (OutT) y;
    }
    if (a.mode == 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        var += __shfl_xor_sync(0xffffffffu, var, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      }
    }
    if (lane == 0) {
      if (a.ret) a.ret[fl] = retv;
      if (a.peak_cand) a.peak_cand[fl] = cand;
      if (a.variance) a.variance[fl] = (a.mode == 1) ? var / (double) cnt : 0.0;
      if (a.mode == 1 && !carry_known && a.unresolved) atomicAdd(a.unresolved, 1);
    }
  }
}

extern "C" int glb_launch_avg(const glb_avg_args *a, void *stream) {
  if (!a || a->mode < 1 || a->mode > 3 || a->depth < 1 || a->minbin < 0 || a->maxbin < a->minbin) {
    glb_set_error("glb_launch_avg: invalid arguments");
    return GLB_EINVAL;
  }
  if (a->nframes <= 0) return GLB_OK;
  const int band = a->maxbin - a->minbin;
  const size_t smem = (size_t) (band > 0 ? band : 1) * sizeof(double);
  if (smem > 200 * 1024) { glb_set_error("glb_launch_avg: band too wide"); return GLB_EINVAL; }
  cudaStream_t st = (cudaStream_t) stream;
  if (!a->sequential && a->depth <= 32) {
    long long ctas = (a->nframes * 32 + 255) / 256;
    if (ctas > 148 * 64) ctas = 148 * 64;
    if (a->out_double) avg_direct_kernel<double><<<(int) ctas, 256, 0, st>>>(*a);
    else avg_direct_kernel<float><<<(int) ctas, 256, 0, st>>>(*a);
    CU(cudaGetLastError());
    g_launches++;
    return GLB_OK;
  }
  int chunk;
  if (a->sequential) {
    chunk = (int) std::min<long long>(a->nframes, 0x7fffffff);
  } else {
    // amortise the direct window sum of a chunk's first frame over >= 8*depth frames
    long long want = std::max<long long>(8LL * a->depth, 16);
    long long by_grid = (a->nframes + 148 * 8 - 1) / (148 * 8);
    chunk = (int) std::max(want, by_grid);
  }
  long long ctas = (a->nframes + chunk - 1) / chunk;
  if (a->out_double) {
    CU(cudaFuncSetAttribute(avg_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    avg_kernel<double><<<(int) ctas, 256, smem, st>>>(*a, chunk);
  } else {
    CU(cudaFuncSetAttribute(avg_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    avg_kernel<float><<<(int) ctas, 256, smem, st>>>(*a, chunk);
  }
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

// carried peak bin: last written candidate at or before each frame
__global__ void __launch_bounds__(1024) peak_carry_kernel(const int *__restrict__ cand, int *__restrict__ out,
                                                          long long n, int init) {
  __shared__ int last[1024];
  const int tid = threadIdx.x;
  const long long per = (n + 1023) / 1024;
  const long long b = tid * per, e = (b + per < n) ? b + per : n;
  int l = -1;
  for (long long i = b; i < e; i++) if (cand[i] >= 0) l = cand[i];
  last[tid] = l;
  __syncthreads();
  int carry = init;
  for (int w = 0; w < tid; w++) if (last[w] >= 0) carry = last[w];
  for (long long i = b; i < e; i++) {
    if (cand[i] >= 0) carry = cand[i];
    out[i] = carry;
  }
}

extern "C" int glb_launch_peak_carry(const int *cand, int *peakbin, long long nframes, int init, void *stream) {
  if (nframes <= 0) return GLB_OK;
  peak_carry_kernel<<<1, 1024, 0, (cudaStream_t) stream>>>(cand, peakbin, nframes, init);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

// ------------------------------------------------------------------------- half-complex PSD
// fft_psd (fft.c:203-226) for a spectrum that did not come from gram_kernel (callers that
// fill outbuf themselves).  phase = atan2(Re, Im), the reference's argument order.
__global__ void halfcomplex_psd_kernel(const float *__restrict__ hc, int n, float *__restrict__ psd,
                                       float *__restrict__ phase) {
  const int half = (n + 1) / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n / 2; i += gridDim.x * blockDim.x) {
    const bool edge = (i == 0) || (n % 2 == 0 && i == n / 2);
    if (!edge && i >= half) continue;
    const float re = hc[i];
    const float im = edge ? 0.f : hc[n - i];
    if (psd) psd[i] = (re * re + im * im) / (float) n;
    if (phase) phase[i] = edge ? 0.f : atan2f(re, im);
  }
}

extern "C" int glb_launch_halfcomplex_psd(const float *hc, int n, float *psd, float *phase, void *stream) {
  if (n < 1) { glb_set_error("halfcomplex_psd: n < 1"); return GLB_EINVAL; }
  const int ctas = (n / 2 + 1 + 255) / 256;
  halfcomplex_psd_kernel<<<ctas, 256, 0, (cudaStream_t) stream>>>(hc, n, psd, phase);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

// ------------------------------------------------------------------------- floor statistics
// compute_floor (fft.c:240-294) per PSD row: one CTA sorts the row descending in shared
// memory (bitonic network over the next power of two, padded with -inf), then takes the
// head (sig), sums the tail from index (int)(n * 0.95) (floor) and scans for the first
// maximum (peak value and bin; peak stays (0, 0) when no bin is > 0, fft.c:284-291).
__global__ void __launch_bounds__(512) floor_stats_kernel(const float *__restrict__ rows, long long stride, int nbins,
                                                          int npow2, float *__restrict__ stats) {
  extern __shared__ float srt[];
  __shared__ float red_v[16];
  __shared__ int red_i[16];
  const float *row = rows + (long long) blockIdx.x * stride;
  const int tid = threadIdx.x;
  float best = 0.f;
  int best_i = 0x7fffffff;
  for (int i = tid; i < npow2; i += 512) {
    const float v = (i < nbins) ? row[i] : -INFINITY;
    srt[i] = v;
    if (i < nbins && v > best) { best = v; best_i = i; }   // thread's bins ascend: first max
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += 512) {
        const int l = i ^ j;
        if (l > i) {
          const float a = srt[i], b = srt[l];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < b) : (a > b)) { srt[i] = b; srt[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  // tail sum (lowest 5 %): partial sums, then one thread adds the partials
  const int start = (int) (nbins * 0.95);
  float part = 0.f;
  for (int i = start + tid; i < nbins; i += 512) part += srt[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    part += __shfl_xor_sync(0xffffffffu, part, o);
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  __shared__ float red_s[16];
  if ((tid & 31) == 0) { red_s[tid >> 5] = part; red_v[tid >> 5] = best; red_i[tid >> 5] = best_i; }
  __syncthreads();
  if (tid == 0) {
    float fsum = 0.f, pv = 0.f;
    int pi = 0x7fffffff;
    for (int w = 0; w < 16; w++) {
      fsum += red_s[w];
      if (red_v[w] > pv || (red_v[w] == pv && red_i[w] < pi)) { pv = red_v[w]; pi = red_i[w]; }
    }
    float fl = fsum / 0.05f;
    fl = fl / (float) nbins;
    float *o = stats + (long long) blockIdx.x * 4;
    o[0] = srt[0];
    o[1] = fl;
    o[2] = (pv > 0.f) ? pv : 0.f;
    o[3] = (pv > 0.f) ? (float) pi : 0.f;
  }
}

extern "C" int glb_launch_floor_stats(const float *rows, long long stride, int nbins, long long nrows, float *stats,
                                      void *stream) {
  if (nrows <= 0) return GLB_OK;
  if (nbins < 1 || nbins > 32768) { glb_set_error("floor_stats: row too wide"); return GLB_EINVAL; }
  int np2 = 1;
  while (np2 < nbins) np2 <<= 1;
  const size_t smem = (size_t) np2 * sizeof(float);
  CU(cudaFuncSetAttribute(floor_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  floor_stats_kernel<<<(unsigned) nrows, 512, smem, (cudaStream_t) stream>>>(rows, stride, nbins, np2, stats);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

// ------------------------------------------------------------------------- display mapping
// main_window_draw, g_main.c:1109-1229, minus the GTK drawing: AGC of the display range from
// the per-row floor statistics, then level -> 8-bit palette index (-> RGB).
//
// agc_kernel: the display range is a recurrence over frames in mixed double/float arithmetic
// with a rounding to float at every step (static float display_max_lvl, g_main.c:1080,
// 1118-1123), so it is walked in order by one thread: O(frames) scalar work, exact semantics.
// state[0..1] = (display_max_lvl, display_min_lvl) carried between calls; stats = floor_stats rows.
__global__ void agc_kernel(const float *__restrict__ stats, long long nframes, long long first_frame, float overlap,
                           int log_scale, float *__restrict__ state, float *__restrict__ range /* [nframes][2] */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float mx = state[0], mn = state[1];
  for (long long i = 0; i < nframes; i++) {
    float sig = stats[4 * i + 0], flo = stats[4 * i + 1];
    if (first_frame + i == 0) {                      // glfer.first_buffer == TRUE (g_main.c:1112-1120)
      if (overlap > 0.0f) { sig /= overlap; flo /= overlap; }
      mx = sig;
      mn = flo;
    } else {                                         // g_main.c:1122-1123
      mx = (float) ((1.0 - 0.99) * (double) sig + 0.99 * (double) mx);
      mn = (float) ((1.0 - 0.99) * (double) flo + 0.99 * (double) mn);
    }
    if (log_scale) {                                 // g_main.c:1132-1135
      range[2 * i + 0] = (float) (10.0 * log10((double) mx));
      range[2 * i + 1] = (float) (10.0 * log10((double) mn));
    } else {
      range[2 * i + 0] = mx;
      range[2 * i + 1] = mn;
    }
  }
  state[0] = mx;
  state[1] = mn;
}

// levels_kernel: one warp per row.  Pixel i of a row shows bin n-1-i (g_main.c:1193-1201); in
// the log scales the level first passes through the reference's `short` level buffer
// (sig_level = levbuf[..] = 10 log10(x): integer-truncated dB, g_main.c:68,1193-1195).
__device__ __forceinline__ float short_db(float x) {
  const double d = 10.0 * log10((double) x);
  if (!(fabs(d) < 2147483648.0)) return 0.f;        // x86 cvttsd2si overflow -> 0x80000000 -> (short) 0
  return (float) (short) (int) d;
}

__global__ void __launch_bounds__(256) levels_kernel(const float *__restrict__ rows, long long stride, int nbins,
                                                     long long nframes, const float *__restrict__ range,
                                                     const float *__restrict__ fixed_range, int log_scale, float thr,
                                                     const unsigned char *__restrict__ colortab,
                                                     unsigned char *__restrict__ levels, unsigned char *__restrict__ rgb) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
  for (long long f = warp; f < nframes; f += nwarps) {
    const float dmax = range ? range[2 * f] : fixed_range[0];
    const float dmin = range ? range[2 * f + 1] : fixed_range[1];
    const float *row = rows + f * stride;
    for (int i = lane; i < nbins; i += 32) {
      const float x = row[nbins - 1 - i];
      const float sig_level = log_scale ? short_db(x) : x;
      const float fl = 255 * ((sig_level - dmin) / (dmax - dmin));
      unsigned char v;
      if ((double) fl < 255.0 * (double) thr) v = 0;
      else if (fl > 255) v = 255;
      else v = (unsigned char) (((double) fl - 255.0 * (double) thr) / (1.0 - (double) thr));
      if (levels) levels[f * nbins + i] = v;
      if (rgb) {
        unsigned char *px = rgb + (f * nbins + i) * 3;
        px[0] = colortab[3 * v];
        px[1] = colortab[3 * v + 1];
        px[2] = colortab[3 * v + 2];
      }
    }
  }
}

extern "C" int glb_launch_agc(const float *stats, long long nframes, long long first_frame, float overlap, int log_scale,
                              float *state, float *range, void *stream) {
  if (nframes <= 0) return GLB_OK;
  agc_kernel<<<1, 32, 0, (cudaStream_t) stream>>>(stats, nframes, first_frame, overlap, log_scale, state, range);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}

extern "C" int glb_launch_levels(const float *rows, long long stride, int nbins, long long nframes, const float *range,
                                 const float *fixed_range, int log_scale, float thr, const unsigned char *colortab,
                                 unsigned char *levels, unsigned char *rgb, void *stream) {
  if (nframes <= 0) return GLB_OK;
  if ((!range && !fixed_range) || (rgb && !colortab)) { glb_set_error("glb_launch_levels: missing range / palette"); return GLB_EINVAL; }
  long long ctas = (nframes * 32 + 255) / 256;
  if (ctas > 148 * 64) ctas = 148 * 64;
  levels_kernel<<<(int) ctas, 256, 0, (cudaStream_t) stream>>>(rows, stride, nbins, nframes, range, fixed_range, log_scale, thr,
                                                              colortab, levels, rgb);
  CU(cudaGetLastError());
  g_launches++;
  return GLB_OK;
}
