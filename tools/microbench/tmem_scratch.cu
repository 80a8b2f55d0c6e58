// Micro-test: tensor memory (TMEM) as per-thread scratch on sm_100a.  Every CTA (128 threads)
// allocates 64 columns; thread t owns TMEM lane t (warp w reaches lanes 32w..32w+31 with the
// 32x32b shape); it stores 48 words, reads them back in chunks of 8 and checks them.  Many CTAs
// per SM allocate concurrently.  Prints mismatches and the cycles of one 8-word load + wait.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(addr)
               : "memory");
}

__global__ void __launch_bounds__(128) k(int *bad, long long *cyc, int iters) {
  __shared__ uint32_t slot;
  const int t = threadIdx.x, w = t >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t) __cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t) (w * 32) << 16);
  int nbad = 0;
  long long t0 = 0, t1 = 0;
  for (int it = 0; it < iters; it++) {
    for (int c = 0; c < 48; c += 8) {
      uint32_t v[8];
      for (int i = 0; i < 8; i++) v[i] = (blockIdx.x * 1315423911u) ^ (t * 4096 + (c + i) * 7 + it);
      tmem_st8(base + c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    t0 = clock64();
    for (int c = 0; c < 48; c += 8) {
      uint32_t v[8];
      tmem_ld8(base + c, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; i++) nbad += v[i] != ((blockIdx.x * 1315423911u) ^ (t * 4096 + (c + i) * 7 + it));
    }
    t1 = clock64();
  }
  if (nbad) atomicAdd(bad, nbad);
  if (blockIdx.x == 0 && t == 0) *cyc = (t1 - t0) / 6;
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot) : "memory");
}

int main() {
  int *bad;
  long long *cyc;
  cudaMallocManaged(&bad, sizeof(int));
  cudaMallocManaged(&cyc, sizeof(long long));
  *bad = 0;
  k<<<148 * 12, 128>>>(bad, cyc, 50);
  cudaError_t e = cudaDeviceSynchronize();
  printf("tmem scratch: %s, mismatches %d, cycles per (ld.x8 + wait) %lld\n", cudaGetErrorString(e), *bad, *cyc);
  return (e != cudaSuccess) || *bad;
}
