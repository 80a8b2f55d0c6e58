// gram_part.cu -- instantiates the spectrogram kernels for one subset of FFT sizes; compiled once
// per subset (-DGLB_PART=0..3) so that the sizes build in parallel.  Returns -1 when the size
// belongs to another part.
#include "gram_common.cuh"

#ifndef GLB_PART
#error "compile with -DGLB_PART=0..3"
#endif

#define GLB_CASE(M_) case M_: return launch_gram_m<M_>(k, multi, groups_hint, st, allow);
#define GLB_PART_FN_(K) glb_gram_part_##K
#define GLB_PART_FN(K) GLB_PART_FN_(K)

int GLB_PART_FN(GLB_PART)(int m, const KParams &k, bool multi, int groups_hint, cudaStream_t st, int allow) {
  switch (m) {
#if GLB_PART == 0
    GLB_CASE(16) GLB_CASE(32) GLB_CASE(64) GLB_CASE(128) GLB_CASE(256) GLB_CASE(512)
#elif GLB_PART == 1
    GLB_CASE(1024) GLB_CASE(4096)
#elif GLB_PART == 2
    GLB_CASE(2048)
#else
    GLB_CASE(8192) GLB_CASE(16384)
#endif
    default: break;
  }
  return -1;
}
