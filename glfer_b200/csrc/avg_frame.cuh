// avg_frame.cuh -- one frame of the sliding per-bin averaging (update_avg_*, avg.c:108-298) by one warp, shared
// by the stand-alone averaging kernel (window sums re-read from PSD rows in HBM / L2) and by the ring
// kernel's fused averaging (window sums from the band history it keeps in shared memory): the same code,
// hence the same bits, whichever path produced the row.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/glb_shim.h"

// RN(x / d) for a small integer divisor d with r = RN(1 / d) (Markstein: q = RN(x r); e = x - q d exactly, by
// FMA; RN(q + e r) = RN(x / d) when r is the correctly rounded reciprocal and q is faithful).  Checked by
// brute force against IEEE division for d = 1 .. 63.  Three dependent operations instead of a division.
__device__ __forceinline__ double avg_div_small(double x, double d, double r) {
  const double q = __dmul_rn(x, r);
  const double e = __fma_rn(-q, d, x);
  return __fma_rn(e, r, q);
}

// psd_at(g, b): PSD of frame g (index since alloc_avg) at bin b.  fl: index of the frame in the launch's
// outputs (f = a.first_frame + fl).  Pass 1 forms the window sums of the band (a lane's bins lane, lane + 32,
// ...; the first four stay in registers) and reduces max / first argmax / sum (/ min) with shuffles, pass 2
// writes the normalised row.  MODE: avgmode_t (1 sumavg, 2 plain, 3 sumextreme), compile time so that each
// carrier only executes what its normalisation needs.
template <typename OutT, int MODE, class PsdAt>
__device__ __forceinline__ void avg_frame_warp_m(const glb_avg_args &a, long long fl, int lane, PsdAt &&psd_at) {
  constexpr int KREG = 4;                                  // window sums kept in registers per lane
  const int band = a.maxbin - a.minbin;
  OutT *out_base = (OutT *) a.avg_rows;
  const long long f = a.first_frame + fl;
  const int eff = (int) ((f + 1 < a.depth) ? f + 1 : a.depth);
  const long long g0 = f - eff + 1;
  auto window_sum = [&](int b) {
    // oldest to newest, as the reference's cum[] would hold them; four loads in flight at a time (the loads do
    // not depend on the sum: issued one by one they cost a full memory latency each)
    double c = 0.0;
    int d = 0;
#pragma unroll 1
    for (; d + 4 <= eff; d += 4) {
      const float v0 = psd_at(g0 + d, b), v1 = psd_at(g0 + d + 1, b), v2 = psd_at(g0 + d + 2, b), v3 = psd_at(g0 + d + 3, b);
      c += (double) v0;
      c += (double) v1;
      c += (double) v2;
      c += (double) v3;
    }
#pragma unroll 1
    for (; d < eff; ++d) c += (double) psd_at(g0 + d, b);
    return c;
  };
  // (loops deliberately NOT unrolled: this code runs once per frame on one warp, inlined into the spectrogram
  // kernel; unrolled it was 3 000 instructions per mode and evicted the FFT loop from the instruction cache)
  double creg[KREG];
  double mx = -1.0, sum = 0.0, mn = 1.0;
  int arg = -1;
#pragma unroll 1
  for (int i = lane, k = 0; i < band; i += 32, ++k) {
    const int b = a.minbin + i;
    const double c = window_sum(b);
    if (k < KREG) creg[k] = c;
    if (arg < 0 || c > mx) { mx = c; arg = b; }
    sum += c;
    if (MODE == 3 && c < mn) mn = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
    if (oarg >= 0 && (arg < 0 || omx > mx || (omx == mx && oarg < arg))) { mx = omx; arg = oarg; }
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (MODE == 3) {
      const double omn = __shfl_xor_sync(0xffffffffu, mn, o);
      if (omn < mn) mn = omn;
    }
  }
  const double m0 = (double) psd_at(f, a.minbin);              // `double max = psd[minbin]` (avg.c:111)
  const int cand = (arg >= 0 && mx > m0) ? arg : -1;
  const double vmax = (cand >= 0) ? mx : m0;
  // *peakbin after this frame: the new candidate, else the caller's value -- known only for
  // the very first frame of the call; otherwise the frame is flagged for the in-order kernel
  const bool carry_known = (cand >= 0) || (fl == 0);
  const int pk = (cand >= 0) ? cand : a.peakbin_init;
  const double deff = (double) (eff + 1), reff = 1.0 / deff;
  double avgspec = 0.0, retv;
  if (MODE == 2) {
    retv = (sum - vmax) / ((double) (band - 1) * deff);
  } else {
    avgspec = (sum - vmax) / (double) (band - 1);
    retv = vmax / avgspec;
  }
  OutT *orow = out_base + fl * a.out_stride;
  double var = 0.0;
  int cnt = 0;
  // outside the band the row is the constant 1e-15 (avg.c:152-153): plain streaming stores
  const bool out_db = sizeof(OutT) == 4 && a.rows_db;
  if (!a.band_only) {
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
  }
  const int ob0 = a.band_only ? a.minbin : 0;
  auto emit = [&](int b, double c) {
    double y;
    if (MODE == 2) {
      y = avg_div_small(c, deff, reff);                         // c / (effdepth + 1), avg.c:150
    } else if (MODE == 3) {
      y = a.max0 ? (c - mn) / (vmax - mn) : c / avgspec;
    } else if (c - avgspec > 0) {
      y = a.max0 ? (c - avgspec) / (vmax - avgspec) : c / avgspec;
      if (b != pk) { const double r = c / avgspec; var += r * r; cnt++; }
    } else {
      y = 1e-15;
    }
    if (out_db) y = 10.0 * log10(y);
    orow[b - ob0] = (OutT) y;
  };
#pragma unroll 1
  for (int b = a.minbin + lane, k = 0; b < a.maxbin && b < a.nbins; b += 32, ++k) emit(b, k < KREG ? creg[k] : window_sum(b));
  if (MODE == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      var += __shfl_xor_sync(0xffffffffu, var, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
  }
  if (lane == 0) {
    if (a.ret) a.ret[fl] = retv;
    if (a.peak_cand) a.peak_cand[fl] = cand;
    if (a.variance) a.variance[fl] = (MODE == 1) ? var / (double) cnt : 0.0;
    if (MODE == 1 && !carry_known && a.unresolved) atomicAdd(a.unresolved, 1);
  }
}

template <typename OutT, class PsdAt>
__device__ __forceinline__ void avg_frame_warp(const glb_avg_args &a, long long fl, int lane, PsdAt &&psd_at) {
  if (a.mode == 2) avg_frame_warp_m<OutT, 2>(a, fl, lane, psd_at);
  else if (a.mode == 3) avg_frame_warp_m<OutT, 3>(a, fl, lane, psd_at);
  else avg_frame_warp_m<OutT, 1>(a, fl, lane, psd_at);
}
