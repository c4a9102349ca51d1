// Last-CTA fold of the per-CTA BatchNorm partial statistics written by the convolution kernels.
#pragma once
#include "ptk_common.cuh"
#include "ptk_kernels.h"

namespace ptk {

// Every CTA has written part[cta][c] = (mean, M2) over its `npix` conv outputs of channel c.
// The last CTA to arrive folds all partials (Chan et al. parallel variance, in double) into
// stats_out = mean[32], rstd[32] (biased variance, eps 1e-3: models/InT.py:102).  >= 256 threads (eight warps fold).
__device__ __forceinline__ void conv_stats_fold(const ConvArgs& a, int nCta, int npix, double (*sFin)[C][3],
                                                int* sIsLast) {
  const int tid = threadIdx.x;
  __threadfence();
  __syncthreads();
  if (tid == 0) *sIsLast = (atomicAdd(a.counter, 1u) == (unsigned)(nCta - 1));
  __syncthreads();
  if (*sIsLast) {
    __threadfence();
    const int c = tid & 31, g = tid >> 5;
    // the whole grid waits for this CTA: keep 16 independent loads in flight per thread (eight warps; further warps idle)
    double s1 = 0, s2 = 0, s3 = 0;
    int i = g < 8 ? g : nCta;
    for (; i + 8 * 15 < nCta; i += 8 * 16) {
      float2 pm[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pm[j] = __ldcg(reinterpret_cast<const float2*>(a.part + ((size_t)(i + 8 * j) * C + c) * 2));
#pragma unroll
      for (int j = 0; j < 16; ++j) { s1 += pm[j].x; s2 += (double)pm[j].x * pm[j].x; s3 += pm[j].y; }
    }
    for (; i < nCta; i += 8) {
      const float2 pm = __ldcg(reinterpret_cast<const float2*>(a.part + ((size_t)i * C + c) * 2));
      s1 += pm.x;
      s2 += (double)pm.x * pm.x;
      s3 += pm.y;
    }
    if (g < 8) { sFin[g][c][0] = s1; sFin[g][c][1] = s2; sFin[g][c][2] = s3; }
    __syncthreads();
    if (tid < C) {
      s1 = s2 = s3 = 0;
      for (int gg = 0; gg < 8; ++gg) { s1 += sFin[gg][tid][0]; s2 += sFin[gg][tid][1]; s3 += sFin[gg][tid][2]; }
      const double n = nCta, m = s1 / n;
      double m2 = s3 + (double)npix * (s2 - n * m * m);
      if (m2 < 0) m2 = 0;
      const double var = m2 / (n * npix);
      a.stats_out[tid] = (float)m;
      a.stats_out[C + tid] = (float)(1.0 / sqrt(var + (double)BN_EPS));
    }
    if (tid == 0) *a.counter = 0u;
  }
}

// The same fold done by the CONSUMER of the statistics (the point-wise kernel that follows the convolution; every CTA folds all
// partials itself, fixed order => every CTA gets the same bits).  The convolution then ends with its last store: no
// __threadfence / atomic arrival / last-CTA fold on the critical path of every forward convolution.
// Threads 0..127: warp w owns channels 8 w .. 8 w + 7, lane = (group g = lane >> 3, channel); group g sums partials g, g + 4, ...
// in double, the four groups meet through shuffles -- no shared memory (the callers run 4 CTAs per SM with none to spare).
// Writes scale = w * rstd and shift = b - mean * scale; stat_out (mean[32], rstd[32]: the checkpoint the reverse loop reads) is
// written by the one CTA that passes a non-null pointer.  The caller syncs before using sScale / sShift.
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void consumer_stats_fold(const float* part, int nCta, int npix, const float* bn_w, const float* bn_b,
                                                    float* sScale, float* sShift, float* stat_out) {
  const int tid = threadIdx.x;
  if (tid >= 128) return;
  const int lane = tid & 31, c = (tid >> 5) * 8 + (lane & 7), g = lane >> 3;
  double s1 = 0, s2 = 0, s3 = 0;
  int i = g;
  for (; i + 4 * 15 < nCta; i += 4 * 16) {
    float2 pm[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pm[j] = __ldcg(reinterpret_cast<const float2*>(part + ((size_t)(i + 4 * j) * C + c) * 2));
#pragma unroll
    for (int j = 0; j < 16; ++j) { s1 += pm[j].x; s2 += (double)pm[j].x * pm[j].x; s3 += pm[j].y; }
  }
  for (; i < nCta; i += 4) {
    const float2 pm = __ldcg(reinterpret_cast<const float2*>(part + ((size_t)i * C + c) * 2));
    s1 += pm.x;
    s2 += (double)pm.x * pm.x;
    s3 += pm.y;
  }
#pragma unroll
  for (int m = 8; m <= 16; m <<= 1) {
    s1 += shfl_xor_f64(s1, m);
    s2 += shfl_xor_f64(s2, m);
    s3 += shfl_xor_f64(s3, m);
  }
  if (g == 0) {
    const double n = nCta, mu = s1 / n;
    double m2 = s3 + (double)npix * (s2 - n * mu * mu);
    if (m2 < 0) m2 = 0;
    const double var = m2 / (n * npix);
    const float mean = (float)mu, rstd = (float)(1.0 / sqrt(var + (double)BN_EPS));
    if (stat_out) { stat_out[c] = mean; stat_out[C + c] = rstd; }
    const float sc = __ldg(bn_w + c) * rstd;
    sScale[c] = sc;
    sShift[c] = __ldg(bn_b + c) - mean * sc;
  }
}

}  // namespace ptk
