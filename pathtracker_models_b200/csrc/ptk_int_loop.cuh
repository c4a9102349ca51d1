// Shared device code of the persistent forward time-loop kernels (ptk_int_loop.cu, ptk_int_loop2.cu): shared-memory plan,
// barriers between co-resident CTAs, the per-CTA BatchNorm fold and the per-tile tcgen05 convolution (schedule of conv7_tc2).
#pragma once
#include <cuda_bf16.h>

#include "ptk_conv_stats.cuh"
#include "ptk_conv_tc2.cuh"
#include "ptk_int.h"
#include "ptk_int_fwd_tiles.cuh"
#include "ptk_profile.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int LP_THREADS = 512;
// convolution view
constexpr int LP_RING = 0;
constexpr int LP_IMG = LP_RING + T2_NST * T2_STAGE;
constexpr int LP_PART = LP_IMG + T2_A_BYTES;
constexpr int LP_BARS = LP_PART + 2 * T2_PART;
constexpr int LP_SMEM = LP_BARS + 256;
// point-wise view of the same bytes
constexpr int LP_PW_W = 0;                                        // up to 4 gate matrices in B-fragment order
constexpr int LP_PW_TILES = LP_PW_W + 4 * C * C * 4;              // 16 warps x 2 tiles
constexpr int LP_PW_FOLD = LP_PW_TILES + 16 * 2 * TILE_F * 4;     // double [8][32][3]
static_assert(LP_PW_FOLD + 8 * C * 3 * 8 <= LP_BARS, "point-wise view must fit under the barriers");
static_assert(LP_IMG % 128 == 0 && LP_PART % 16 == 0 && LP_BARS % 8 == 0, "alignment");

// tuning build only (-DPTK_DEBUG_STAMPS=1, tools/loop_stamps.py): per-CTA cycle sums of the phases of a frame, thread 0's view.
// The translation unit that owns the counters defines PTK_LOOP_STAMPS_TU before including this header.
#if PTK_DEBUG_STAMPS && defined(PTK_LOOP_STAMPS_TU)
__device__ long long g_loop_acc[148][16];
#define LP_STAMP_DECL long long lp_acc[16] = {0}; long long lp_prev = clock64();
#define LP_STAMP(i) do { if (threadIdx.x == 0) { const long long now_ = clock64(); lp_acc[i] += now_ - lp_prev; lp_prev = now_; } } while (0)
#define LP_STAMP_STORE(cta) do { if (threadIdx.x == 0 && (cta) < 148) for (int i_ = 0; i_ < 16; ++i_) g_loop_acc[cta][i_] = lp_acc[i_]; } while (0)
#else
#define LP_STAMP_DECL
#define LP_STAMP(i) do { } while (0)
#define LP_STAMP_STORE(cta) do { } while (0)
#endif

struct LoopArgs {
  float* part[2];       // BatchNorm partial statistics [nCta][32][2], one buffer per BatchNorm
  unsigned* gctr;       // grid barrier counter (monotonic, zeroed by the host before the launch)
  unsigned* sctr;       // [B] per-sample barrier counters (monotonic)
  int tiles_per_sample;
};

// ---- barriers between CTAs (all CTAs are co-resident: cooperative launch) ------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// arrive: everything this CTA wrote before is released to the GPU scope
__device__ __forceinline__ void cta_arrive(unsigned* ctr) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    red_release_gpu(ctr);
  }
}
// wait until `target` arrivals; afterwards every thread of the CTA may read what the arriving CTAs wrote
__device__ __forceinline__ void cta_wait(const unsigned* ctr, unsigned target) {
  if (threadIdx.x == 0) {
    while (ld_acquire_gpu(ctr) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// ---- BatchNorm statistics: every CTA folds all partials itself, in the order of conv_stats_fold (ptk_conv_stats.cuh) ----
__device__ __forceinline__ void loop_fold(const float* part, int nCta, int npix, double (*sFin)[C][3], float* sScale, float* sShift,
                                          const float* bn_w, const float* bn_b, float* stat_out) {
  const int tid = threadIdx.x;
  if (tid < 256) {
    const int c = tid & 31, g = tid >> 5;
    double s1 = 0, s2 = 0, s3 = 0;
    int i = g;
    for (; i + 8 * 15 < nCta; i += 8 * 16) {
      float2 pm[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pm[j] = __ldcg(reinterpret_cast<const float2*>(part + ((size_t)(i + 8 * j) * C + c) * 2));
#pragma unroll
      for (int j = 0; j < 16; ++j) { s1 += pm[j].x; s2 += (double)pm[j].x * pm[j].x; s3 += pm[j].y; }
    }
    for (; i < nCta; i += 8) {
      const float2 pm = __ldcg(reinterpret_cast<const float2*>(part + ((size_t)i * C + c) * 2));
      s1 += pm.x;
      s2 += (double)pm.x * pm.x;
      s3 += pm.y;
    }
    sFin[g][c][0] = s1; sFin[g][c][1] = s2; sFin[g][c][2] = s3;
  }
  __syncthreads();
  if (tid < C) {
    double s1 = 0, s2 = 0, s3 = 0;
    for (int gg = 0; gg < 8; ++gg) { s1 += sFin[gg][tid][0]; s2 += sFin[gg][tid][1]; s3 += sFin[gg][tid][2]; }
    const double n = nCta, m = s1 / n;
    double m2 = s3 + (double)npix * (s2 - n * m * m);
    if (m2 < 0) m2 = 0;
    const double var = m2 / (n * npix);
    const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)BN_EPS));
    if (stat_out) { stat_out[tid] = mean; stat_out[C + tid] = rstd; }   // checkpoint for the reverse loop
    const float sc = __ldg(bn_w + tid) * rstd;
    sScale[tid] = sc;
    sShift[tid] = __ldg(bn_b + tid) - mean * sc;
  }
  __syncthreads();
}

// ---- one 7x7 convolution of this CTA's tile (the schedule of conv7_tc2_kernel, 512 threads) ----------------------------------
// kpar: 0 for the first convolution of a frame, 1 for the second.  The mbarriers live for the whole kernel; per convolution
// the ring barriers of slots 0 / 1 complete 11 phases, those of slots 2 / 3 ten, dFull[0] / dEmpty[0] two, dFull[1] / dEmpty[1],
// stA, stB one: the parity a wait has to look for is the one of the stand-alone kernel xor (odd phase count ? kpar : 0).
template <bool SINGLE>
__device__ __forceinline__ void loop_conv(const float* __restrict__ in, const uint8_t* __restrict__ wsrc, float* out, float* part,
                                          int cta, int H, int W, int x0, int y0, int b, uint8_t* smem, float* sMerge,
                                          uint32_t tmem_base, uint32_t kpar, bool own_rows_staged = false) {
  uint8_t* sWr = smem + LP_RING;
  uint8_t* sA = smem + LP_IMG;
  float* sPart = reinterpret_cast<float*>(smem + LP_PART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LP_BARS);
  uint64_t* full = bars;
  uint64_t* empty = bars + T2_NST;
  uint64_t* dFull = bars + 2 * T2_NST;
  uint64_t* dEmpty = dFull + 2;
  uint64_t* stA = dEmpty + 2;
  uint64_t* stB = stA + 1;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const size_t bs = (size_t)C * H * W;
  constexpr int NSTG = T2_NCH * T2_NG;                      // 42 weight stages

  // the region was used by the point-wise phase (generic proxy): zero the image, then hand the ring to the async proxy.
  // own_rows_staged (ptk_int_loop2.cu): the point-wise phase has zeroed the image and written the tile's own 16 rows into it
  // as bf16 hi/lo already; only the halo rows (and, for W > 32, the halo columns) are still to be fetched.
  if (!own_rows_staged)
    for (int i = tid; i < T2_A_BYTES / 16; i += LP_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < T2_NST; ++i) {
      mbar_expect_tx(&full[i], T2_STAGE);
      bulk_g2s(sWr + i * T2_STAGE, wsrc + (size_t)i * T2_STAGE, T2_STAGE, &full[i]);
    }
  }

  // ---- stage the input tile: fp32 NCHW -> bf16 hi/lo, channel-chunk planar (the image of ptk_conv_tc2.cu) ------------
  // one item = (chunk, row, 4-pixel group): 8 LDG.128 (one per channel) -> 4 pixels; one item per thread and phase
  auto stage_item = [&](int it, int limit) {
    const int q = (it >> 1) & 3, row = it >> 5, g = ((it >> 3) & 3) * 2 + (it & 1);
    const int gy = y0 + row - 3;
    if (it >= limit || gy < 0 || gy >= H) return;
    const int p = row * T2_WP + 3 + 4 * g;
    const float* src = in + (size_t)b * bs + ((size_t)(q * 8) * H + gy) * W + x0 + 4 * g;
    float4 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)e * H * W));
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v0 = reinterpret_cast<const float*>(&v[2 * e])[px], v1 = reinterpret_cast<const float*>(&v[2 * e + 1])[px];
        split_bf16x2(v0, v1, hi[e], lo[e]);
      }
      *reinterpret_cast<uint4*>(sA + q * T2_PLANE + (p + px) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + 4 * T2_PLANE + q * T2_PLANE + (p + px) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  };
  constexpr int NIT_A = 4 * 12 * (T2_TW / 4), NIT = 4 * T2_ROWS * (T2_TW / 4);      // 384, 704
  if (!own_rows_staged) stage_item(tid, NIT_A);             // phase A: input rows 0-11 (first row chunk)
  else if (tid < 192) stage_item(tid < 96 ? tid : tid + 32 * 16, NIT);   // halo rows 0-2 and 19-21 (items of row r start at 32 r)
  if (W > T2_TW) {   // halo columns (only inside the image when W > 32)
    for (int it = tid; it < 4 * T2_ROWS * 6; it += LP_THREADS) {
      const int q = it / (T2_ROWS * 6), rem = it - q * (T2_ROWS * 6), r = rem / 6, hc = rem % 6;
      const int c = hc < 3 ? hc : (T2_TW + hc);          // padded columns 0..2 and 35..37
      const int gy = y0 + r - 3, gx = x0 + c - 3;
      if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
      const float* src = in + (size_t)b * bs + ((size_t)(q * 8) * H + gy) * W + gx;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v0 = __ldcg(src + (size_t)(2 * e) * H * W), v1 = __ldcg(src + (size_t)(2 * e + 1) * H * W);
        split_bf16x2(v0, v1, hi[e], lo[e]);
      }
      const int p = r * T2_WP + c;
      *reinterpret_cast<uint4*>(sA + q * T2_PLANE + p * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + 4 * T2_PLANE + q * T2_PLANE + p * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
  mbar_arrive(stA);
  if (warp >= 2) {                                          // phase B: input rows 12-21, warps 2-15 (448 threads, 320 items)
    if (!own_rows_staged) stage_item(NIT_A + tid - 64, NIT);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(stB);
  }
  const uint32_t p01 = kpar;                                // parity offset of barriers with an odd phase count per convolution

  if (warp == 0) {
    // ---- weight producer ---------------------------------------------------------------------------------------
    if (lane == 0) {
      for (int i = T2_NST; i < NSTG; ++i) {
        const int st = i % T2_NST, round = i / T2_NST;
        mbar_wait(&empty[st], (uint32_t)((round - 1) & 1) ^ (st < 2 ? p01 : 0u));
        mbar_expect_tx(&full[st], T2_STAGE);
        bulk_g2s(sWr + st * T2_STAGE, wsrc + (size_t)(i % T2_NG) * T2_STAGE, T2_STAGE, &full[st]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer --------------------------------------------------------------------------------------------
    const uint32_t bHiA = smem_u32(sA), bLoA = bHiA + 4 * T2_PLANE, wBase = smem_u32(sWr);
    const uint64_t dB0 = umma_desc(0, T2_PLANE, 128);          // pixels: K-major, LBO = next 8-channel plane, SBO = next 8 pixels
    const uint64_t dW0 = umma_desc(0, 128 * 16, 128);          // weights: K-major, LBO = next k-chunk, SBO = next 8 rows
    const uint32_t bHiW = (uint32_t)(dB0 >> 32), bLoW = (uint32_t)dB0;
    const uint32_t wHiW = (uint32_t)(dW0 >> 32), wLoW = (uint32_t)dW0;
    uint32_t bH[2], bL[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      bH[s] = bLoW + ((bHiA + (uint32_t)(2 * s) * T2_PLANE) >> 4);
      bL[s] = bLoW + ((bLoA + (uint32_t)(2 * s) * T2_PLANE) >> 4);
    }
    int i = 0;
#pragma unroll 1
    for (int c = 0; c < T2_NCH; ++c) {
      const int buf = c & 1;
      if (c < 2) {                                             // chunk 0 reads input rows 0-11, chunks 1 / 2 rows 6-21
        mbar_wait(c == 0 ? stA : stB, p01);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (c >= 2) {                                            // the buffer's previous chunk must have been drained
        mbar_wait(&dEmpty[buf], 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t idesc = umma_idesc(t2_chunk_n(c));
      const uint32_t dcol = tmem_base + (uint32_t)(buf * T2_DBUF);
#pragma unroll 1
      for (int g = 0; g < T2_NG; ++g, ++i) {
        const int st = i % T2_NST, round = i / T2_NST;
        mbar_wait(&full[st], (uint32_t)(round & 1) ^ (st < 2 ? p01 : 0u));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t po = (uint32_t)(t2_chunk_start(c) + (g >> 1) * T2_WP + 4 * (g & 1));     // 16-byte units == pixels
        const uint32_t wS = wLoW + ((wBase + (uint32_t)st * T2_STAGE) >> 4);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const uint64_t dWh = ((uint64_t)wHiW << 32) | (wS + (uint32_t)((s * 2 * T2_WTILE) >> 4));
          const uint64_t dWl = ((uint64_t)wHiW << 32) | (wS + (uint32_t)((s * 2 * T2_WTILE + T2_WTILE) >> 4));
          const uint64_t dIh = ((uint64_t)bHiW << 32) | (bH[s] + po), dIl = ((uint64_t)bHiW << 32) | (bL[s] + po);
          if (elect_one()) {
            umma_bf16(dcol, dWh, dIh, idesc, (g | s) ? 1u : 0u);
            if (!SINGLE) {
              umma_bf16(dcol, dWh, dIl, idesc, 1u);
              umma_bf16(dcol, dWl, dIh, idesc, 1u);
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty[st]);
      }
      __syncwarp();
      if (elect_one()) umma_commit(&dFull[buf]);
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 4 + T2_EPI_WARPS) {
    // ---- epilogue: two groups of four warps drain alternate 32-pixel blocks (t2_epilogue, ptk_conv_tc2.cuh) --------------------
    const T2Epi e{out + (size_t)b * bs, H, W, x0, y0, 0, part + (size_t)cta * C * 2};
    t2_epilogue(e, tmem_base, sPart, sMerge, dFull, dEmpty, p01, warp - 4, lane);
  }
  // every asynchronous operation of this convolution has completed once the epilogue warps are through (they waited for the
  // last accumulator commit, which follows every ring commit); the caller's next barrier releases the region to the point-wise view
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

}  // namespace ptk
