// Persistent forward time loop of the InT / hGRU circuit: ONE cooperative kernel walks all T frames.
//
// Reference being replaced: the Python loop `for t in range(x.shape[2])` of models/InT.py:223-235 (and
// models/ffhgru_hierarchy.py:246-262) around rCell.forward / hConvGRUCell.forward.  The PDL chain of ptk_int_fwd.cu runs
// four launches per frame (conv, mid, conv, post); here one CTA per 16 x 32-pixel tile (the tile of conv7_tc2) stays
// resident for the whole clip and runs the same four phases itself:
//
//   conv0 (tcgen05, Ge_t -> Y0_t + BatchNorm partials)   -- grid barrier: ARRIVE
//   mid, the part that needs no batch statistics (excitation gate (9), stem, inhibition-gate pre-activation (7))
//                                                        -- grid barrier: WAIT, fold the partials (every CTA, same order)
//   mid, the BatchNorm0-dependent rest (eqs 3b-8) -> I_{t+1}        -- per-sample barrier (the 3-row halo of the next conv)
//   conv1 (I_{t+1} -> Y1_t + partials)                   -- grid barrier: ARRIVE
//   post operand loads                                   -- grid barrier: WAIT, fold
//   post (eqs 10b-13 -> E_{t+1}; attention (1)-(2) of frame t+1 -> A, Ge)        -- per-sample barrier
//
// What residency buys (there is no second copy of the arithmetic: the tile bodies are the ones of the chain, ptk_int_fwd_tiles.cuh,
// and the MMA / epilogue schedule is conv7_tc2's, so results are bit-identical to the chain):
//  * TMEM allocation, the per-channel parameter vectors and the kernel prologues happen once, not 4 T times;
//  * the two BatchNorm reductions are split arrive / wait barriers, so the statistics-independent half of `mid` and the
//    operand loads of `post` run while the slowest CTA is still convolving;
//  * all 16 warps stage the operand image (the stand-alone convolution has 8).
// Training and inference: the checkpoints the reverse loop reads are written exactly as by the chain.
//
// Shared memory is one 215 KB region used in turn as the convolution's [weight ring | operand image | quadrant tiles]
// and as the point-wise [gate matrices | 16 x 2 warp tiles | fold scratch]; the image is re-zeroed per convolution.
#include <cuda_bf16.h>

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
                                          int cta, int H, int W, int x0, int y0, int b, uint8_t* smem, uint32_t tmem_base,
                                          uint32_t kpar) {
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

  // the region was used by the point-wise phase (generic proxy): zero the image, then hand the ring to the async proxy
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
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        hi[e] = pack_bf16x2(h0, h1);
        lo[e] = pack_bf16x2(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
      }
      *reinterpret_cast<uint4*>(sA + q * T2_PLANE + (p + px) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + 4 * T2_PLANE + q * T2_PLANE + (p + px) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  };
  constexpr int NIT_A = 4 * 12 * (T2_TW / 4), NIT = 4 * T2_ROWS * (T2_TW / 4);      // 384, 704
  stage_item(tid, NIT_A);                                   // phase A: input rows 0-11 (first row chunk)
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
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        hi[e] = pack_bf16x2(h0, h1);
        lo[e] = pack_bf16x2(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
      }
      const int p = r * T2_WP + c;
      *reinterpret_cast<uint4*>(sA + q * T2_PLANE + p * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + 4 * T2_PLANE + q * T2_PLANE + p * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
  mbar_arrive(stA);
  if (warp >= 2) {                                          // phase B: input rows 12-21, warps 2-15 (448 threads, 320 items)
    stage_item(NIT_A + tid - 64, NIT);
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
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue warps: quadrant dl = warp - 4; out[oc][m] = sum_dl D[dl][oc][m + dl] -------------------------------
    const int dl = warp - 4, et = tid - 128;
    const int oc = et >> 2, jq = et & 3;
    float sK = 0.f, s1 = 0.f, s2 = 0.f;
    int sn = 0;
    float* outc = out + (size_t)b * bs + (size_t)oc * H * W;
    int blk = 0;
#pragma unroll 1
    for (int c = 0; c < T2_NCH; ++c) {
      const int buf = c & 1;
      mbar_wait(&dFull[buf], (uint32_t)((c >> 1) & 1) ^ (buf ? p01 : 0u));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nb = t2_chunk_blocks(c);
#pragma unroll 1
      for (int kb = 0; kb < nb; ++kb, ++blk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(dl * 32) << 16) + (uint32_t)(buf * T2_DBUF + kb * 32 + dl), v);   // lane = oc
        float* pw = sPart + (size_t)(blk & 1) * (T2_PART / 4) + (size_t)(dl * C + lane) * T2_PP;
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(pw + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        if (kb == nb - 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&dEmpty[buf]);
        }
        named_bar_sync(1, 128);
        const float* pr = sPart + (size_t)(blk & 1) * (T2_PART / 4) + (size_t)oc * T2_PP + 8 * jq;
        float o[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 p0 = *reinterpret_cast<const float4*>(pr + 4 * h), p1 = *reinterpret_cast<const float4*>(pr + C * T2_PP + 4 * h);
          const float4 p2 = *reinterpret_cast<const float4*>(pr + 2 * C * T2_PP + 4 * h), p3 = *reinterpret_cast<const float4*>(pr + 3 * C * T2_PP + 4 * h);
          o[4 * h + 0] = (p0.x + p1.x) + (p2.x + p3.x);
          o[4 * h + 1] = (p0.y + p1.y) + (p2.y + p3.y);
          o[4 * h + 2] = (p0.z + p1.z) + (p2.z + p3.z);
          o[4 * h + 3] = (p0.w + p1.w) + (p2.w + p3.w);
        }
        const int m0 = t2_chunk_start(c) + kb * 32 + 8 * jq;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = m0 + 2 * u, r = m / T2_WP, col = m - r * T2_WP;
          if (r < T2_TH && col < T2_TW) {
            *reinterpret_cast<float2*>(outc + (size_t)(y0 + r) * W + x0 + col) = make_float2(o[2 * u], o[2 * u + 1]);
            if (sn == 0) sK = o[2 * u];
            const float d0 = o[2 * u] - sK, d1 = o[2 * u + 1] - sK;
            s1 += d0 + d1;
            s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
            sn += 2;
          }
        }
      }
    }
    // per-thread (n, mean, M2) -> Chan merge over the 4 threads of a channel -> per-CTA partial (mean, M2) of 16x32 outputs
    float n = (float)sn, mean = sn ? sK + s1 / n : 0.f, m2 = sn ? fmaxf(s2 - s1 * s1 / n, 0.f) : 0.f;
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
      const float n2 = __shfl_xor_sync(0xffffffffu, n, off), mean2 = __shfl_xor_sync(0xffffffffu, mean, off);
      const float m22 = __shfl_xor_sync(0xffffffffu, m2, off);
      const float nt = n + n2, dlt = mean2 - mean;
      if (nt > 0.f) {
        m2 = m2 + m22 + dlt * dlt * (n * n2 / nt);
        mean = mean + dlt * (n2 / nt);
      }
      n = nt;
    }
    if (jq == 0) {
      float* dstp = part + ((size_t)cta * C + oc) * 2;
      dstp[0] = mean;
      dstp[1] = m2;
    }
  }
  // every asynchronous operation of this convolution has completed once the epilogue warps are through (they waited for the
  // last accumulator commit, which follows every ring commit); the caller's next barrier releases the region to the point-wise view
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

template <int CELL, int ACT, bool SINGLE>
__global__ void __launch_bounds__(LP_THREADS, 1) int_loop_kernel(IntDev d, LoopArgs la) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sGW = reinterpret_cast<float*>(smem_raw + LP_PW_W);
  float* tiles = reinterpret_cast<float*>(smem_raw + LP_PW_TILES);
  double (*sFin)[C][3] = reinterpret_cast<double (*)[C][3]>(smem_raw + LP_PW_FOLD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + LP_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T2_NST + 6);
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sbi[C], sbe[C], sba[C], sAl[C], sMu[C], sKa[C], sGa[C], sWr[C], sScale[C], sShift[C];

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int H = d.H, W = d.W, P = H * W;
  const int x0 = blockIdx.x * T2_TW, y0 = blockIdx.y * T2_TH, b = blockIdx.z;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int nCta = gridDim.x * gridDim.y * gridDim.z;

  // ---- once per clip: barriers, TMEM, per-channel parameter vectors ------------------------------------------------------
  if (tid == 0) {
    uint64_t* full = bars;
    uint64_t* empty = bars + T2_NST;
    uint64_t* dFull = bars + 2 * T2_NST;
    uint64_t* dEmpty = dFull + 2;
    for (int i = 0; i < T2_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&dFull[0], 1); mbar_init(&dFull[1], 1);
    mbar_init(&dEmpty[0], 4); mbar_init(&dEmpty[1], 4);
    mbar_init(dEmpty + 2, LP_THREADS); mbar_init(dEmpty + 3, LP_THREADS - 64);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < C * 3; i += LP_THREADS) sWp[i] = __ldg(d.Wp + i);
  if (tid < C) {
    const int c = tid;
    sbp[c] = __ldg(d.bp + c);
    sba[c] = __ldg(d.gate_b[0] + c) + __ldg(d.gate_b[1] + c);
    sbi[c] = __ldg(d.gate_b[2] + c) + __ldg(d.gate_b[3] + c);
    sbe[c] = __ldg(d.gate_b[4] + c) + __ldg(d.gate_b[5] + c);
    sAl[c] = __ldg(d.alpha + c); sMu[c] = __ldg(d.mu + c);
    sKa[c] = __ldg(d.kappa + c); sGa[c] = __ldg(d.gamma + c);
    sWr[c] = __ldg(d.Wr + c);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const Lane L;
  // this warp's point-wise tile: image row y0 + warp of the CTA's tile = 32 consecutive pixels
  const int pix_base = (y0 + warp) * W + x0, pix = pix_base + L.pxo;
  const size_t base = (size_t)b * C * P + pix;
  float* sT0 = tiles + warp * 2 * TILE_F;
  float* sT1 = sT0 + TILE_F;
  const float br = __ldg(d.br);
  const MidW mw{sGW, sGW + C * C, sGW + 2 * C * C, sGW + 3 * C * C, sWp, sbp, sbi, sbe, sAl, sMu, sScale, sShift};
  const PostW pw{sGW, sGW + C * C, nullptr, nullptr, sWp, sbp, sba, sbe, sKa, sGa, sScale, sShift, sWr};
  unsigned gtarget = 0, starget = 0;
  unsigned* sctr = la.sctr + b;
  const unsigned tps = (unsigned)la.tiles_per_sample;

#pragma unroll 1
  for (int t = 0; t < d.T; ++t) {
    // ================= conv0: Ge_t -> Y0_t ==================================================================================
    loop_conv<SINGLE>(d.Ge, reinterpret_cast<const uint8_t*>(d.w_inh_f), d.Y0_at(t), la.part[0], cta, H, W, x0, y0, b, smem_raw,
                      tmem_base, 0u);
    cta_arrive(la.gctr);
    gtarget += (unsigned)nCta;
    // ================= mid, statistics-independent half ===================================================================
    {
      cta_fill_async(sGW, d.gateTF[2], C * C);
      cta_fill_async(sGW + C * C, d.gateTF[3], C * C);
      cta_fill_async(sGW + 2 * C * C, d.gateTF[4], C * C);
      cta_fill_async(sGW + 3 * C * C, d.gateTF[5], C * C);
      const float* It = d.I_at(t);
      if (CELL == PTK_CELL_INT) warp_tile_load_async(sT0, It, (size_t)C * P, b, pix_base, P);
      else warp_tile_load_async(sT0, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P);
      warp_tile_load_async(sT1, d.Ge, (size_t)C * P, b, pix_base, P);
      cp_async_commit();
      float xin[3][4];
      frame_px(d.fv, b, t, pix, W, xin);
      cp_async_wait<0>();
      __syncthreads();
      const float* Sg = (CELL == PTK_CELL_INT) ? It + base : d.A_at(t) + (size_t)b * d.a_bstride + pix;
      MidIn cur;
      mid_tile_head<CELL, ACT>(d, t, base, Sg, sT0, sT1, mw, L, xin, cur, It, d.Y0_at(t));
      // ---- BatchNorm0: all convolutions of the batch are in ----
      cta_wait(la.gctr, gtarget);
      loop_fold(la.part[0], nCta, T2_TH * T2_TW, sFin, sScale, sShift, d.bn_w[0], d.bn_b[0], cta == 0 ? d.stat(t, 0) : nullptr);
      mid_tile_tail<CELL, ACT>(d, t, base, Sg, sT0, sT1, mw, L, cur, It, d.Y0_at(t), d.I_at(t + 1));
    }
    // the other tiles of this sample need the halo rows of I_{t+1}
    cta_arrive(sctr);
    starget += tps;
    cta_wait(sctr, starget);
    // ================= conv1: I_{t+1} -> Y1_t ==============================================================================
    loop_conv<SINGLE>(d.I_at(t + 1), reinterpret_cast<const uint8_t*>(d.w_exc_f), d.Y1_at(t), la.part[1], cta, H, W, x0, y0, b,
                      smem_raw, tmem_base, 1u);
    cta_arrive(la.gctr);
    gtarget += (unsigned)nCta;
    // ================= post ===================================================================================================
    {
      const int has_next = (t + 1 < d.T);
      cta_fill_async(sGW, d.gateTF[0], C * C);
      cta_fill_async(sGW + C * C, d.gateTF[1], C * C);
      cp_async_commit();
      float xin[3][4];
      if (has_next) frame_px(d.fv, b, t + 1, pix, W, xin);
      PostIn cur;
      post_load(d.Y1_at(t), d.I_at(t + 1), d.E_at(t), d.Gex_at(t), base + (size_t)L.oc0 * P, cur);
      cp_async_wait<0>();
      // ---- BatchNorm1 ----
      cta_wait(la.gctr, gtarget);
      loop_fold(la.part[1], nCta, T2_TH * T2_TW, sFin, sScale, sShift, d.bn_w[1], d.bn_b[1], cta == 0 ? d.stat(t, 1) : nullptr);
      post_tile<ACT, false>(d, t, has_next, b, pix_base, sT0, sT1, pw, br, L, xin, cur);
      if (has_next) {   // the other tiles of this sample need the halo rows of Ge_{t+1}
        cta_arrive(sctr);
        starget += tps;
        cta_wait(sctr, starget);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// ---- host ------------------------------------------------------------------------------------------------------------------------
size_t int_loop_sync_words(int B) { return (size_t)64 + (size_t)B; }   // grid counter (own 256-byte line) + one counter per sample

template <int CELL, int ACT, bool SINGLE>
static int launch_loop(const IntDev& d, const LoopArgs& la, cudaStream_t st) {
  auto kern = int_loop_kernel<CELL, ACT, SINGLE>;
  PTK_CUDA_OK(ptk_func_smem(kern, LP_SMEM));
  dim3 grid(d.W / T2_TW, d.H / T2_TH, d.B);
  IntDev dd = d;
  LoopArgs ll = la;
  void* args[] = {&dd, &ll};
  ProfileScope ps(PTK_PROF_CONV_FWD, st);
  PTK_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(LP_THREADS), args, (size_t)LP_SMEM, st));
  return PTK_OK;
}

// can the whole grid be co-resident?  (one 215 KB CTA per SM)
bool int_loop_fits(int B, int H, int W) {
  static int sms = -1, coop = 0;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return coop && (long long)B * (H / T2_TH) * (W / T2_TW) <= (long long)sms;
}

int run_int_forward_loop(const IntDev& d, int cell, int act, float* part1, unsigned* sync_words, cudaStream_t st) {
  LoopArgs la;
  la.part[0] = d.bn_part; la.part[1] = part1;
  la.gctr = sync_words; la.sctr = sync_words + 64;
  la.tiles_per_sample = (d.H / T2_TH) * (d.W / T2_TW);
  PTK_CUDA_OK(cudaMemsetAsync(sync_words, 0, int_loop_sync_words(d.B) * sizeof(unsigned), st));
  const bool single = d.conv_mode == PTK_CONV_BF16;
  const bool tanh_ = act == PTK_ACT_TANH;
#define PTK_LOOP_CASE(CELL, ACTV)                                                                      \
  return single ? launch_loop<CELL, ACTV | ACT_FAST, true>(d, la, st) : launch_loop<CELL, ACTV | ACT_FAST, false>(d, la, st)
  if (cell == PTK_CELL_INT) {
    if (tanh_) { PTK_LOOP_CASE(PTK_CELL_INT, PTK_ACT_TANH); }
    PTK_LOOP_CASE(PTK_CELL_INT, PTK_ACT_SOFTPLUS);
  }
  if (tanh_) { PTK_LOOP_CASE(PTK_CELL_HGRU, PTK_ACT_TANH); }
  PTK_LOOP_CASE(PTK_CELL_HGRU, PTK_ACT_SOFTPLUS);
#undef PTK_LOOP_CASE
}

}  // namespace ptk
