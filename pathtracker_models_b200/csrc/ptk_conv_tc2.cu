// Tensor-core horizontal-connection convolution, second generation: weights are the A operand, pixels the B operand.
//
//   Y[b, oc, y, x] = sum_{tap, ic} In[b, ic, y+ky-3, x+kx-3] * W[oc, ic, tap]      (models/InT.py:161,172; dgrad = flipped W)
//
// What changed against ptk_conv_tc.cu (pixels as A with M = 128, three taps stacked along N = 96) and why:
//  * that kernel's MMA loop ran at the shared-memory operand floor (7 KB of operands per 48 tensor-pipe cycles: 68 cycles
//    per instruction measured, tensor pipe 42 % active), then drained all of TMEM in a serial epilogue.
//  * here an instruction is M = 128 = 4 horizontally adjacent taps x 32 output channels (A = weights, 4 KB, K-major) times
//    N = 240 or 160 padded pixels (B = the SAME zero-padded bf16 image as before, addressed at start pixel
//    chunk + ky*38 + 4*kxg): 11.5 KB of operands per 120 tensor-pipe cycles = 96 B/clk, under the ~105-128 B/clk the
//    shared memory delivers, so the loop is bound by the tensor pipe.  3-pass split as before (Wh*Ih + Wh*Il + Wl*Ih).
//  * the 16x38 padded output tile is produced in three row chunks (6, 6, 4 rows) whose accumulators alternate between two
//    240-column TMEM buffers: four epilogue warps drain chunk c while the MMA warp issues chunk c+1.
//  * D[(dl, oc)][n] holds tap kx = 4*kxg + dl's contribution to output pixel n - dl: the realignment is a column offset of
//    the tcgen05.ld, the sum over the four lane quadrants goes through a small shared-memory tile, and because lanes are
//    output CHANNELS the result is already row-major NCHW: coalesced stores, per-channel BatchNorm partial statistics in
//    registers.
//  * price: the weight stream (14 stages of 16 KB) passes once per chunk, 3 x 224 KB per CTA, through a 4-stage
//    cp.async.bulk ring.
#include <cuda_bf16.h>

#include "ptk_common.cuh"
#include "ptk_conv_stats.cuh"
#include "ptk_conv_tc2.cuh"
#include "ptk_kernels.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int T2_THREADS = 384;   // warp 0: weight producer, 1: MMA issuer, 2-3: staging only, 4-11: two epilogue groups; all warps stage

template <bool SINGLE>
__global__ void __launch_bounds__(T2_THREADS, 1) conv7_tc2_kernel(ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;                                   // [hi|lo][4 chunks][880 px][16 B]
  uint8_t* sWr = smem_raw + T2_A_BYTES;                     // weight ring
  float* sPart = reinterpret_cast<float*>(smem_raw + T2_A_BYTES + T2_NST * T2_STAGE);   // [2][4][32][36]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + T2_A_BYTES + T2_NST * T2_STAGE + 2 * T2_PART);
  uint64_t* full = bars;                                    // [T2_NST]
  uint64_t* empty = bars + T2_NST;                          // [T2_NST]
  uint64_t* dFull = bars + 2 * T2_NST;                      // [2] accumulator buffer complete
  uint64_t* dEmpty = dFull + 2;                             // [2] accumulator buffer drained
  uint64_t* stA = dEmpty + 2;                               // operand image: input rows 0-11 (+ halo columns) staged
  uint64_t* stB = stA + 1;                                  // input rows 12-21 staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stB + 1);
  __shared__ double sFin[8][C][3];
  __shared__ float sMerge[C * 4];                           // BatchNorm partials of the second epilogue group
  __shared__ float4 sBN[C];                                 // fused BatchNorm backward: k, mean(dV), mean, rstd * mean(dV*xhat)
  __shared__ int sIsLast;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int x0 = blockIdx.x * T2_TW, y0 = blockIdx.y * T2_TH, b = blockIdx.z;
  const int H = a.H, W = a.W;
  const size_t in_bs = a.in_bs ? (size_t)a.in_bs : (size_t)C * H * W, out_bs = a.out_bs ? (size_t)a.out_bs : (size_t)C * H * W;
  constexpr int NSTG = T2_NCH * T2_NG;                      // 42 weight stages per CTA

  // ---- one-time setup: barriers, TMEM, zero the operand image ------------------------------------
  if (tid == 0) {
    for (int i = 0; i < T2_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&dFull[0], 1); mbar_init(&dFull[1], 1);
    mbar_init(&dEmpty[0], T2_EPI_WARPS); mbar_init(&dEmpty[1], T2_EPI_WARPS);
    mbar_init(stA, T2_THREADS); mbar_init(stB, T2_THREADS - 64);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < T2_A_BYTES / 16; i += T2_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // weights do not depend on the input tile: start streaming right away
  if (tid == 0) {
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wT);
    for (int i = 0; i < T2_NST; ++i) {
      mbar_expect_tx(&full[i], T2_STAGE);
      bulk_g2s(sWr + i * T2_STAGE, wsrc + (size_t)i * T2_STAGE, T2_STAGE, &full[i]);
    }
  }
  if (a.bn_y != nullptr) {   // the saved conv output is a forward checkpoint: pull the tile's lines HBM -> L2 before the wait
    for (int it = tid; it < C * T2_ROWS; it += T2_THREADS) {
      const int ch = it / T2_ROWS, gy = y0 - 3 + (it - ch * T2_ROWS);
      if (gy >= 0 && gy < H) prefetch_l2(a.bn_y + (size_t)b * C * H * W + ((size_t)ch * H + gy) * W + x0);
    }
  }
  // everything above touched only shared memory, TMEM and the prepared weights (written >= 2 kernels ago)
  pdl_wait();
  pdl_launch_dependents();

  // ---- stage the input tile: fp32 NCHW -> bf16 hi/lo, channel-chunk planar (same image as ptk_conv_tc.cu) -------
  // With bn_y set the tile is dV and the BatchNorm backward (b3 / b9) is applied on the way: per channel
  //   dY = k * (dV - m - (Y - mean) * rn),  k = w*rstd, m = mean(dV), rn = rstd * mean(dV*xhat)
  const bool bnf = (a.bn_y != nullptr);
  if (bnf) {
    if (tid < C) {
      const float rstd = a.bn_stat[C + tid];
      sBN[tid] = make_float4(__ldg(a.bn_w + tid) * rstd, a.bn_tot[tid], a.bn_stat[tid], rstd * a.bn_tot[C + tid]);
    }
    __syncthreads();
  }
  // Two phases, so that the MMA warp can start the first row chunk (input rows 0-11) while rows 12-21 are still being staged:
  //   phase A: rows 0-11 (items 0..383), one item per thread, + the halo columns  -> mbarrier stA
  //   phase B: rows 12-21 (items 384..703), warps 2-11 only (warps 0 / 1 go to their producer / MMA roles) -> mbarrier stB
  // one item = (chunk, row, 4-pixel group): 8 LDG.128 (one per channel) -> 4 pixels (the tile is L2-latency bound: one item per
  // thread and phase, 12 warps' worth of loads in flight).
  auto stage_item = [&](int it, int limit) {
    // a quarter warp = (2 adjacent pixel groups) x (4 chunks): conflict-free 16-byte stores; a warp covers one whole input row
    const int q = (it >> 1) & 3, row = it >> 5, g = ((it >> 3) & 3) * 2 + (it & 1);
    const int gy = y0 + row - 3;
    if (it >= limit || gy < 0 || gy >= H) return;
    const int p = row * T2_WP + 3 + 4 * g;
    const size_t off = ((size_t)(q * 8) * H + gy) * W + x0 + 4 * g;
    const float* src = a.in + (size_t)b * in_bs + off;
    float4 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = *reinterpret_cast<const float4*>(src + (size_t)e * H * W);
    if (bnf) {
      const float* ys = a.bn_y + (size_t)b * C * H * W + off;
      float4 yv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) yv[e] = *reinterpret_cast<const float4*>(ys + (size_t)e * H * W);
      const bool own = row >= 3 && row < 3 + T2_TH;        // interior rows: this CTA owns their dY
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 cst = sBN[q * 8 + e], y = yv[e];
        float4 t = v[e];
        t.x = cst.x * (t.x - cst.y - (y.x - cst.z) * cst.w);
        t.y = cst.x * (t.y - cst.y - (y.y - cst.z) * cst.w);
        t.z = cst.x * (t.z - cst.y - (y.z - cst.z) * cst.w);
        t.w = cst.x * (t.w - cst.y - (y.w - cst.z) * cst.w);
        v[e] = t;
        if (own) *reinterpret_cast<float4*>(a.bn_out + (size_t)b * C * H * W + off + (size_t)e * H * W) = t;
      }
    }
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
  stage_item(tid, NIT_A);
  if (W > T2_TW) {   // halo columns (only inside the image when W > 32)
    for (int it = tid; it < 4 * T2_ROWS * 6; it += T2_THREADS) {
      const int q = it / (T2_ROWS * 6), rem = it - q * (T2_ROWS * 6), r = rem / 6, hc = rem % 6;
      const int c = hc < 3 ? hc : (T2_TW + hc);          // padded columns 0..2 and 35..37
      const int gy = y0 + r - 3, gx = x0 + c - 3;
      if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
      const size_t o = ((size_t)(q * 8) * H + gy) * W + gx;
      const float* src = a.in + (size_t)b * in_bs + o;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v0 = src[(size_t)(2 * e) * H * W], v1 = src[(size_t)(2 * e + 1) * H * W];
        if (bnf) {
          const float* ys = a.bn_y + (size_t)b * C * H * W + o;
          const float4 c0 = sBN[q * 8 + 2 * e], c1 = sBN[q * 8 + 2 * e + 1];
          v0 = c0.x * (v0 - c0.y - (ys[(size_t)(2 * e) * H * W] - c0.z) * c0.w);
          v1 = c1.x * (v1 - c1.y - (ys[(size_t)(2 * e + 1) * H * W] - c1.z) * c1.w);
        }
        split_bf16x2(v0, v1, hi[e], lo[e]);
      }
      const int p = r * T2_WP + c;
      *reinterpret_cast<uint4*>(sA + q * T2_PLANE + p * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + 4 * T2_PLANE + q * T2_PLANE + p * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
  mbar_arrive(stA);
  if (warp >= 2) {
    stage_item(NIT_A + tid - 64, NIT);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(stB);
  }
  const bool stats = (a.part != nullptr);

  if (warp == 0) {
    // ---- weight producer: refill a ring slot as soon as the MMAs that read it have retired --------------------
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wT);
      for (int i = T2_NST; i < NSTG; ++i) {
        const int st = i % T2_NST, round = i / T2_NST;
        mbar_wait(&empty[st], (uint32_t)((round - 1) & 1));
        mbar_expect_tx(&full[st], T2_STAGE);
        bulk_g2s(sWr + st * T2_STAGE, wsrc + (size_t)(i % T2_NG) * T2_STAGE, T2_STAGE, &full[st]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer -------------------------------------------------------------------------------------------
    const uint32_t bHiA = smem_u32(sA), bLoA = bHiA + 4 * T2_PLANE, wBase = smem_u32(sWr);
    const uint64_t dB0 = umma_desc(0, T2_PLANE, 128);          // pixels: K-major, LBO = next 8-channel plane, SBO = next 8 pixels
    const uint64_t dW0 = umma_desc(0, 128 * 16, 128);          // weights: K-major, LBO = next k-chunk (128 rows x 16 B), SBO = next 8 rows
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
      mbar_wait(c == 0 ? stA : stB, 0u);                      // chunk 0 reads input rows 0-11, chunks 1 / 2 rows 6-21
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (c >= 2) {                                            // the buffer's previous chunk must have been drained
        mbar_wait(&dEmpty[buf], 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t idesc = umma_idesc(t2_chunk_n(c));
      const uint32_t dcol = tmem_base + (uint32_t)(buf * T2_DBUF);
#pragma unroll 1
      for (int g = 0; g < T2_NG; ++g, ++i) {
        const int st = i % T2_NST, round = i / T2_NST;
        mbar_wait(&full[st], (uint32_t)(round & 1));
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
            if (!SINGLE) {   // PTK_CONV_BF16 keeps Wh*Ih only
              umma_bf16(dcol, dWh, dIl, idesc, 1u);
              umma_bf16(dcol, dWl, dIh, idesc, 1u);
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty[st]);   // frees this ring slot once the MMAs above have read it
      }
      __syncwarp();
      if (elect_one()) umma_commit(&dFull[buf]);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ---- epilogue: two groups of four warps drain alternate 32-pixel blocks (t2_epilogue, ptk_conv_tc2.cuh) --------------------
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const T2Epi e{a.out + (size_t)b * out_bs, H, W, x0, y0, a.accumulate, stats ? a.part + (size_t)cta * C * 2 : nullptr};
    t2_epilogue(e, tmem_base, sPart, sMerge, dFull, dEmpty, 0u, warp - 4, lane);
  }
  if (stats && !a.defer_stats) conv_stats_fold(a, gridDim.x * gridDim.y * gridDim.z, T2_TH * T2_TW, sFin, &sIsLast);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// weights fp32 [oc][ic][49] -> bf16 hi/lo in the staged layout (t2_woff); the dummy tap kx = 7 is written as zeros.
//   forward: n = oc, k = ic, tap as is;  dgrad (conv_transpose): n = ic, k = oc, tap -> 48 - tap.
__global__ void prep_weights_tc2_kernel(const float* __restrict__ w, __nv_bfloat16* wf, __nv_bfloat16* wd, int oc_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C * KS * 8) return;
  const int oc = i / (C * KS * 8), ic = (i / (KS * 8)) % C, ky = (i / 8) % KS, kx = i % 8;
  const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
  if (kx == KS) {
    if (wf) { wf[t2_woff(ky, kx, oc, ic, false) / 2] = z; wf[t2_woff(ky, kx, oc, ic, true) / 2] = z; }
    if (wd) { wd[t2_woff(ky, kx, ic, oc, false) / 2] = z; wd[t2_woff(ky, kx, ic, oc, true) / 2] = z; }
    return;
  }
  const float v = w[(size_t)oc * oc_stride + ic * TAPS + ky * KS + kx];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  if (wf) {
    wf[t2_woff(ky, kx, oc, ic, false) / 2] = h;
    wf[t2_woff(ky, kx, oc, ic, true) / 2] = l;
  }
  if (wd) {
    wd[t2_woff(KS - 1 - ky, KS - 1 - kx, ic, oc, false) / 2] = h;
    wd[t2_woff(KS - 1 - ky, KS - 1 - kx, ic, oc, true) / 2] = l;
  }
}

size_t conv7_tc2_weight_bytes() { return (size_t)T2_NG * T2_STAGE; }

int launch_conv7_tc2(const ConvArgs& a, cudaStream_t st) {
  dim3 grid(a.W / T2_TW, a.H / T2_TH, a.B);
  if (a.mode == PTK_CONV_BF16) {
    PTK_CUDA_OK(ptk_func_smem(conv7_tc2_kernel<true>, T2_SMEM));
    PTK_CUDA_OK(launch_pdl(a.pdl != 0, conv7_tc2_kernel<true>, grid, dim3(T2_THREADS), T2_SMEM, st, a));
  } else {
    PTK_CUDA_OK(ptk_func_smem(conv7_tc2_kernel<false>, T2_SMEM));
    PTK_CUDA_OK(launch_pdl(a.pdl != 0, conv7_tc2_kernel<false>, grid, dim3(T2_THREADS), T2_SMEM, st, a));
  }
  PTK_LAUNCH_OK("conv7_tc2_kernel");
  return PTK_OK;
}

int launch_prep_weights_tc2(const float* w, void* wf, void* wd, cudaStream_t st, int oc_stride) {
  const int n = C * C * KS * 8;
  prep_weights_tc2_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, oc_stride ? oc_stride : C * TAPS);
  PTK_LAUNCH_OK("prep_weights_tc2_kernel");
  return PTK_OK;
}

}  // namespace ptk
