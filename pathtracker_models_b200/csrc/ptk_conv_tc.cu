// Tensor-core horizontal-connection convolution: implicit GEMM on tcgen05 / TMEM (sm_100a).
//
//   Y[b, :, y, x] = sum_{tap, ic} In[b, ic, y+ky-3, x+kx-3] * W[:, ic, tap]      (models/InT.py:161,172)
//
// fp32 parity through a 3-pass bf16 split: In = Ih + Il, W = Wh + Wl (bf16 each),
//   Y ~= Ih*Wh + Ih*Wl + Il*Wh, accumulated in fp32 in TMEM.
// Per CTA: 16 rows x 32 columns of one sample (same tiling / arguments as conv7_ffma_kernel).
//   * the zero-padded input tile lives in shared memory as bf16 hi / lo in "channel-chunk planar"
//     form [chunk = ic/8][pixel][8 ic] (pixel pitch 16 B, row pitch 38 pixels), which is exactly the
//     K-major no-swizzle UMMA canonical layout with SBO = 128 B, LBO = plane size -- so every tap
//     of the 7x7 window is the SAME shared-memory image addressed at a different start pixel
//     (ky*38 + kx): no im2col copies.
//   * GEMM shape per MMA: M = 128 consecutive padded pixels (5 M-tiles cover the 16x38 padded tile),
//     K = 16 input channels, N = 64 = [Wh | Wl] for the Ih operand (one instruction yields both
//     Ih*Wh and Ih*Wl) and N = 32 (Wh only) for the Il operand.  The two N-halves are summed in the epilogue.
//   * weights stream per tap (4 KB: hi+lo) through a cp.async.bulk + mbarrier ring; one elected
//     thread issues tcgen05.mma; tcgen05.commit releases ring slots and signals the epilogue.
//   * epilogue: tcgen05.ld -> registers -> BatchNorm batch statistics partials -> global (NCHW fp32).
#include <cuda_bf16.h>

#include "ptk_common.cuh"
#include "ptk_conv_stats.cuh"
#include "ptk_kernels.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int TC_TH = 16, TC_TW = 32;
constexpr int TC_WP = TC_TW + 6;                       // 38: padded row pitch (pixels)
constexpr int TC_ROWS = TC_TH + 6;                     // 22 input rows
constexpr int TC_MT = 5;                               // M-tiles: 5*128 = 640 >= 16*38 = 608
constexpr int TC_NPIX = 880;                           // >= 640 + 6*38 + 6 = 874 pixels incl. slack
constexpr int TC_PLANE = TC_NPIX * 16;                 // bytes per channel-chunk plane
constexpr int TC_A_BYTES = 2 * 4 * TC_PLANE;           // hi + lo, 4 chunks
constexpr int TC_WTAP = 4 * 64 * 16;                   // bytes of one tap's weights: 4 chunks x 64 rows x 8 bf16
constexpr int TC_STAGE = KS * TC_WTAP;                 // one ring stage = one kernel row (7 taps, 28 KB)
constexpr int TC_NST = 3;                              // weight ring stages
constexpr int TC_SMEM = TC_A_BYTES + TC_NST * TC_STAGE + 256;

// accumulator row of one pixel: 32 channels = cols [0,32) + cols [32,64) of M-tile j
__device__ __forceinline__ void tc_load_pixel(uint32_t tmem_base, int quad, int j, float (&v)[32], bool single) {
  float lo[32];
  const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * 64);
  tmem_ld32(ta, v);
  if (single) return;
  tmem_ld32(ta + 32, lo);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] += lo[i];
}
__device__ long long g_tc_dbg[8];
__device__ unsigned long long g_tc_times[1024][3];   // per CTA: start, mma-done, end (globaltimer ns)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_STAMP(i) do { if (blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.x == 64) g_tc_dbg[i] = clock64(); } while (0)

template <bool SINGLE>
__global__ void __launch_bounds__(256, 1) conv7_tc_kernel(ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;                                   // [hi|lo][4 chunks][880 px][16 B]
  uint8_t* sWr = smem_raw + TC_A_BYTES;                     // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + TC_A_BYTES + TC_NST * TC_STAGE);
  uint64_t* full = bars;                                    // [TC_NST]
  uint64_t* empty = bars + TC_NST;                          // [TC_NST]
  uint64_t* accum_bar = bars + 2 * TC_NST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_NST + 1);
  __shared__ float sRed[8][C];
  __shared__ float sMean[C];
  __shared__ double sFin[8][C][3];
  __shared__ int sIsLast;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform => uniform branches / registers
  const int x0 = blockIdx.x * TC_TW, y0 = blockIdx.y * TC_TH, b = blockIdx.z;
  const int H = a.H, W = a.W;

  TC_STAMP(0);
  const int dbg_cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (threadIdx.x == 64 && dbg_cta < 1024) g_tc_times[dbg_cta][0] = gtimer();
  // ---- one-time setup: barriers, TMEM, zero the operand image ------------------------------------
  if (tid == 0) {
    for (int i = 0; i < TC_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < TC_A_BYTES / 16; i += 256) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  TC_STAMP(1);
  // weights do not depend on the input tile: start streaming the first kernel rows right away
  if (tid == 0) {
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wT);
    for (int ky = 0; ky < TC_NST; ++ky) {
      mbar_expect_tx(&full[ky], TC_STAGE);
      bulk_g2s(sWr + ky * TC_STAGE, wsrc + (size_t)ky * TC_STAGE, TC_STAGE, &full[ky]);
    }
  }

  // ---- stage the input tile: fp32 NCHW -> bf16 hi/lo, channel-chunk planar -------------------------
  // interior columns: one item = (chunk, row, 4-pixel group): 8 LDG.128 (one per channel) -> 4 pixels
  for (int it = tid; it < 4 * TC_ROWS * (TC_TW / 4); it += 256) {
    const int q = it / (TC_ROWS * (TC_TW / 4)), rem = it - q * (TC_ROWS * (TC_TW / 4)), r = rem / (TC_TW / 4), g = rem % (TC_TW / 4);
    const int gy = y0 + r - 3;
    if (gy < 0 || gy >= H) continue;
    const float* src = a.in + ((size_t)(b * C + q * 8) * H + gy) * W + x0 + 4 * g;
    float4 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = *reinterpret_cast<const float4*>(src + (size_t)e * H * W);
    const int p = r * TC_WP + 3 + 4 * g;
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
      *reinterpret_cast<uint4*>(sA + q * TC_PLANE + (p + px) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + 4 * TC_PLANE + q * TC_PLANE + (p + px) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  // halo columns (only inside the image when W > 32)
  if (W > TC_TW) {
    for (int it = tid; it < 4 * TC_ROWS * 6; it += 256) {
      const int q = it / (TC_ROWS * 6), rem = it - q * (TC_ROWS * 6), r = rem / 6, hc = rem % 6;
      const int c = hc < 3 ? hc : (TC_TW + hc);          // padded columns 0..2 and 35..37
      const int gy = y0 + r - 3, gx = x0 + c - 3;
      if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
      const float* src = a.in + ((size_t)(b * C + q * 8) * H + gy) * W + gx;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v0 = src[(size_t)(2 * e) * H * W], v1 = src[(size_t)(2 * e + 1) * H * W];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        hi[e] = pack_bf16x2(h0, h1);
        lo[e] = pack_bf16x2(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
      }
      const int p = r * TC_WP + c;
      *reinterpret_cast<uint4*>(sA + q * TC_PLANE + p * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sA + 4 * TC_PLANE + q * TC_PLANE + p * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  TC_STAMP(2);

  // ---- main loop: weight producer (warp 0) / MMA issuer (warp 1) ----------------------------------
  if (warp == 0 && lane == 0) {
    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wT);
    for (int ky = TC_NST; ky < KS; ++ky) {   // refill a slot as soon as the MMAs that read it have retired
      const int st = ky % TC_NST, round = ky / TC_NST;
      mbar_wait(&empty[st], (uint32_t)((round - 1) & 1));
      mbar_expect_tx(&full[st], TC_STAGE);
      bulk_g2s(sWr + st * TC_STAGE, wsrc + (size_t)ky * TC_STAGE, TC_STAGE, &full[st]);
    }
  } else if (warp == 1) {
    // descriptors differ only in the 14-bit start-address field of the low word: precompute the
    // per-(operand, k-step) bases and add the pixel offset per instruction (1 IADD per MMA).
    const uint32_t aHi = smem_u32(sA), aLo = aHi + 4 * TC_PLANE, wBase = smem_u32(sWr);
    const uint64_t dA0 = umma_desc(0, TC_PLANE, 128), dB0 = umma_desc(0, 1024, 128);
    const uint32_t aHiW = (uint32_t)(dA0 >> 32), bHiW = (uint32_t)(dB0 >> 32);
    const uint32_t aLoW = (uint32_t)dA0, bLoW = (uint32_t)dB0;
    uint32_t aH[2], aL[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      aH[s] = aLoW + ((aHi + (uint32_t)(2 * s) * TC_PLANE) >> 4);
      aL[s] = aLoW + ((aLo + (uint32_t)(2 * s) * TC_PLANE) >> 4);
    }
    constexpr uint32_t ID64 = umma_idesc(64), ID32 = umma_idesc(32);
    for (int ky = 0; ky < KS; ++ky) {
      const int st = ky % TC_NST, round = ky / TC_NST;
      mbar_wait(&full[st], (uint32_t)(round & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int kx = 0; kx < KS; ++kx) {
        const uint32_t poff = (uint32_t)(ky * TC_WP + kx);          // in 16-byte units == pixels
        const uint32_t bW = bLoW + ((wBase + (uint32_t)st * TC_STAGE + (uint32_t)kx * TC_WTAP) >> 4);
#pragma unroll
        for (int j = 0; j < TC_MT; ++j) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const uint32_t po = poff + (uint32_t)(j * 128);
            const uint64_t dB = ((uint64_t)bHiW << 32) | (bW + (uint32_t)(2 * s) * (1024 >> 4));
            const uint32_t dcol = tmem_base + (uint32_t)(j * 64);
            if (elect_one()) {
              if (SINGLE) {   // PTK_CONV_BF16: Ih*Wh only
                umma_bf16(dcol, ((uint64_t)aHiW << 32) | (aH[s] + po), dB, ID32, (ky | kx | s) ? 1u : 0u);
              } else {
                umma_bf16(dcol, ((uint64_t)aHiW << 32) | (aH[s] + po), dB, ID64, (ky | kx | s) ? 1u : 0u);
                umma_bf16(dcol, ((uint64_t)aHiW << 32) | (aL[s] + po), dB, ID32, 1u);
              }
            }
          }
        }
      }
      __syncwarp();
      if (elect_one()) umma_commit(&empty[st]);   // frees this ring slot once the MMAs above have read it
    }
    if (elect_one()) umma_commit(accum_bar);
  }
  __syncwarp();

  // ---- epilogue -----------------------------------------------------------------------------------
  mbar_wait(accum_bar, 0);
  TC_STAMP(3);
  if (threadIdx.x == 64 && dbg_cta < 1024) g_tc_times[dbg_cta][1] = gtimer();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int quad = warp & 3, jpar = warp >> 2;   // warps 0-3: tiles 0,2,4; warps 4-7: tiles 1,3
  const int nCta = gridDim.x * gridDim.y * gridDim.z;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const bool stats = (a.part != nullptr);
  if (stats) {
    float s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = 0.f;
    for (int j = jpar; j < TC_MT; j += 2) {
      const int m = j * 128 + quad * 32 + lane, ro = m / TC_WP, xo = m - ro * TC_WP;
      float v[32];
      tc_load_pixel(tmem_base, quad, j, v, SINGLE);
      if (ro < TC_TH && xo < TC_TW) {
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] += v[i];
      }
    }
    const float tot = warp_transpose_sum(s, lane);
    sRed[warp][lane] = tot;
    __syncthreads();
    if (tid < C) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sRed[w][tid];
      sMean[tid] = t * (1.0f / (TC_TH * TC_TW));
    }
    __syncthreads();
  }
  {
    float s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = 0.f;
    for (int j = jpar; j < TC_MT; j += 2) {
      const int m = j * 128 + quad * 32 + lane, ro = m / TC_WP, xo = m - ro * TC_WP;
      float v[32];
      tc_load_pixel(tmem_base, quad, j, v, SINGLE);
      if (ro < TC_TH && xo < TC_TW) {
        float* dst = a.out + ((size_t)(b * C) * H + y0 + ro) * W + x0 + xo;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (stats) {
            const float dl = v[i] - sMean[i];
            s[i] = fmaf(dl, dl, s[i]);
          }
          float* p = dst + (size_t)i * H * W;
          *p = a.accumulate ? (*p + v[i]) : v[i];
        }
      }
    }
    if (stats) {
      const float tot = warp_transpose_sum(s, lane);
      __syncthreads();
      sRed[warp][lane] = tot;
      __syncthreads();
      if (tid < C) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sRed[w][tid];
        float* dstp = a.part + ((size_t)cta * C + tid) * 2;
        dstp[0] = sMean[tid];
        dstp[1] = t;
      }
      conv_stats_fold(a, nCta, TC_TH * TC_TW, sFin, &sIsLast);
    }
  }
  TC_STAMP(4);
  if (threadIdx.x == 64 && dbg_cta < 1024) g_tc_times[dbg_cta][2] = gtimer();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// weights fp32 [oc][ic][49] -> bf16 hi/lo per tap: [tap][chunk = k/8][row n: 0-31 hi, 32-63 lo][k%8]
//   forward: n = oc, k = ic, tap as is;  dgrad (conv_transpose): n = ic, k = oc, tap -> 48 - tap.
__global__ void prep_weights_tc_kernel(const float* __restrict__ w, __nv_bfloat16* wf, __nv_bfloat16* wd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C * TAPS) return;
  const int oc = i / (C * TAPS), ic = (i / TAPS) % C, tap = i % TAPS;
  const float v = w[i];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  if (wf) {
    const size_t base = (size_t)tap * (TC_WTAP / 2) + (size_t)(ic >> 3) * 512 + (ic & 7);
    wf[base + (size_t)oc * 8] = h;
    wf[base + (size_t)(32 + oc) * 8] = l;
  }
  if (wd) {
    const size_t base = (size_t)(TAPS - 1 - tap) * (TC_WTAP / 2) + (size_t)(oc >> 3) * 512 + (oc & 7);
    wd[base + (size_t)ic * 8] = h;
    wd[base + (size_t)(32 + ic) * 8] = l;
  }
}

int launch_conv7_tc(const ConvArgs& a, cudaStream_t st) {
  dim3 grid(a.W / TC_TW, a.H / TC_TH, a.B);
  if (a.mode == PTK_CONV_BF16) {
    PTK_CUDA_OK(cudaFuncSetAttribute(conv7_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    conv7_tc_kernel<true><<<grid, 256, TC_SMEM, st>>>(a);
  } else {
    PTK_CUDA_OK(cudaFuncSetAttribute(conv7_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    conv7_tc_kernel<false><<<grid, 256, TC_SMEM, st>>>(a);
  }
  PTK_LAUNCH_OK("conv7_tc_kernel");
  return PTK_OK;
}

int read_tc_times(unsigned long long* out, int n) {
  PTK_CUDA_OK(cudaMemcpyFromSymbol(out, g_tc_times, sizeof(unsigned long long) * 3 * n));
  return PTK_OK;
}

int read_tc_debug(long long* out) {
  PTK_CUDA_OK(cudaMemcpyFromSymbol(out, g_tc_dbg, sizeof(long long) * 8));
  return PTK_OK;
}

int launch_prep_weights_tc(const float* w, void* wf, void* wd, cudaStream_t st) {
  const int n = C * C * TAPS;
  prep_weights_tc_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd);
  PTK_LAUNCH_OK("prep_weights_tc_kernel");
  return PTK_OK;
}

}  // namespace ptk
