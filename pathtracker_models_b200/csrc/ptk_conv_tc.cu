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
//     K = 16 input channels, N = 96 = three horizontally adjacent taps x 32 output channels ("tap
//     stacking").  The MMA phase is bound by shared-memory operand bandwidth (~128 B/clk/SM: an M=128,
//     K=16 A tile is 4 KB however few columns it feeds), so one A read must feed as many columns as
//     TMEM allows: with the A operand at start pixel ky*38 + kx0, column block dl holds
//     sum In[p + (ky,kx0)] * W[ky, kx0+dl], i.e. the contribution to output pixel p - dl.  The blocks
//     are realigned in the epilogue (out[m] = D0[m] + D1[m+1] + D2[m+2], a lane shift).  5 tiles x 96
//     columns = 480 of the 512 TMEM columns.  Taps per kernel row: groups {0,1,2}, {3,4,5}, {6} (N = 32).
//     Per group three instructions: Ih*Wh, Ih*Wl, Il*Wh, all accumulating into the same columns.
//   * weights stream per kernel row (28 KB: hi+lo) through a cp.async.bulk + mbarrier ring; one elected
//     thread issues tcgen05.mma; tcgen05.commit releases ring slots and signals the epilogue.
//   * epilogue: tcgen05.ld -> registers -> realign -> BatchNorm batch statistics partials -> global.
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
constexpr int TC_PLANE = TC_NPIX * 16 + 16;            // bytes per channel-chunk plane; +16 so the 4 chunks of one pixel hit 4 different bank groups
constexpr int TC_A_BYTES = 2 * 4 * TC_PLANE;           // hi + lo, 4 chunks
constexpr int TC_WTAP = 4 * 64 * 16;                   // bytes of one tap's weights: 4 chunks x (32 hi + 32 lo rows) x 8 bf16
constexpr int TC_STAGE = KS * TC_WTAP;                 // one ring stage = one kernel row (7 taps, 28 KB)
constexpr int TC_SP = TC_TH * TC_WP + 4;                // row pitch (floats) of the realigned output staging tile
constexpr int TC_DCOLS = 96;                           // TMEM columns per M-tile: 3 stacked taps x 32 channels
constexpr int TC_GRP_BYTES = 3 * TC_WTAP;              // a 3-tap group: [Wh: 4 chunks x 96 rows x 16 B][Wl: same]
// byte offset of weight element (row n, k) of tap (ky, kx) inside the staged image
__host__ __device__ constexpr size_t tc_woff(int ky, int kx, int n, int k, bool lo) {
  const int gi = kx / 3, dl = kx - 3 * gi, N = gi < 2 ? 96 : 32;
  return (size_t)ky * TC_STAGE + (size_t)gi * TC_GRP_BYTES + (lo ? 64 * N : 0) + (size_t)(k >> 3) * (N * 16) +
         (size_t)(dl * 32 + n) * 16 + (size_t)(k & 7) * 2;
}
constexpr int TC_NST = 3;                              // weight ring stages
constexpr int TC_SMEM = TC_A_BYTES + TC_NST * TC_STAGE + 256;

// Tuning instrumentation (tools/step_stamps.py, gap_probe.py, conv_cta_times.py): compiled in only with -DPTK_DEBUG_STAMPS=1
// (PTK_DEBUG_STAMPS=1 python -m pathtracker_models_b200._build --force); the shipped library carries none of it.
#if PTK_DEBUG_STAMPS
constexpr int TC_TIMES_MAX = 1024;
__device__ long long g_tc_dbg[8];
__device__ unsigned long long g_tc_times[TC_TIMES_MAX][3];   // per CTA: start, mma-done, end (globaltimer ns)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_STAMP(i) do { if (blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.x == 64) g_tc_dbg[i] = clock64(); } while (0)
#define TC_TIME(cta, i) do { if (threadIdx.x == 64 && (cta) < TC_TIMES_MAX) g_tc_times[cta][i] = gtimer(); } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#define TC_TIME(cta, i) do { } while (0)
#endif

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
  __shared__ double sFin[8][C][3];
  __shared__ int sIsLast;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform => uniform branches / registers
  const int x0 = blockIdx.x * TC_TW, y0 = blockIdx.y * TC_TH, b = blockIdx.z;
  const int H = a.H, W = a.W;
  const size_t in_bs = a.in_bs ? (size_t)a.in_bs : (size_t)C * H * W, out_bs = a.out_bs ? (size_t)a.out_bs : (size_t)C * H * W;

  TC_STAMP(0);
#if PTK_DEBUG_STAMPS
  const int dbg_cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
#endif
  TC_TIME(dbg_cta, 0);
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
  // everything above touched only shared memory, TMEM and the prepared weights (written >= 2 kernels ago)
  pdl_wait();
  pdl_launch_dependents();

  // ---- stage the input tile: fp32 NCHW -> bf16 hi/lo, channel-chunk planar -------------------------
  // interior columns: one item = (chunk, row, 4-pixel group): 8 LDG.128 (one per channel) -> 4 pixels.
  // All of a thread's items (<= 3) are requested before the first conversion: the tile is L2-latency bound.
  {
    constexpr int NIT = 4 * TC_ROWS * (TC_TW / 4), NR = (NIT + 255) / 256;
    float4 v[NR][8];
    int pp[NR], qq[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      // a quarter warp = (2 adjacent pixel groups) x (4 chunks): its eight 16-byte stores land 64*g1 + 16*q bytes apart
      // modulo the 128-byte bank period, i.e. conflict-free (one row per chunk-plane per quarter warp was 4-way conflicted);
      // a warp covers one whole input row of all 32 channels, so the loads are still full 128-byte lines.
      const int it = tid + 256 * r;
      const int q = (it >> 1) & 3, row = it >> 5, g = ((it >> 3) & 3) * 2 + (it & 1);
      const int gy = y0 + row - 3;
      pp[r] = -1; qq[r] = q;
      if (it < NIT && gy >= 0 && gy < H) {
        pp[r] = row * TC_WP + 3 + 4 * g;
        const float* src = a.in + (size_t)b * in_bs + ((size_t)(q * 8) * H + gy) * W + x0 + 4 * g;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[r][e] = *reinterpret_cast<const float4*>(src + (size_t)e * H * W);
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (pp[r] < 0) continue;
      const int p = pp[r], q = qq[r];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = reinterpret_cast<const float*>(&v[r][2 * e])[px], v1 = reinterpret_cast<const float*>(&v[r][2 * e + 1])[px];
          split_bf16x2(v0, v1, hi[e], lo[e]);
        }
        *reinterpret_cast<uint4*>(sA + q * TC_PLANE + (p + px) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sA + 4 * TC_PLANE + q * TC_PLANE + (p + px) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  // halo columns (only inside the image when W > 32)
  if (W > TC_TW) {
    for (int it = tid; it < 4 * TC_ROWS * 6; it += 256) {
      const int q = it / (TC_ROWS * 6), rem = it - q * (TC_ROWS * 6), r = rem / 6, hc = rem % 6;
      const int c = hc < 3 ? hc : (TC_TW + hc);          // padded columns 0..2 and 35..37
      const int gy = y0 + r - 3, gx = x0 + c - 3;
      if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
      const float* src = a.in + (size_t)b * in_bs + ((size_t)(q * 8) * H + gy) * W + gx;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v0 = src[(size_t)(2 * e) * H * W], v1 = src[(size_t)(2 * e + 1) * H * W];
        split_bf16x2(v0, v1, hi[e], lo[e]);
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
    const uint64_t dA0 = umma_desc(0, TC_PLANE, 128);
    const uint64_t dB96 = umma_desc(0, 96 * 16, 128), dB32 = umma_desc(0, 32 * 16, 128);   // LBO = rows x 16 B between k-chunks
    const uint32_t aHiW = (uint32_t)(dA0 >> 32), aLoW = (uint32_t)dA0;
    uint32_t aH[2], aL[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      aH[s] = aLoW + ((aHi + (uint32_t)(2 * s) * TC_PLANE) >> 4);
      aL[s] = aLoW + ((aLo + (uint32_t)(2 * s) * TC_PLANE) >> 4);
    }
    constexpr uint32_t ID96 = umma_idesc(96), ID32 = umma_idesc(32);
    for (int ky = 0; ky < KS; ++ky) {
      const int st = ky % TC_NST, round = ky / TC_NST;
      mbar_wait(&full[st], (uint32_t)(round & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t stBase = wBase + (uint32_t)st * TC_STAGE;
#pragma unroll
      for (int gi = 0; gi < 3; ++gi) {
        const int N = gi < 2 ? 96 : 32;
        const uint32_t idesc = gi < 2 ? ID96 : ID32;
        const uint64_t dB0 = gi < 2 ? dB96 : dB32;
        const uint32_t bHiW = (uint32_t)(dB0 >> 32);
        const uint32_t bH = (uint32_t)dB0 + ((stBase + (uint32_t)gi * TC_GRP_BYTES) >> 4);   // Wh block of the group
        const uint32_t bL = bH + (uint32_t)((64 * N) >> 4);                                  // Wl block
        const uint32_t poff = (uint32_t)(ky * TC_WP + 3 * gi);                               // in 16-byte units == pixels
#pragma unroll
        for (int j = 0; j < TC_MT; ++j) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const uint32_t po = poff + (uint32_t)(j * 128);
            const uint32_t kadv = (uint32_t)(s * ((2 * N * 16) >> 4));
            const uint64_t dBh = ((uint64_t)bHiW << 32) | (bH + kadv), dBl = ((uint64_t)bHiW << 32) | (bL + kadv);
            const uint32_t dcol = tmem_base + (uint32_t)(j * TC_DCOLS);
            if (elect_one()) {
              umma_bf16(dcol, ((uint64_t)aHiW << 32) | (aH[s] + po), dBh, idesc, (ky | gi | s) ? 1u : 0u);
              if (!SINGLE) {   // PTK_CONV_BF16 keeps Ih*Wh only
                umma_bf16(dcol, ((uint64_t)aHiW << 32) | (aH[s] + po), dBl, idesc, 1u);
                umma_bf16(dcol, ((uint64_t)aHiW << 32) | (aL[s] + po), dBh, idesc, 1u);
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
  TC_TIME(dbg_cta, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int quad = warp & 3, jpar = warp >> 2;   // warps 0-3: tiles 0,2,4; warps 4-7: tiles 1,3
  const int nCta = gridDim.x * gridDim.y * gridDim.z;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const bool stats = (a.part != nullptr);
  // Realign the stacked-tap column blocks: out[m] = D0[m] + D1[m+1] + D2[m+2] (m = TMEM lane = padded
  // pixel).  Inside a 32-lane quadrant that is a shuffle; the two rows past a quadrant are handed over
  // through a small edge buffer.  The realigned tile is staged in shared memory (the operand image is
  // dead once accum_bar fired) as [channel][padded pixel], from which the output pass reads whole rows:
  // compact rolled loops (an unrolled register-resident version spent 3x longer fetching instructions).
  float* sS = reinterpret_cast<float*>(sA);      // [32][TC_SP]
  float* sEdge = sS + C * TC_SP;                 // [20 quadrants][3][32]: D1 of lane 0, D2 of lane 0, D2 of lane 1
#pragma unroll 1
  for (int j = jpar; j < TC_MT; j += 2) {
    float o[32], v1[32], v2[32];
    const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * TC_DCOLS);
    tmem_ld32(ta, o);
    tmem_ld32(ta + 32, v1);
    tmem_ld32(ta + 64, v2);
    float* e = sEdge + (size_t)((j * 4 + quad) * 3) * 32;
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) { e[i] = v1[i]; e[32 + i] = v2[i]; }
    } else if (lane == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) e[64 + i] = v2[i];
    }
    const int m = j * 128 + quad * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float t1 = __shfl_down_sync(0xffffffffu, v1[i], 1), t2 = __shfl_down_sync(0xffffffffu, v2[i], 2);
      const float r = o[i] + ((lane < 31 ? t1 : 0.f) + (lane < 30 ? t2 : 0.f));
      if (m < TC_TH * TC_WP) sS[i * TC_SP + m] = r;
    }
  }
  __syncthreads();
  // rows 30 / 31 of every quadrant take the missing terms from the next quadrant's edge rows
  for (int e = tid; e < (TC_TH * TC_WP / 32) * 2 * C; e += 256) {
    const int q = e >> 6, r = (e >> 5) & 1, c = e & 31, m = q * 32 + 30 + r;
    const float* E = sEdge + (size_t)(q + 1) * 96;
    if (m < TC_TH * TC_WP) sS[c * TC_SP + m] += r ? (E[c] + E[64 + c]) : E[32 + c];
  }
  __syncthreads();
  // output pass: warp w owns channels 4w..4w+3; lane = (row, half row) -> 16 consecutive pixels
  {
    const int row = lane >> 1, hx = (lane & 1) * 16;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int c = warp * 4 + cc;
      const float* src = sS + c * TC_SP + row * TC_WP + hx;
      float v[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 t = *reinterpret_cast<const float2*>(src + 2 * k);
        v[2 * k] = t.x; v[2 * k + 1] = t.y;
      }
      if (stats) {   // per-CTA (mean, M2) of this channel over the 16x32 outputs
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) sum += v[k];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float mean = sum * (1.0f / (TC_TH * TC_TW));
        float m2 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const float dl = v[k] - mean; m2 = fmaf(dl, dl, m2); }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
        if (lane == 0) {
          float* dstp = a.part + ((size_t)cta * C + c) * 2;
          dstp[0] = mean;
          dstp[1] = m2;
        }
      }
      float4* dst = reinterpret_cast<float4*>(a.out + (size_t)b * out_bs + ((size_t)c * H + y0 + row) * W + x0 + hx);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 t = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        if (a.accumulate) { const float4 p = dst[k]; t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w; }
        dst[k] = t;
      }
    }
    if (stats) conv_stats_fold(a, nCta, TC_TH * TC_TW, sFin, &sIsLast);
  }
  TC_STAMP(4);
  TC_TIME(dbg_cta, 2);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// weights fp32 [oc][ic][49] -> bf16 hi/lo in the staged layout (tc_woff): per kernel row, per tap group,
// [Wh | Wl] blocks of [chunk = k/8][row = (tap in group, n)][k%8].
//   forward: n = oc, k = ic, tap as is;  dgrad (conv_transpose): n = ic, k = oc, tap -> 48 - tap.
__global__ void prep_weights_tc_kernel(const float* __restrict__ w, __nv_bfloat16* wf, __nv_bfloat16* wd, int oc_stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C * TAPS) return;
  const int oc = i / (C * TAPS), ic = (i / TAPS) % C, tap = i % TAPS, ky = tap / KS, kx = tap - ky * KS;
  const float v = w[(size_t)oc * oc_stride + ic * TAPS + tap];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  if (wf) {
    wf[tc_woff(ky, kx, oc, ic, false) / 2] = h;
    wf[tc_woff(ky, kx, oc, ic, true) / 2] = l;
  }
  if (wd) {
    wd[tc_woff(KS - 1 - ky, KS - 1 - kx, ic, oc, false) / 2] = h;
    wd[tc_woff(KS - 1 - ky, KS - 1 - kx, ic, oc, true) / 2] = l;
  }
}

int launch_conv7_tc(const ConvArgs& a, cudaStream_t st) {
  if (a.v2) return launch_conv7_tc2(a, st);
  dim3 grid(a.W / TC_TW, a.H / TC_TH, a.B);
  if (a.mode == PTK_CONV_BF16) {
    PTK_CUDA_OK(ptk_func_smem(conv7_tc_kernel<true>, TC_SMEM));
    PTK_CUDA_OK(launch_pdl(a.pdl != 0, conv7_tc_kernel<true>, grid, dim3(256), TC_SMEM, st, a));
  } else {
    PTK_CUDA_OK(ptk_func_smem(conv7_tc_kernel<false>, TC_SMEM));
    PTK_CUDA_OK(launch_pdl(a.pdl != 0, conv7_tc_kernel<false>, grid, dim3(256), TC_SMEM, st, a));
  }
  PTK_LAUNCH_OK("conv7_tc_kernel");
  return PTK_OK;
}

#if PTK_DEBUG_STAMPS
int read_tc_times(unsigned long long* out, int n) {
  if (n < 0 || n > TC_TIMES_MAX) { set_error("ptk_debug_tc_times: n = %d outside [0, %d]", n, TC_TIMES_MAX); return PTK_EINVAL; }
  PTK_CUDA_OK(cudaMemcpyFromSymbol(out, g_tc_times, sizeof(unsigned long long) * 3 * n));
  return PTK_OK;
}

int read_tc_debug(long long* out) {
  PTK_CUDA_OK(cudaMemcpyFromSymbol(out, g_tc_dbg, sizeof(long long) * 8));
  return PTK_OK;
}
#endif

int launch_prep_weights_tc(const float* w, void* wf, void* wd, cudaStream_t st, int oc_stride) {
  const int n = C * C * TAPS;
  prep_weights_tc_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, oc_stride ? oc_stride : C * TAPS);
  PTK_LAUNCH_OK("prep_weights_tc_kernel");
  return PTK_OK;
}

}  // namespace ptk
