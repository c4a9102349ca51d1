// Tensor-core weight gradient of the 7x7 horizontal-connection convolutions (tcgen05 / TMEM).
//
//   dW[oc][ic][ky][kx] += sum_{y,x} dY[b,oc,y,x] * In[b,ic,y+ky-3,x+kx-3]       (per CTA: 16 rows x 32 cols of one sample)
//
// GEMM view per tap: D[m = oc][n = ic] = sum_{k = pixel} dY[k][m] * In[k + shift(tap)][n]; the pixel
// axis is K, so both operands are MN-major.  Both tiles sit in shared memory in the channel-chunk
// planar form of ptk_conv_tc.cu ([chunk][pixel][8 ch], 16 B per pixel) with the hi planes followed by
// the lo planes.  One instruction with M = 64 rows = [dYh(32 oc) ; dYl(32 oc)] and N = 64 columns =
// [Ih(32 ic) | Il(32 ic)] therefore produces all four products of the split
//   (dYh + dYl) * (Ih + Il)
// at once; the epilogue adds the two row halves and the two column halves.  A tap is just a different
// start pixel of the In operand (ky*38 + kx), exactly as in the forward kernel.
//
// Taps are processed in rounds of 4 (4 x 64 = 256 TMEM columns, double buffered): while the MMA
// thread accumulates round r+1 over the 38 K-steps, four epilogue warps drain round r and add it into
// this CTA's private accumulator slot in global memory (read-add-write by the owning CTA only =>
// deterministic; `finalize` folds the slots).
#include "ptk_common.cuh"
#include "ptk_kernels.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int WT_TH = 16, WT_TW = 32, WT_WP = 38, WT_ROWS = 22;
constexpr int WT_KPIX = WT_TH * WT_TW;               // 512 pixels = 32 K-steps of 16: the dY tile is stored compact (row pitch 32)
constexpr int WT_KSTEPS = WT_KPIX / 16;              // a K-step = 16 consecutive pixels of ONE image row, so the In operand's
                                                     // matching 16 pixels are contiguous too and the 6 pad columns cost no MMAs
constexpr int WT_DY_PLANE = WT_KPIX * 16 + 16;       // 8208 B: +16 so the 8 channel-chunk planes of one K row hit 8 different bank groups
constexpr int WT_IN_NPIX = 880;                      // >= 608 + 6*38 + 6
constexpr int WT_IN_PLANE = WT_IN_NPIX * 16 + 16;    // 14096 B (same reason)
constexpr int WT_DY_BYTES = 8 * WT_DY_PLANE;         // hi chunks 0-3, lo chunks 4-7
constexpr int WT_IN_BYTES = 8 * WT_IN_PLANE;
constexpr int WT_RT = 4;                             // taps per round
constexpr int WT_ROUNDS = (TAPS + WT_RT - 1) / WT_RT;
constexpr int WT_OP = 36;                            // staging row pitch (floats)
constexpr int WT_OUT_BYTES = WT_RT * 2 * C * WT_OP * 4;
constexpr int WT_SMEM = WT_DY_BYTES + WT_IN_BYTES + WT_OUT_BYTES + 128;

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both MN-major, M=64, N=64
__host__ __device__ constexpr uint32_t wt_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
         ((uint32_t)(64 >> 4) << 24);
}

// fp32 NCHW rows -> bf16 hi/lo channel-chunk planar tile (pixel pitch 16 B, row pitch 38 pixels).
// rows [r0, r0+nrows) of the tile map to image rows gy0 + r; columns: 32 interior (+3 halo each side when
// `halo` and the neighbour exists).  `col0` = padded column of image column x0.
__device__ __forceinline__ void wt_stage(uint8_t* sHi, uint8_t* sLo, int plane_bytes, const float* __restrict__ src,
                                         size_t bs, int b, int gy0, int nrows, int x0, int H, int W, int col0, bool halo, int pitch) {
  const int tid = threadIdx.x;
  // two items per round: 16 LDG.128 in flight per thread before the first conversion (L2-latency bound)
  const int nit = 4 * nrows * (WT_TW / 4);
  for (int it0 = tid; it0 < nit; it0 += 512) {
    float4 v[2][8];
    int pp[2], qq[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int it = it0 + 256 * r;
      // quarter warp = (2 adjacent pixel groups) x (4 chunk planes, 16 B apart modulo the bank period): conflict-free
      // 16-byte stores (see conv7_tc_kernel); a warp covers one tile row of all 32 channels
      const int q = (it >> 1) & 3, row = it >> 5, g = ((it >> 3) & 3) * 2 + (it & 1);
      const int gy = gy0 + row;
      pp[r] = -1; qq[r] = q;
      if (it < nit && gy >= 0 && gy < H) {
        pp[r] = row * pitch + col0 + 4 * g;
        const float* p0 = src + (size_t)b * bs + ((size_t)(q * 8) * H + gy) * W + x0 + 4 * g;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[r][e] = *reinterpret_cast<const float4*>(p0 + (size_t)e * H * W);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (pp[r] < 0) continue;
      const int p = pp[r], q = qq[r];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          split_bf16x2(reinterpret_cast<const float*>(&v[r][2 * e])[px], reinterpret_cast<const float*>(&v[r][2 * e + 1])[px], hi[e], lo[e]);
        }
        *reinterpret_cast<uint4*>(sHi + q * plane_bytes + (p + px) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sLo + q * plane_bytes + (p + px) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  if (halo && W > WT_TW) {
    for (int it = tid; it < 4 * nrows * 6; it += 256) {
      const int q = it / (nrows * 6), rem = it - q * (nrows * 6), r = rem / 6, hc = rem % 6;
      const int c = hc < 3 ? hc : (WT_TW + hc);
      const int gy = gy0 + r, gx = x0 + c - 3;
      if (gy < 0 || gy >= H || gx < 0 || gx >= W) continue;
      const float* p0 = src + (size_t)b * bs + ((size_t)(q * 8) * H + gy) * W + gx;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        split_bf16x2(p0[(size_t)(2 * e) * H * W], p0[(size_t)(2 * e + 1) * H * W], hi[e], lo[e]);
      }
      const int p = r * pitch + c;
      *reinterpret_cast<uint4*>(sHi + q * plane_bytes + p * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sLo + q * plane_bytes + p * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

#if PTK_DEBUG_STAMPS   // tuning instrumentation, not in the shipped library (see ptk_conv_tc.cu)
__device__ long long g_wt_dbg[8];
#define WT_STAMP(i, t) do { if (blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.x == (t)) g_wt_dbg[i] = clock64(); } while (0)
#else
#define WT_STAMP(i, t) do { } while (0)
#endif

template <bool SINGLE>
__global__ void __launch_bounds__(256, 1) wgrad7_tc_kernel(WgradArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sDY = smem_raw;                                  // [8 planes][608 px][16 B]
  uint8_t* sIN = smem_raw + WT_DY_BYTES;                    // [8 planes][880 px][16 B]
  float* sOut = reinterpret_cast<float*>(smem_raw + WT_DY_BYTES + WT_IN_BYTES);   // [4 taps][hi/lo][32 oc][36]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + WT_DY_BYTES + WT_IN_BYTES + WT_OUT_BYTES);
  uint64_t* tfull = bars;        // [2]
  uint64_t* tempty = bars + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform => uniform branches / registers
  const int x0 = blockIdx.x * WT_TW, y0 = blockIdx.y * WT_TH, b = blockIdx.z;
  const int H = a.H, W = a.W;
  const size_t in_bs = a.in_bs ? (size_t)a.in_bs : (size_t)C * H * W, dy_bs = a.dy_bs ? (size_t)a.dy_bs : (size_t)C * H * W;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;

  WT_STAMP(0, 0);
  if (tid == 0) {
    mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 128); mbar_init(&tempty[1], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < (WT_DY_BYTES + WT_IN_BYTES) / 16; i += 256) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  // the In operand is usually a forward checkpoint (I_{t+1}): pull the tile's lines HBM -> L2 before the wait
  for (int it = tid; it < C * WT_ROWS; it += 256) {
    const int ch = it / WT_ROWS, gy = y0 - 3 + (it - ch * WT_ROWS);
    if (gy >= 0 && gy < H) prefetch_l2(a.in + (size_t)b * in_bs + ((size_t)ch * H + gy) * W + x0);
  }
  __syncthreads();
  pdl_wait();                 // the prologue above touched only shared memory / TMEM (ptk_common.cuh, PDL)
  pdl_launch_dependents();
  WT_STAMP(1, 0);
  // dY tile: 16 rows, interior columns only (pad columns stay zero => contribute nothing)
  wt_stage(sDY, sDY + 4 * WT_DY_PLANE, WT_DY_PLANE, a.dy, dy_bs, b, y0, WT_TH, x0, H, W, 0, false, WT_TW);
  // In tile: 22 rows with the 3-pixel halo, interior at padded column 3
  wt_stage(sIN, sIN + 4 * WT_IN_PLANE, WT_IN_PLANE, a.in, in_bs, b, y0 - 3, WT_ROWS, x0, H, W, 3, true, WT_WP);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  WT_STAMP(2, 0);

  if (warp == 0) {
    // ---- MMA issuer (whole warp converged; one elected lane issues) ------------------------------------------------------------------------------
    // MN-major no-swizzle descriptors: LBO = stride between 8-pixel K groups (128 B), SBO = plane stride
    const uint64_t dA0 = umma_desc(smem_u32(sDY), 128, WT_DY_PLANE);
    const uint64_t dB0 = umma_desc(smem_u32(sIN), 128, WT_IN_PLANE);
    const uint32_t aHiW = (uint32_t)(dA0 >> 32), aLoW = (uint32_t)dA0;
    const uint32_t bHiW = (uint32_t)(dB0 >> 32), bLoW = (uint32_t)dB0;
    constexpr uint32_t IDESC = SINGLE ? ((wt_idesc() & ~(0x3Fu << 17)) | ((uint32_t)(32 >> 3) << 17)) : wt_idesc();   // bf16: N = 32 (Ih only)
    for (int r = 0; r < WT_ROUNDS; ++r) {
      const int buf = r & 1, use = r >> 1;
      if (use > 0) mbar_wait(&tempty[buf], (uint32_t)((use - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int t0 = r * WT_RT, nt = (TAPS - t0) < WT_RT ? (TAPS - t0) : WT_RT;
      for (int tl = 0; tl < nt; ++tl) {
        const int tap = t0 + tl, ky = tap / KS, kx = tap - ky * KS;
        const uint32_t bOff = bLoW + (uint32_t)(ky * WT_WP + kx);      // 16-byte units == pixels
        const uint32_t dcol = tmem_base + (uint32_t)(buf * 256 + tl * 64);
#pragma unroll 2
        for (int ks = 0; ks < WT_KSTEPS; ++ks) {
          // K-step ks = pixels 16h .. 16h+15 of tile row r: dY compact (pitch 32), In padded (pitch 38)
          const uint64_t dA = ((uint64_t)aHiW << 32) | (aLoW + (uint32_t)(ks * 16));
          const uint64_t dB = ((uint64_t)bHiW << 32) | (bOff + (uint32_t)((ks >> 1) * WT_WP + (ks & 1) * 16));
          if (elect_one()) umma_bf16(dcol, dA, dB, IDESC, ks ? 1u : 0u);
        }
      }
      __syncwarp();
      if (elect_one()) umma_commit(&tfull[buf]);
    }
    WT_STAMP(3, 0);
  } else if (warp >= 4) {
    // ---- epilogue warps: drain a round, combine the four split products, add into the CTA's slot ----
    const int quad = warp & 3, et = tid - 128;
    const int half = quad >> 1, oc = (quad & 1) * 16 + lane;      // valid for lane < 16 (M = 64 TMEM layout)
    float* slot = a.wacc + (size_t)cta * TAPS * C * C;
    for (int r = 0; r < WT_ROUNDS; ++r) {
      const int buf = r & 1, use = r >> 1;
      const int t0 = r * WT_RT, nt = (TAPS - t0) < WT_RT ? (TAPS - t0) : WT_RT;
      mbar_wait(&tfull[buf], (uint32_t)(use & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int tl = 0; tl < nt; ++tl) {
        float v[32], w[32];
        const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + tl * 64);
        tmem_ld32(ta, v);
        if (SINGLE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = 0.f;
        } else {
          tmem_ld32(ta + 32, w);
        }
        if (lane < 16) {
          float* dst = sOut + ((tl * 2 + half) * C + oc) * WT_OP;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(v[i] + w[i], v[i + 1] + w[i + 1], v[i + 2] + w[i + 2], v[i + 3] + w[i + 3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tempty[buf]);                     // TMEM buffer may be overwritten by round r+2
      named_bar_sync(1, 128);
      // 4 taps x 1024 floats = 8 float4 per thread: issue all loads first, then add and store
      {
        float4 prev[8], val[8];
        float* gp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i4 = et + 128 * k;                         // float4 index inside the round
          const int tl = i4 >> 8, o = (i4 >> 3) & 31, ic = (i4 & 7) * 4;
          gp[k] = slot + (size_t)(t0 + tl) * C * C + o * C + ic;
          const bool on = tl < nt;
          if (on && !a.first) prev[k] = *reinterpret_cast<const float4*>(gp[k]);
          else prev[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 h = *reinterpret_cast<const float4*>(sOut + ((tl * 2 + 0) * C + o) * WT_OP + ic);
          const float4 l = *reinterpret_cast<const float4*>(sOut + ((tl * 2 + 1) * C + o) * WT_OP + ic);
          val[k] = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int tl = (et + 128 * k) >> 8;
          if (tl < nt)
            *reinterpret_cast<float4*>(gp[k]) = make_float4(prev[k].x + val[k].x, prev[k].y + val[k].y,
                                                            prev[k].z + val[k].z, prev[k].w + val[k].w);
        }
      }
      named_bar_sync(1, 128);
    }
  }
  WT_STAMP(4, 128);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

#if PTK_DEBUG_STAMPS
int read_wt_debug(long long* out) {
  PTK_CUDA_OK(cudaMemcpyFromSymbol(out, g_wt_dbg, sizeof(long long) * 8));
  return PTK_OK;
}
#endif

size_t wgrad_tc_acc_floats(int B, int H, int W) { return (size_t)B * (H / WT_TH) * (W / WT_TW) * TAPS * C * C; }

int launch_wgrad7_tc(const WgradArgs& a, cudaStream_t st) {
  dim3 grid(a.W / WT_TW, a.H / WT_TH, a.B);
  if (a.mode == PTK_CONV_BF16) {
    PTK_CUDA_OK(ptk_func_smem(wgrad7_tc_kernel<true>, WT_SMEM));
    PTK_CUDA_OK(launch_pdl(true, wgrad7_tc_kernel<true>, grid, dim3(256), WT_SMEM, st, a));
  } else {
    PTK_CUDA_OK(ptk_func_smem(wgrad7_tc_kernel<false>, WT_SMEM));
    PTK_CUDA_OK(launch_pdl(true, wgrad7_tc_kernel<false>, grid, dim3(256), WT_SMEM, st, a));
  }
  PTK_LAUNCH_OK("wgrad7_tc_kernel");
  return PTK_OK;
}

}  // namespace ptk
