// Strict-fp32 horizontal-connection convolutions of the InT / hGRU cell (CUDA-core path).
//
//   conv7_ffma_kernel   Y = In *7 W            (forward:  F.conv2d(gated_excitation, w_inh, padding=3),
//                                               models/InT.py:161,172; dgrad: conv_transpose2d with
//                                               flipped / in-out-swapped weights, prepared by the caller)
//                       + per-channel batch statistics of Y for the BatchNorm that follows
//                         (models/InT.py:102,161,172: batch statistics, biased variance, eps 1e-3)
//   wgrad7_ffma_kernel  dW += In (star) dY     (autograd's conv weight gradient, mainclean.py:204)
//
// Layout: state tensors [B,32,H,W] fp32 (NCHW), H % 16 == 0, W % 32 == 0.
// One CTA = 16 rows x 32 columns x 32 output channels of one sample; 256 threads, each thread an
// 8-pixel x 8-channel register tile (64 accumulators); weights and the zero-padded input tile
// are staged in shared memory by cp.async in chunks of 8 input channels.
#include "ptk_common.cuh"
#include "ptk_conv_stats.cuh"
#include "ptk_kernels.h"

namespace ptk {

constexpr int TH = 16, TW = 32;            // output tile
constexpr int IR = TH + KS - 1;            // 22 input rows
constexpr int IP = 44;                     // input row pitch (cols -4 .. 39), conflict-free for LDS.128
constexpr int ICC = 8;                     // input channels per staged chunk
constexpr int IN_CH_PITCH = IR * IP;       // 968
constexpr int W_CHUNK = ICC * TAPS * C;    // 12544 floats
constexpr int ROW_OPS = 14;                // 8 x 16B interior + 3 + 3 halo scalars per input row

__device__ __forceinline__ void stage_input_chunk(float* sIn, const float* __restrict__ in, int b, int ic0,
                                                  int y0, int x0, int H, int W, int ch_pitch, int tid,
                                                  int nthreads) {
  for (int op = tid; op < ICC * IR * ROW_OPS; op += nthreads) {
    const int row = op / ROW_OPS, j = op - row * ROW_OPS;
    const int icl = row / IR, r = row - icl * IR;
    const int gy = y0 + r - KS / 2;
    const bool rv = (gy >= 0) && (gy < H);
    const float* grow = in + ((size_t)(b * C + ic0 + icl) * H + (rv ? gy : 0)) * W;
    float* srow = sIn + icl * ch_pitch + r * IP;
    if (j < 8) {
      cp_async16_zfill(srow + 4 + 4 * j, grow + x0 + 4 * j, rv);
    } else {
      const int c = (j < 11) ? (1 + j - 8) : (36 + j - 11);
      const int gx = x0 + c - 4;
      const bool v = rv && (gx >= 0) && (gx < W);
      cp_async4_zfill(srow + c, grow + (v ? gx : 0), v);
    }
  }
}

__global__ void __launch_bounds__(256, 1) conv7_ffma_kernel(ConvArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;                       // [8][968]
  float* sW = smem + ICC * IN_CH_PITCH;    // [8][49][32]
  __shared__ float sRed[8][8];
  __shared__ double sFin[8][C][3];
  __shared__ int sIsLast;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cg = lane & 3, rl = lane >> 2;     // column group (8 px), row within 8
  const int ocg = warp & 3, rh = warp >> 2;    // output-channel group (8 oc), row half
  const int ty = rh * 8 + rl;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
  const int H = a.H, W = a.W;

  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[j][o] = 0.f;

  for (int chunk = 0; chunk < C / ICC; ++chunk) {
    stage_input_chunk(sIn, a.in, b, chunk * ICC, y0, x0, H, W, IN_CH_PITCH, tid, 256);
    const float* wsrc = a.wT + (size_t)chunk * W_CHUNK;
    for (int i = tid; i < W_CHUNK / 4; i += 256) cp_async16(sW + 4 * i, wsrc + 4 * i);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

#pragma unroll 1
    for (int ic = 0; ic < ICC; ++ic) {
      const float* sI = sIn + ic * IN_CH_PITCH + ty * IP + cg * 8;
      const float* sWi = sW + ic * TAPS * C + ocg * 8;
#pragma unroll 1
      for (int ky = 0; ky < KS; ++ky) {
        float in[16];
        {
          const float4* p = reinterpret_cast<const float4*>(sI + ky * IP);
          const float4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
          in[0] = v0.x; in[1] = v0.y; in[2] = v0.z; in[3] = v0.w;
          in[4] = v1.x; in[5] = v1.y; in[6] = v1.z; in[7] = v1.w;
          in[8] = v2.x; in[9] = v2.y; in[10] = v2.z; in[11] = v2.w;
          in[12] = v3.x; in[13] = v3.y; in[14] = v3.z; in[15] = v3.w;
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(sWi + (ky * KS + kx) * C);
          const float4 w0 = wp[0], w1 = wp[1];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = in[j + kx + 1];
            acc[j][0] = fmaf(v, w0.x, acc[j][0]);
            acc[j][1] = fmaf(v, w0.y, acc[j][1]);
            acc[j][2] = fmaf(v, w0.z, acc[j][2]);
            acc[j][3] = fmaf(v, w0.w, acc[j][3]);
            acc[j][4] = fmaf(v, w1.x, acc[j][4]);
            acc[j][5] = fmaf(v, w1.y, acc[j][5]);
            acc[j][6] = fmaf(v, w1.z, acc[j][6]);
            acc[j][7] = fmaf(v, w1.w, acc[j][7]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- per-channel batch statistics of this CTA's 512 conv outputs (before any accumulate) ----
  if (a.part != nullptr) {
    const int nCta = gridDim.x * gridDim.y * gridDim.z;
    const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    float s[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) v += acc[j][o];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      s[o] = v;
    }
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 8; ++o) sRed[warp][o] = s[o];
    }
    __syncthreads();
    float mean[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) mean[o] = (sRed[ocg][o] + sRed[ocg + 4][o]) * (1.0f / (TH * TW));
    __syncthreads();
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = acc[j][o] - mean[o];
        v = fmaf(dlt, dlt, v);
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      s[o] = v;
    }
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 8; ++o) sRed[warp][o] = s[o];
    }
    __syncthreads();
    if (warp < 4 && lane == 0) {  // warps 0..3 hold ocg == warp
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        float* dst = a.part + ((size_t)cta * C + warp * 8 + o) * 2;
        dst[0] = mean[o];
        dst[1] = sRed[warp][o] + sRed[warp + 4][o];
      }
    }
    conv_stats_fold(a, nCta, TH * TW, sFin, &sIsLast);
  }

  // ---- store ----------------------------------------------------------------------------------
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float* dst = a.out + ((size_t)(b * C + ocg * 8 + o) * H + y0 + ty) * W + x0 + cg * 8;
    float4 v0 = make_float4(acc[0][o], acc[1][o], acc[2][o], acc[3][o]);
    float4 v1 = make_float4(acc[4][o], acc[5][o], acc[6][o], acc[7][o]);
    if (a.accumulate) {
      const float4 p0 = *reinterpret_cast<const float4*>(dst);
      const float4 p1 = *reinterpret_cast<const float4*>(dst + 4);
      v0.x += p0.x; v0.y += p0.y; v0.z += p0.z; v0.w += p0.w;
      v1.x += p1.x; v1.y += p1.y; v1.z += p1.z; v1.w += p1.w;
    }
    *reinterpret_cast<float4*>(dst) = v0;
    *reinterpret_cast<float4*>(dst + 4) = v1;
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient:  dW[oc][ic][ky][kx] += sum_{b,y,x} dY[b,oc,y,x] * In[b,ic,y+ky-3,x+kx-3]
// grid = (4 input-channel chunks, G tile groups); thread = one (oc, ic) pair holding all 49 taps.
// The per-CTA accumulators live in `wacc` ([4*G][49][256]) across timesteps (read-add-write by the
// owning CTA only => deterministic), and are folded once by the finalize kernel.
constexpr int WG_IN_PITCH = IN_CH_PITCH + 4;  // 972: (8 ic) x LDS.128 conflict-free

__global__ void __launch_bounds__(256, 1) wgrad7_ffma_kernel(WgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sDY = smem;                     // [32 oc][16][32]
  float* sIn = smem + C * TH * TW;       // [8 ic][972]
  const int tid = threadIdx.x;
  const int ic = tid & 7, oc = tid >> 3;
  const int chunk = blockIdx.x, grp = blockIdx.y, G = gridDim.y;
  const int H = a.H, W = a.W;
  const int tilesX = W / TW, tilesY = H / TH;
  const int nTiles = a.B * tilesY * tilesX;

  float acc[TAPS];
#pragma unroll
  for (int i = 0; i < TAPS; ++i) acc[i] = 0.f;

  for (int tile = grp; tile < nTiles; tile += G) {
    const int b = tile / (tilesY * tilesX);
    const int rem = tile - b * tilesY * tilesX;
    const int y0 = (rem / tilesX) * TH, x0 = (rem % tilesX) * TW;
    __syncthreads();
    for (int i = tid; i < C * TH * (TW / 4); i += 256) {
      const int o = i / (TH * TW / 4), r = (i / (TW / 4)) % TH, q = i % (TW / 4);
      cp_async16(sDY + (o * TH + r) * TW + 4 * q, a.dy + ((size_t)(b * C + o) * H + y0 + r) * W + x0 + 4 * q);
    }
    stage_input_chunk(sIn, a.in, b, chunk * ICC, y0, x0, H, W, WG_IN_PITCH, tid, 256);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

#pragma unroll 1
    for (int y = 0; y < TH; ++y) {
#pragma unroll 1
      for (int xc = 0; xc < TW / 8; ++xc) {
        float dy[8];
        {
          const float4* p = reinterpret_cast<const float4*>(sDY + (oc * TH + y) * TW + xc * 8);
          const float4 v0 = p[0], v1 = p[1];
          dy[0] = v0.x; dy[1] = v0.y; dy[2] = v0.z; dy[3] = v0.w;
          dy[4] = v1.x; dy[5] = v1.y; dy[6] = v1.z; dy[7] = v1.w;
        }
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          float in[16];
          const float4* p = reinterpret_cast<const float4*>(sIn + ic * WG_IN_PITCH + (y + ky) * IP + xc * 8);
          const float4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
          in[0] = v0.x; in[1] = v0.y; in[2] = v0.z; in[3] = v0.w;
          in[4] = v1.x; in[5] = v1.y; in[6] = v1.z; in[7] = v1.w;
          in[8] = v2.x; in[9] = v2.y; in[10] = v2.z; in[11] = v2.w;
          in[12] = v3.x; in[13] = v3.y; in[14] = v3.z; in[15] = v3.w;
#pragma unroll
          for (int kx = 0; kx < KS; ++kx)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[ky * KS + kx] = fmaf(dy[j], in[j + kx + 1], acc[ky * KS + kx]);
        }
      }
    }
  }
  float* dst = a.wacc + (size_t)(grp * gridDim.x + chunk) * TAPS * 256;
  if (a.first) {
#pragma unroll
    for (int i = 0; i < TAPS; ++i) dst[i * 256 + tid] = acc[i];
  } else {
#pragma unroll
    for (int i = 0; i < TAPS; ++i) dst[i * 256 + tid] += acc[i];
  }
}

// ---- host launchers ----------------------------------------------------------------------------
static constexpr size_t CONV_SMEM = (size_t)(ICC * IN_CH_PITCH + W_CHUNK) * sizeof(float);
static constexpr size_t WGRAD_SMEM = (size_t)(C * TH * TW + ICC * WG_IN_PITCH) * sizeof(float);

int conv_num_ctas(int B, int H, int W) { return B * (H / TH) * (W / TW); }

int launch_conv7(const ConvArgs& a, cudaStream_t st) {
  if (a.mode != PTK_CONV_FP32) return launch_conv7_tc(a, st);
  static thread_local int configured_dev = -1;
  int dev;
  PTK_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    PTK_CUDA_OK(cudaFuncSetAttribute(conv7_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV_SMEM));
    PTK_CUDA_OK(cudaFuncSetAttribute(wgrad7_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGRAD_SMEM));
    configured_dev = dev;
  }
  dim3 grid(a.W / TW, a.H / TH, a.B);
  conv7_ffma_kernel<<<grid, 256, CONV_SMEM, st>>>(a);
  PTK_LAUNCH_OK("conv7_ffma_kernel");
  return PTK_OK;
}

int launch_wgrad7(const WgradArgs& a, int groups, cudaStream_t st) {
  static thread_local int configured_dev = -1;
  int dev;
  PTK_CUDA_OK(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    PTK_CUDA_OK(cudaFuncSetAttribute(wgrad7_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGRAD_SMEM));
    configured_dev = dev;
  }
  dim3 grid(C / ICC, groups);
  wgrad7_ffma_kernel<<<grid, 256, WGRAD_SMEM, st>>>(a);
  PTK_LAUNCH_OK("wgrad7_ffma_kernel");
  return PTK_OK;
}

}  // namespace ptk
