// Warp-tile machinery shared by the point-wise (gate / blend / BatchNorm-apply) kernels.
//
// A warp owns a tile of 32 consecutive pixels of one sample and all 32 channels.  Lane (pg, og):
// pg = lane & 7 -> 4 consecutive pixels, og = lane >> 3 -> 8 consecutive channels, so every thread
// owns a 4 px x 8 ch patch for the element-wise math, and a 1x1 gate convolution is a 32x32x32
// register-tiled product against [c][TP] shared-memory tiles (3 LDS.128 per 32 FFMA).  Warps are
// independent (only __syncwarp), the CTA only shares the weight matrices.
#pragma once
#include "ptk_common.cuh"

namespace ptk {

constexpr int WPX = 32;   // pixels per warp tile
constexpr int TP = 36;    // tile row pitch (floats), 16-byte aligned rows

struct Lane {
  int pg, og, oc0, pxo;   // pixel group, channel group, first channel, pixel offset inside the tile
  __device__ __forceinline__ Lane() {
    const int lane = threadIdx.x & 31;
    pg = lane & 7; og = lane >> 3; oc0 = og * 8; pxo = pg * 4;
  }
};

// acc[p][o] += sum_ic tile[ic][pxo+p] * W[ic][oc0+o]      (W: [in][out] in shared memory)
__device__ __forceinline__ void warp_gemm(float (&acc)[4][8], const float* __restrict__ tile,
                                          const float* __restrict__ sW, const Lane& L) {
  const float* tp = tile + L.pxo;
  const float* wp = sW + L.oc0;
#ifndef PTK_GEMM_IC
#define PTK_GEMM_IC C
#endif
#pragma unroll 8
  for (int ic = 0; ic < PTK_GEMM_IC; ++ic) {
    const float4 v = *reinterpret_cast<const float4*>(tp + ic * TP);
    const float4 w0 = *reinterpret_cast<const float4*>(wp + ic * C);
    const float4 w1 = *reinterpret_cast<const float4*>(wp + ic * C + 4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[p][o] = fmaf(vv[p], ww[o], acc[p][o]);
  }
}

// async copy of a [32 ch][32 px] tile of a [B,C,P] tensor into the warp's shared tile
__device__ __forceinline__ void warp_tile_load_async(float* tile, const float* __restrict__ src, size_t bstride,
                                                     int b, int pix_base, int P) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int q = lane + 32 * j, c = q >> 3, part = q & 7;
    cp_async16(tile + c * TP + part * 4, src + (size_t)b * bstride + (size_t)c * P + pix_base + part * 4);
  }
}

// L2 prefetch of the warp's [32 ch][32 px] tile of a [B,C,P] tensor: lane = channel, one 128-byte line each
__device__ __forceinline__ void warp_tile_prefetch(const float* __restrict__ src, size_t tile_base, int P) {
  prefetch_l2(src + tile_base + (size_t)(threadIdx.x & 31) * P);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void ld4v(const float* p, float (&v)[4]) {
  const float4 t = ld4(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
// own patch of a tile (thread's 4 px of channel oc0+o)
__device__ __forceinline__ void tile_ld(const float* tile, const Lane& L, int o, float (&v)[4]) {
  ld4v(tile + (L.oc0 + o) * TP + L.pxo, v);
}
__device__ __forceinline__ void tile_st(float* tile, const Lane& L, int o, const float (&v)[4]) {
  st4(tile + (L.oc0 + o) * TP + L.pxo, v);
}

// the three input channels of the thread's 4 pixels of frame t
__device__ __forceinline__ void frame_px(const FrameView& fv, int b, int t, int pix, int W, float (&xin)[3][4]) {
  const int y = pix / W, x = pix - y * W;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int p = 0; p < 4; ++p) xin[i][p] = fv.at(b, i, t, y, x + p);
}
// stem pre-activation of channel c for the 4 pixels (models/InT.py:192,212)
__device__ __forceinline__ void stem_pre(const float* sWp, const float* sbp, int c, const float (&xin)[3][4],
                                         float (&pre)[4]) {
  const float w0 = sWp[c * 3], w1 = sWp[c * 3 + 1], w2 = sWp[c * 3 + 2], bb = sbp[c];
#pragma unroll
  for (int p = 0; p < 4; ++p) pre[p] = fmaf(w2, xin[2][p], fmaf(w1, xin[1][p], fmaf(w0, xin[0][p], bb)));
}

__device__ __forceinline__ void cta_fill(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
}
// asynchronous variant (cp.async, 16 B per request): n % 4 == 0, both pointers 16-byte aligned.  The caller
// commits / waits; used so that the weight fill overlaps the first tile's operand loads.
__device__ __forceinline__ void cta_fill_async(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) cp_async16(dst + i, src + i);
}

// registers <-> the thread's patch of a tile, all 8 channels (static indices; used to park GEMM
// accumulators in shared memory so that the element-wise math can run in a rolled loop)
__device__ __forceinline__ void acc_to_tile(float* tile, const Lane& L, const float (&acc)[4][8]) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float v[4] = {acc[0][o], acc[1][o], acc[2][o], acc[3][o]};
    tile_st(tile, L, o, v);
  }
}
__device__ __forceinline__ void acc_fill(float (&acc)[4][8], const float* sBias, const Lane& L) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float bb = sBias ? sBias[L.oc0 + o] : 0.0f;
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p][o] = bb;
  }
}
// acc <- [B,C,P] global patch (static)
__device__ __forceinline__ void acc_from_global(float (&acc)[4][8], const float* src, size_t base, int P, const Lane& L) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float4 t = ld4(src + base + (size_t)(L.oc0 + o) * P);
    acc[0][o] = t.x; acc[1][o] = t.y; acc[2][o] = t.z; acc[3][o] = t.w;
  }
}
__device__ __forceinline__ void acc_to_global(float* dst, size_t base, int P, const Lane& L, const float (&acc)[4][8]) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float v[4] = {acc[0][o], acc[1][o], acc[2][o], acc[3][o]};
    st4(dst + base + (size_t)(L.oc0 + o) * P, v);
  }
}

// reduce per-thread channel partials (thread owns channels oc0..oc0+7) over the 8 pixel-group lanes
__device__ __forceinline__ float reduce_pg(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// per-warp channel-sum accumulators in shared memory: sPart[warp][q][c]
template <int NQ>
__device__ __forceinline__ void part_add(float* sPartWarp, int c, const Lane& L, const float (&v)[NQ]) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const float r = reduce_pg(v[q]);
    if (L.pg == 0) sPartWarp[q * C + c] += r;
  }
}
// fold the 4 warps' partials: sTot[q][c]  (all threads; contains __syncthreads)
template <int NQ>
__device__ __forceinline__ void part_fold(const float* sPart /*[4][NQ][32]*/, float* sTot) {
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * C; i += blockDim.x)
    sTot[i] = (sPart[i] + sPart[NQ * C + i]) + (sPart[2 * NQ * C + i] + sPart[3 * NQ * C + i]);
  __syncthreads();
}

}  // namespace ptk
