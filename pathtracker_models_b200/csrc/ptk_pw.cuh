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
#pragma unroll 8
  for (int ic = 0; ic < C; ++ic) {
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

// ---- tensor-core variant of the gate products (tensor-core modes only) -----------------------------------------
// out[px][n] += sum_k tile[k][px] * W[k][n] on mma.sync.m16n8k8 tf32 with the operands split x = hi + lo in
// registers (mask + subtract) and three products hi*hi + hi*lo + lo*hi: ~2^-20 relative error per product, the
// same idea as the split-bf16 convolutions.  W sits in shared memory in B-fragment order (prep_gates_kernel):
//   WF[((kt*4 + nt)*32 + lane)*2 + j] = W[kt*8 + (lane&3) + 4j][nt*8 + (lane>>2)].
// The accumulators come out in MMA fragment layout; frag_to_tile writes them to a [c][TP] tile (conflict-free),
// from which every thread re-reads its own 4 px x 8 ch patch.  4x fewer issue slots than the FFMA product.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;           // truncation: x - hi is exact and |lo| < 2^-10 |x|
  lo = __float_as_uint(x - __uint_as_float(hi));   // the MMA reads only the tf32 bits of lo (error ~2^-20 |x|)
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
struct Frag { float c[2][4][4]; };   // [m-tile: 16 px][n-tile: 8 ch][4]
__device__ __forceinline__ void frag_zero(Frag& f) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) f.c[m][n][e] = 0.f;
}
__device__ __forceinline__ void warp_mma_gemm(Frag& f, const float* __restrict__ tile, const float* __restrict__ sWF) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float* a = tile + (kt * 8 + t) * TP + m * 16 + g;
      split_tf32(a[0], ah[m][0], al[m][0]);
      split_tf32(a[8], ah[m][1], al[m][1]);
      split_tf32(a[4 * TP], ah[m][2], al[m][2]);
      split_tf32(a[4 * TP + 8], ah[m][3], al[m][3]);
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float2 b = *reinterpret_cast<const float2*>(sWF + ((kt * 4 + n) * 32 + lane) * 2);
      uint32_t bh[2], bl[2];
      split_tf32(b.x, bh[0], bl[0]);
      split_tf32(b.y, bh[1], bl[1]);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        mma_tf32(f.c[m][n], al[m], bh);
        mma_tf32(f.c[m][n], ah[m], bl);
        mma_tf32(f.c[m][n], ah[m], bh);
      }
    }
  }
}
// fragment layout -> [c][TP] tile; callers bracket it with __syncwarp()
__device__ __forceinline__ void frag_to_tile(float* tile, const Frag& f) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        tile[(n * 8 + 2 * t + (e & 1)) * TP + m * 16 + g + 8 * (e >> 1)] = f.c[m][n][e];
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
// streaming store (st.global.cs, evict-first in L2): for tensors nobody reads again inside the reverse loop (the per-step
// operand stores of the deferred weight-gradient launches), so that they do not push the step's working set out of L2
__device__ __forceinline__ void st4s(float* p, const float (&v)[4], bool stream) {
  if (stream) __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  else st4(p, v);
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

// [B,C,P] global tile (base = b*C*P + first pixel of the warp tile) += fragment, in fragment layout: every
// instruction touches 4 channels x 8 consecutive pixels (full 32-byte sectors).  Used where a second
// accumulator set would not fit in registers next to the one still being formed.
__device__ __forceinline__ void frag_add_to_global(float* dst, size_t tile_base, int P, const Frag& f) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float* q = dst + tile_base + (size_t)(n * 8 + 2 * t + (e & 1)) * P + m * 16 + g + 8 * (e >> 1);
        *q += f.c[m][n][e];
      }
}
// acc += the thread's patch of a tile
__device__ __forceinline__ void acc_add_tile(float (&acc)[4][8], const float* tile, const Lane& L) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float4 t = ld4(tile + (L.oc0 + o) * TP + L.pxo);
    acc[0][o] += t.x; acc[1][o] += t.y; acc[2][o] += t.z; acc[3][o] += t.w;
  }
}
// tile patch += [B,C,P] global patch (each thread its own patch)
__device__ __forceinline__ void tile_add_global(float* tile, const float* src, size_t base, int P, const Lane& L) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const float4 a = ld4(src + base + (size_t)(L.oc0 + o) * P);
    float* tp = tile + (L.oc0 + o) * TP + L.pxo;
    const float4 t = ld4(tp);
    *reinterpret_cast<float4*>(tp) = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
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
