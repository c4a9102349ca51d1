// Tile geometry and shared-memory layout constants of the second-generation tensor-core convolution (ptk_conv_tc2.cu);
// shared with the persistent time-loop kernel (ptk_int_loop.cu), which runs the same MMA / epilogue schedule per tile.
#pragma once
#include "ptk_common.cuh"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int T2_TH = 16, T2_TW = 32;
constexpr int T2_WP = T2_TW + 6;                       // 38: padded row pitch (pixels)
constexpr int T2_ROWS = T2_TH + 6;                     // 22 input rows
constexpr int T2_NPIX = 880;                           // pixels of the operand image incl. slack (last read: 456 + 232 + 160 = 848)
constexpr int T2_PLANE = T2_NPIX * 16 + 16;            // bytes per channel-chunk plane (+16: the 4 chunks of a pixel hit 4 bank groups)
constexpr int T2_A_BYTES = 2 * 4 * T2_PLANE;           // hi + lo, 4 chunks
constexpr int T2_NG = 2 * KS;                          // tap groups (ky, kxg): kx = 4*kxg + dl, dl = 0..3 (kx = 7 is a zero dummy)
constexpr int T2_WTILE = 2 * 128 * 16;                 // one A operand: [2 k-chunks][128 rows = (dl, n)][8 bf16]
constexpr int T2_STAGE = 4 * T2_WTILE;                 // ic half 0: Wh, Wl; ic half 1: Wh, Wl  = 16 KB
constexpr int T2_NST = 4;                              // weight ring stages
constexpr int T2_NCH = 3;                              // row chunks of the output tile
constexpr int T2_DBUF = 240;                           // TMEM columns per accumulator buffer
constexpr int T2_PP = 36;                              // row pitch (floats) of the quadrant-sum tile
constexpr int T2_PART = 4 * C * T2_PP * 4;             // [4 quadrants][32 oc][32 px (+4)] floats
constexpr int T2_SMEM = T2_A_BYTES + T2_NST * T2_STAGE + 2 * T2_PART + 256;

__host__ __device__ constexpr int t2_chunk_start(int c) { return c * 6 * T2_WP; }                    // 0, 228, 456
__host__ __device__ constexpr int t2_chunk_n(int c) { return c < 2 ? 240 : 160; }                    // >= outputs + 3
__host__ __device__ constexpr int t2_chunk_blocks(int c) { return c < 2 ? 7 : 5; }                   // 32-pixel output blocks

// byte offset of weight element (tap (ky, kx), row n, k) in the staged stream
__host__ __device__ constexpr size_t t2_woff(int ky, int kx, int n, int k, bool lo) {
  return (size_t)(ky * 2 + (kx >> 2)) * T2_STAGE + (size_t)(k >> 4) * (2 * T2_WTILE) + (lo ? T2_WTILE : 0) +
         (size_t)((k >> 3) & 1) * (128 * 16) + (size_t)((kx & 3) * 32 + n) * 16 + (size_t)(k & 7) * 2;
}

// ---- epilogue of one tile, shared by conv7_tc2_kernel and the persistent loop kernels ----------------------------------------
// Eight warps in two groups of four: quadrant dl = ewarp & 3 of TMEM holds tap kx = 4*kxg + dl's share, out[oc][m] = sum_dl
// D[dl][oc][m + dl].  The 32-pixel blocks of a row chunk alternate between the groups (group g takes blocks g, g + 2, ...), each
// group sums its block's four quadrants through its own shared-memory tile, so two blocks are in flight at any time: the drain of
// a chunk takes half as long as with one group, which is what the convolution's tail (the last chunk, after the last MMA) costs.
struct T2Epi {
  float* out_b;        // output of this sample: out + b * out_bs
  int H, W, x0, y0;
  int accumulate;      // out += conv
  float* part_dst;     // this CTA's BatchNorm partials [32][2] (mean, M2) or nullptr
};
constexpr int T2_EPI_WARPS = 8;
// sPart: [2 groups][4 quadrants][32 oc][T2_PP] floats; sMerge: [32][4] floats; dFull / dEmpty: the accumulator barriers
// (dEmpty counts T2_EPI_WARPS arrivals); par1: parity offset of dFull[1] (persistent kernels), 0 for a stand-alone launch.
__device__ __forceinline__ void t2_epilogue(const T2Epi& e, uint32_t tmem_base, float* sPartAll, float* sMerge, uint64_t* dFull,
                                            uint64_t* dEmpty, uint32_t par1, int ewarp, int lane) {
  const int grp = ewarp >> 2, dl = ewarp & 3, et = dl * 32 + lane;
  const int oc = et >> 2, jq = et & 3;              // sum phase: channel, 8-pixel slice of a 32-pixel block
  float* sPart = sPartAll + (size_t)grp * (T2_PART / 4);
  float sK = 0.f, s1 = 0.f, s2 = 0.f;               // shifted sums of this thread's outputs: sum (v - K), sum (v - K)^2
  int sn = 0;
  float* outc = e.out_b + (size_t)oc * e.H * e.W;
  const bool stats = (e.part_dst != nullptr);
  bool firstblk = true;
#pragma unroll 1
  for (int c = 0; c < T2_NCH; ++c) {
    const int buf = c & 1;
    mbar_wait(&dFull[buf], (uint32_t)((c >> 1) & 1) ^ (buf ? par1 : 0u));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int nb = t2_chunk_blocks(c);
    const int last = nb - 1 - ((nb - 1 - grp) & 1);     // this group's last block of the chunk
#pragma unroll 1
    for (int kb = grp; kb < nb; kb += 2) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(dl * 32) << 16) + (uint32_t)(buf * T2_DBUF + kb * 32 + dl), v);   // lane = oc
      if (!firstblk) named_bar_sync(3 + grp, 128);      // the whole group has read the previous block's sums
      firstblk = false;
      float* pw = sPart + (size_t)(dl * C + lane) * T2_PP;
#pragma unroll
      for (int k = 0; k < 8; ++k) *reinterpret_cast<float4*>(pw + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      if (kb == last) {                                 // this warp is done with the TMEM buffer
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&dEmpty[buf]);
      }
      named_bar_sync(1 + grp, 128);
      const float* pr = sPart + (size_t)oc * T2_PP + 8 * jq;
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
      const int m0 = t2_chunk_start(c) + kb * 32 + 8 * jq;      // padded output pixel of o[0]
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = m0 + 2 * u, r = m / T2_WP, col = m - r * T2_WP;       // col is even: both pixels share a row
        if (r < T2_TH && col < T2_TW) {
          float2 t = make_float2(o[2 * u], o[2 * u + 1]);
          float2* dst = reinterpret_cast<float2*>(outc + (size_t)(e.y0 + r) * e.W + e.x0 + col);
          if (e.accumulate) { const float2 q = *dst; t.x += q.x; t.y += q.y; }
          *dst = t;
          if (stats) {
            if (sn == 0) sK = o[2 * u];
            const float d0 = o[2 * u] - sK, d1 = o[2 * u + 1] - sK;
            s1 += d0 + d1;
            s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
            sn += 2;
          }
        }
      }
    }
  }
  if (stats) {
    // per-thread (n, mean, M2) -> Chan merge over the 4 threads of a channel, then over the two groups -> (mean, M2) of 16x32 outputs
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
    if (grp == 1 && jq == 0) { sMerge[oc * 4 + 0] = n; sMerge[oc * 4 + 1] = mean; sMerge[oc * 4 + 2] = m2; }
    named_bar_sync(5, 32 * T2_EPI_WARPS);
    if (grp == 0 && jq == 0) {
      const float n2 = sMerge[oc * 4 + 0], mean2 = sMerge[oc * 4 + 1], m22 = sMerge[oc * 4 + 2];
      const float nt = n + n2, dlt = mean2 - mean;
      if (nt > 0.f) {
        m2 = m2 + m22 + dlt * dlt * (n * n2 / nt);
        mean = mean + dlt * (n2 / nt);
      }
      e.part_dst[oc * 2 + 0] = mean;
      e.part_dst[oc * 2 + 1] = m2;
    }
  }
}

}  // namespace ptk
