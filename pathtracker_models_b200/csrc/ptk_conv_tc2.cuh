// Tile geometry and shared-memory layout constants of the second-generation tensor-core convolution (ptk_conv_tc2.cu);
// shared with the persistent time-loop kernel (ptk_int_loop.cu), which runs the same MMA / epilogue schedule per tile.
#pragma once
#include "ptk_common.cuh"

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

}  // namespace ptk
