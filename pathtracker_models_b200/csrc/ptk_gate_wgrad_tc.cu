// Gate weight / bias gradients of ALL time steps on tcgen05 (tensor-core modes; one launch after the reverse loop).
//
//   dWe_w = sum ge (x) S,  dWe_u = sum ge (x) Ge,  dWi_w = sum gi (x) X,  dWi_u = sum gi (x) S,  dWa_w = sum a (x) X,
//   dWa_u = sum a (x) E;   d(bias) = sum ge / gi / a        (sums over pixels, samples and time steps; the backward of the six
//   nn.Conv2d(32, 32, 1) of models/InT.py:110-121 / models/ffhgru_hierarchy.py:104-113)
//
// The mma.sync version (gate_wgrad_kernel<CELL, true>, ptk_int_bwd.cu) streams its 3.7 GB of operands at 2.9 TB/s: it is bound by
// the legacy tensor pipe and the per-warp operand splits, not by HBM (profiles/r2z_kernels_full.txt: tensor pipe 44 %, issue 35 %).
// Here the six products are ONE GEMM per 16 pixels:  D[128][144] += A[128][16 px] * B[144][16 px]^T  with
//   A rows:  0-31 ge, 32-63 gi, 64-95 a, 96-127 zero               (the three gate deltas)
//   B rows:  0-31 S, 32-63 Ge, 64-95 X, 96-127 E, 128 ones, 129-143 zero   (the four gate inputs + the bias column)
// of which six 32 x 32 blocks and three bias columns are the gradients (the tensor pipe has room for the unused blocks: 216 cycles
// per 16 pixels and SM against ~560 of HBM time).  Pixels are the K axis and contiguous in memory, so both operands are K-major:
// producers convert fp32 rows to bf16 hi/lo chunk planes [8 px][row][16 B] (3-pass split hi*hi + hi*lo + lo*hi, as everywhere).
// Accumulation is two-level, as in the mma.sync version: a TMEM buffer holds one time step's sum, three flush warps retire it into
// fp32 running totals in registers while the next step accumulates in the second buffer (one long chain over all steps lost 1e-4
// on the hGRU gate gradients at T = 64).  Persistent CTAs, one per SM; slots folded in fp64 by int_finalize (deterministic).
#include <cuda.h>
#include <cuda_bf16.h>

#include "ptk_int.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int GW_THREADS = 512;                 // warp 0: MMA issuer; warps 4-6: flush (TMEM quadrants 0-2); the other 12: converters
constexpr int GW_NCONV = 12;
constexpr int GW_PX = 32;                       // pixels per tile (= per stage)
constexpr int GW_KC = GW_PX / 8;                // 8-pixel chunks per stage
constexpr int GW_MA = 128, GW_NB = 144;
constexpr int GW_PLANE_A = GW_MA * 16 + 16;     // +16: the chunk lanes of a quarter warp hit different bank groups
constexpr int GW_PLANE_B = GW_NB * 16 + 16;
constexpr int GW_SKEW = (64 + 128 - (GW_KC * 16) % 128) % 128;   // lo planes follow the hi planes at an offset = 64 (mod 128): a hi / lo
constexpr int GW_LO_A = GW_KC * GW_PLANE_A + GW_SKEW;           // lane pair of the same chunk then hits different bank groups
constexpr int GW_LO_B = GW_KC * GW_PLANE_B + GW_SKEW;
static_assert(GW_LO_A % 128 == 64 && GW_LO_B % 128 == 64, "bank skew");
constexpr int GW_A_BYTES = GW_LO_A + GW_KC * GW_PLANE_A + 64;
constexpr int GW_B_BYTES = GW_LO_B + GW_KC * GW_PLANE_B + 64;
constexpr int GW_OP_STAGE = GW_A_BYTES + GW_B_BYTES;  // one operand stage (bf16 hi/lo chunk planes)
constexpr int GW_NOP = 2;
constexpr int GW_RAW_STAGE = 7 * C * GW_PX * 4;       // one raw stage: [7 tensors][32 ch][32 px] fp32, rows of 128 B
constexpr int GW_NRAW = 5;                            // four tiles in flight while one is converted
constexpr int GW_OP0 = GW_NRAW * GW_RAW_STAGE;
constexpr int GW_BARS = GW_OP0 + GW_NOP * GW_OP_STAGE;
constexpr int GW_SMEM = GW_BARS + 256;
constexpr int GW_DCOLS = 160;                   // TMEM columns per accumulator buffer (144 used)

// the seven operand tensors as 2-D tensor maps [rows = (t, b, c)][P pixels] (fp32, box = 32 channels x 32 pixels): order ge, gi, a, S, Ge, X, E
struct GwMaps { CUtensorMap m[7]; };

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

template <int CELL>
__global__ void __launch_bounds__(GW_THREADS, 1) gate_wgrad_tc_kernel(IntDev d, int nT, const __grid_constant__ GwMaps maps) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + GW_BARS);
  uint64_t* rawFull = bars;                     // [GW_NRAW] raw tile landed (bulk copies, tx count)
  uint64_t* rawEmpty = rawFull + GW_NRAW;       // [GW_NRAW] ... and converted by all converter warps
  uint64_t* opFull = rawEmpty + GW_NRAW;        // [GW_NOP] operand stage written (all converter warps)
  uint64_t* opEmpty = opFull + GW_NOP;          // [GW_NOP] ... and consumed (tcgen05.commit)
  uint64_t* dFull = opEmpty + GW_NOP;           // [2] a time step's sum is complete
  uint64_t* dEmpty = dFull + 2;                 // [2] ... and has been retired (3 flush warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dEmpty + 2);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int P = d.H * d.W, tpi = P / GW_PX, perT = d.B * tpi, nTiles = nT * perT;

  if (tid == 0) {
    for (int i = 0; i < GW_NRAW; ++i) { mbar_init(&rawFull[i], 1); mbar_init(&rawEmpty[i], GW_NCONV); }
    for (int i = 0; i < GW_NOP; ++i) { mbar_init(&opFull[i], GW_NCONV); mbar_init(&opEmpty[i], 1); }
    mbar_init(&dFull[0], 1); mbar_init(&dFull[1], 1);
    mbar_init(&dEmpty[0], 3); mbar_init(&dEmpty[1], 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows the converters never write: A rows 96-127 and B rows 129-143 stay zero, B row 128 is the bias column (ones)
  for (int i = tid; i < GW_NOP * GW_OP_STAGE / 16; i += GW_THREADS) reinterpret_cast<uint4*>(smem_raw + GW_OP0)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid < GW_NOP * GW_KC) {
    const int st = tid / GW_KC, kc = tid % GW_KC;
    *reinterpret_cast<uint4*>(smem_raw + GW_OP0 + (size_t)st * GW_OP_STAGE + GW_A_BYTES + kc * GW_PLANE_B + 128 * 16) =
        make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ---- MMA issuer ---------------------------------------------------------------------------------------------------------
    const uint64_t dA0 = umma_desc(0, GW_PLANE_A, 128), dB0 = umma_desc(0, GW_PLANE_B, 128);
    const uint32_t aW = (uint32_t)(dA0 >> 32), bW = (uint32_t)(dB0 >> 32);
    const uint32_t idesc = umma_idesc(GW_NB);
    const uint32_t sbase = smem_u32(smem_raw + GW_OP0);
    int it = 0, ev = 0, t_cur = -1;
    bool fresh = true;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++it) {
      const int t = tile / perT;
      if (t != t_cur) {
        if (t_cur >= 0) {                         // the finished step's sum goes to the flush warps, the next one to the other buffer
          __syncwarp();
          if (elect_one()) umma_commit(&dFull[ev & 1]);
          ++ev;
          if (ev >= 2) {
            mbar_wait(&dEmpty[ev & 1], (uint32_t)(((ev >> 1) - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
        }
        t_cur = t;
        fresh = true;
      }
      const int st = it % GW_NOP;
      mbar_wait(&opFull[st], (uint32_t)((it / GW_NOP) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aH0 = sbase + (uint32_t)st * GW_OP_STAGE, bH0 = aH0 + GW_A_BYTES;
      const uint32_t dcol = tmem_base + (uint32_t)((ev & 1) * GW_DCOLS);
#pragma unroll
      for (int ks = 0; ks < GW_PX / 16; ++ks) {
        const uint32_t aH = (uint32_t)dA0 + ((aH0 + (uint32_t)(2 * ks) * GW_PLANE_A) >> 4);
        const uint32_t aL = aH + ((uint32_t)GW_LO_A >> 4);
        const uint32_t bH = (uint32_t)dB0 + ((bH0 + (uint32_t)(2 * ks) * GW_PLANE_B) >> 4);
        const uint32_t bL = bH + ((uint32_t)GW_LO_B >> 4);
        const uint64_t dAh = ((uint64_t)aW << 32) | aH, dAl = ((uint64_t)aW << 32) | aL;
        const uint64_t dBh = ((uint64_t)bW << 32) | bH, dBl = ((uint64_t)bW << 32) | bL;
        if (elect_one()) {
          umma_bf16(dcol, dAh, dBh, idesc, (fresh && ks == 0) ? 0u : 1u);
          umma_bf16(dcol, dAh, dBl, idesc, 1u);
          umma_bf16(dcol, dAl, dBh, idesc, 1u);
        }
      }
      fresh = false;
      __syncwarp();
      if (elect_one()) umma_commit(&opEmpty[st]);
    }
    __syncwarp();
    if (t_cur >= 0 && elect_one()) umma_commit(&dFull[ev & 1]);
    __syncwarp();
  } else if (warp >= 4 && warp < 7) {
    // ---- flush warps: quadrant q = the gate delta (0: ge, 1: gi, 2: a); thread = output channel o -----------------------------
    const int q = warp - 4;
    const int c1 = q == 0 ? 0 : 64;               // first input block:  ge x S,  gi x X,  a x X
    const int c2 = q == 0 ? 32 : (q == 1 ? 0 : 96);   // second:         ge x Ge, gi x S,  a x E
    float tot1[32], tot2[32], btot = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { tot1[i] = 0.f; tot2[i] = 0.f; }
    // the steps this CTA visits: same walk as the MMA warp
    int nev = 0, t_cur = -1;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
      const int t = tile / perT;
      if (t != t_cur) { t_cur = t; ++nev; }
    }
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int ev = 0; ev < nev; ++ev) {
      const int buf = ev & 1;
      mbar_wait(&dFull[buf], (uint32_t)((ev >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[32];
      tmem_ld32(tq + (uint32_t)(buf * GW_DCOLS + c1), v);
#pragma unroll
      for (int i = 0; i < 32; ++i) tot1[i] += v[i];
      tmem_ld32(tq + (uint32_t)(buf * GW_DCOLS + c2), v);
#pragma unroll
      for (int i = 0; i < 32; ++i) tot2[i] += v[i];
      tmem_ld32(tq + (uint32_t)(buf * GW_DCOLS + 128), v);   // column 128 = the row sum (bias gradient), 129.. are zero
      btot += v[0];
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&dEmpty[buf]);
    }
    // matrix order of the accumulator slot (GACC_*): We_w, We_u, Wi_w, Wi_u, Wa_w, Wa_u; [out][in]
    float* ga = d.gacc + (size_t)blockIdx.x * GACC_N;
    const int m1 = q == 0 ? 0 : (q == 1 ? 2 : 4), m2 = m1 + 1;
    float4* o1 = reinterpret_cast<float4*>(ga + m1 * 1024 + lane * C);
    float4* o2 = reinterpret_cast<float4*>(ga + m2 * 1024 + lane * C);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o1[i] = make_float4(tot1[4 * i], tot1[4 * i + 1], tot1[4 * i + 2], tot1[4 * i + 3]);
      o2[i] = make_float4(tot2[4 * i], tot2[4 * i + 1], tot2[4 * i + 2], tot2[4 * i + 3]);
    }
    ga[GACC_BE + q * C + lane] = btot;
  } else {
    // ---- loaders + converters.  Raw fp32 tiles arrive by TMA: ONE cp.async.bulk.tensor per operand tensor and tile (a 32-channel x
    // 32-pixel box of the [rows][P] tensor map), no registers held, four tiles in flight per SM.  (Measured alternatives: LDG-fed
    // registers hold one tile, 1.25 ms, one load latency per tile; one cp.async.bulk per 128-byte row is bound by the TMA issue rate,
    // ~20 cycles per copy, 2.07 ms -- as the mma.sync version's 256-byte row copies were.)  Conversion: a warp owns 8 channels of 3 (or 2)
    // tensors; lane = (channel, 8-pixel chunk): 32 bytes of fp32 in, one hi and one lo 16-byte unit of the operand planes out.
    const int pw = warp < 4 ? warp - 1 : warp - 4;            // 0..11
    const int kc = lane & 3, r8 = lane >> 2;                  // 8-pixel chunk of the tile, row inside the warp's 8-row group
    const int row = (pw & 3) * 8 + r8;                        // converters 0-3: tensors 0, 3, 6; 4-7: tensors 1, 4; 8-11: tensors 2, 5; 8 channels each
    constexpr int AHEAD = GW_NRAW - 1;                         // tiles in flight; the slot of tile it + AHEAD is the one tile it - 1 used
    // Every converter warp waits for the slot (= for all warps to be through with tile it - 1), converter 0 issues: keeping the
    // warps in step was measured faster (1.00 ms) than letting them run ahead of each other (1.05 ms: a warp that is a tile ahead
    // takes issue slots from the ones the MMA warp is waiting for).
    auto issue = [&](int tile, int j) {                        // j = sequence number of the tile in this CTA
      const int slot = j % GW_NRAW;
      if (j >= GW_NRAW) mbar_wait(&rawEmpty[slot], (uint32_t)(((j / GW_NRAW) - 1) & 1));
      if (pw == 0 && lane < 7) {                                          // one tensor-map copy per operand tensor: [32 channels][32 pixels]
        const int t = tile / perT, tile_t = tile - t * perT;
        const int b = tile_t / tpi, pix_base = (tile_t - b * tpi) * GW_PX;
        if (lane == 0) mbar_expect_tx(&rawFull[slot], GW_RAW_STAGE);
        __syncwarp(0x7fu);
        tma_load_2d(smem_raw + (size_t)slot * GW_RAW_STAGE + lane * (C * GW_PX * 4), &maps.m[lane], pix_base, (t * d.B + b) * C, &rawFull[slot]);
      }
    };
    const int G = gridDim.x;
    for (int j = 0; j < AHEAD; ++j)
      if ((int)blockIdx.x + j * G < nTiles) issue(blockIdx.x + j * G, j);
    int it = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += G, ++it) {
      if (tile + AHEAD * G < nTiles) issue(tile + AHEAD * G, it + AHEAD);
      const int slot = it % GW_NRAW, st = it % GW_NOP;
      mbar_wait(&rawFull[slot], (uint32_t)((it / GW_NRAW) & 1));
      if (it >= GW_NOP) mbar_wait(&opEmpty[st], (uint32_t)(((it / GW_NOP) - 1) & 1));
      const uint8_t* raw = smem_raw + (size_t)slot * GW_RAW_STAGE + row * (GW_PX * 4) + kc * 32;
      uint8_t* sa = smem_raw + GW_OP0 + (size_t)st * GW_OP_STAGE;
      // lane = (row, 8-pixel chunk): 32 bytes in, one hi and one lo 16-byte unit out; the tensor index is a compile-time constant
#define GW_CONVERT(K)                                                                                                        \
      {                                                                                                                      \
        const float4 a4 = *reinterpret_cast<const float4*>(raw + (K) * (C * GW_PX * 4));                                      \
        const float4 b4 = *reinterpret_cast<const float4*>(raw + (K) * (C * GW_PX * 4) + 16);                                 \
        const float f[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};                                                  \
        uint32_t hi[4], lo[4];                                                                                               \
        _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                                      \
          split_bf16x2(f[2 * e], f[2 * e + 1], hi[e], lo[e]); \
        }                                                                                                                    \
        uint8_t* dst = (K) < 3 ? sa + kc * GW_PLANE_A + ((K) * 32 + row) * 16                                                 \
                               : sa + GW_A_BYTES + kc * GW_PLANE_B + (((K) - 3) * 32 + row) * 16;                             \
        *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);                                             \
        *reinterpret_cast<uint4*>(dst + ((K) < 3 ? GW_LO_A : GW_LO_B)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);             \
      }
      if (pw < 4) { GW_CONVERT(0) GW_CONVERT(3) GW_CONVERT(6) }
      else if (pw < 8) { GW_CONVERT(1) GW_CONVERT(4) }
      else { GW_CONVERT(2) GW_CONVERT(5) }
#undef GW_CONVERT
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(&opFull[st]); mbar_arrive(&rawEmpty[slot]); }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// returns the number of accumulator slots written (one per CTA) or a negative error code
typedef CUresult (*PtkEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PtkEncodeTiled encode_tiled_fn() {
  static PtkEncodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (PtkEncodeTiled)p;
  }();
  return fn;
}
static bool make_map(CUtensorMap* m, const float* base, long long rows, int P) {
  const cuuint64_t gdim[2] = {(cuuint64_t)P, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)P * 4};
  const cuuint32_t box[2] = {(cuuint32_t)GW_PX, (cuuint32_t)C};
  const cuuint32_t estr[2] = {1, 1};
  return encode_tiled_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
template <int CELL>
static cudaError_t launch_gw(const IntDev& d, int nT, int grid, const GwMaps& maps, cudaStream_t st) {
  cudaError_t e = ptk_func_smem(gate_wgrad_tc_kernel<CELL>, GW_SMEM);
  if (e == cudaSuccess) e = launch_pdl(true, gate_wgrad_tc_kernel<CELL>, dim3(grid), dim3(GW_THREADS), GW_SMEM, st, d, nT, maps);
  return e;
}
int launch_gate_wgrad_tc(const IntDev& d, int cell, int nT, cudaStream_t st) {
  const int P = d.H * d.W;
  const long long tiles = (long long)nT * d.B * (P / GW_PX);
  const int grid = tiles < GATE_CTAS ? (int)tiles : GATE_CTAS;
  if (!encode_tiled_fn()) { set_error("gate_wgrad_tc: cuTensorMapEncodeTiled not available from the driver"); return -1; }
  if (d.dy_tstride != (size_t)d.B * C * P || d.ck_stride != (size_t)d.B * C * P || (cell == PTK_CELL_HGRU && d.a_tstride != (long long)d.B * C * P)) {
    set_error("gate_wgrad_tc: per-step operand tensors are not laid out [T][B,C,P]");
    return -1;
  }
  // operand tensors of all steps are contiguous [T (or T+1)][B][C][P]: row (t, b, c) of a [rows][P] matrix
  const long long rT = (long long)nT * d.B * C, rT1 = (long long)(nT + 1) * d.B * C;
  const float* S = cell == PTK_CELL_INT ? d.I_at(0) : (cell == PTK_CELL_HGRU ? d.A_at(0) : d.E_at(0));
  const long long rS = cell == PTK_CELL_HGRU ? rT : rT1;
  GwMaps maps;
  const bool ok = make_map(&maps.m[0], d.geb_at(0), rT, P) && make_map(&maps.m[1], d.gi_at(0), rT, P) && make_map(&maps.m[2], d.a_at(0), rT, P) &&
                  make_map(&maps.m[3], S, rS, P) && make_map(&maps.m[4], d.GeB_at(0), rT, P) && make_map(&maps.m[5], d.X_at(0), rT, P) &&
                  make_map(&maps.m[6], d.E_at(0), rT1, P);
  if (!ok) { set_error("gate_wgrad_tc: cuTensorMapEncodeTiled failed"); return -1; }
  const cudaError_t e = cell == PTK_CELL_INT ? launch_gw<PTK_CELL_INT>(d, nT, grid, maps, st)
                      : cell == PTK_CELL_HGRU ? launch_gw<PTK_CELL_HGRU>(d, nT, grid, maps, st) : launch_gw<2>(d, nT, grid, maps, st);
  if (e != cudaSuccess) {
    set_error("gate_wgrad_tc_kernel launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return grid;
}

}  // namespace ptk
