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
#include <cuda_bf16.h>

#include "ptk_int.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int GW_THREADS = 384;                 // warp 0: MMA issuer; warps 4-6: flush (TMEM quadrants 0-2); the other 8: producers
constexpr int GW_PX = 64;                       // pixels per stage
constexpr int GW_KC = GW_PX / 8;                // 8-pixel chunks per stage
constexpr int GW_MA = 128, GW_NB = 144;
constexpr int GW_PLANE_A = GW_MA * 16 + 16;     // +16: the 8 chunk lanes of a quarter warp hit 8 bank groups
constexpr int GW_PLANE_B = GW_NB * 16 + 16;
constexpr int GW_LO_A = GW_KC * GW_PLANE_A + 64;     // lo planes follow the hi planes, skewed by 64 B: a hi / lo lane pair of the same chunk
constexpr int GW_LO_B = GW_KC * GW_PLANE_B + 64;     // then hits different bank groups
constexpr int GW_A_BYTES = GW_LO_A + GW_KC * GW_PLANE_A + 64;
constexpr int GW_B_BYTES = GW_LO_B + GW_KC * GW_PLANE_B + 64;
constexpr int GW_STAGE = GW_A_BYTES + GW_B_BYTES;
constexpr int GW_NST = 3;
constexpr int GW_SMEM = GW_NST * GW_STAGE + 256;
constexpr int GW_DCOLS = 160;                   // TMEM columns per accumulator buffer (144 used)

template <int CELL>
__global__ void __launch_bounds__(GW_THREADS, 1) gate_wgrad_tc_kernel(IntDev d, int nT) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + GW_NST * GW_STAGE);
  uint64_t* full = bars;                        // [GW_NST] stage filled (8 producer warps)
  uint64_t* empty = bars + GW_NST;              // [GW_NST] stage consumed (tcgen05.commit)
  uint64_t* dFull = bars + 2 * GW_NST;          // [2] a time step's sum is complete
  uint64_t* dEmpty = dFull + 2;                 // [2] ... and has been retired (3 flush warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dEmpty + 2);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int P = d.H * d.W, tpi = P / GW_PX, perT = d.B * tpi, nTiles = nT * perT;

  if (tid == 0) {
    for (int i = 0; i < GW_NST; ++i) { mbar_init(&full[i], 8); mbar_init(&empty[i], 1); }
    mbar_init(&dFull[0], 1); mbar_init(&dFull[1], 1);
    mbar_init(&dEmpty[0], 3); mbar_init(&dEmpty[1], 3);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows the producers never write: A rows 96-127 and B rows 129-143 stay zero, B row 128 is the bias column (ones)
  for (int i = tid; i < GW_NST * GW_STAGE / 16; i += GW_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (tid < GW_NST * GW_KC) {
    const int st = tid / GW_KC, kc = tid % GW_KC;
    *reinterpret_cast<uint4*>(smem_raw + (size_t)st * GW_STAGE + GW_A_BYTES + kc * GW_PLANE_B + 128 * 16) =
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
    const uint32_t sbase = smem_u32(smem_raw);
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
      const int st = it % GW_NST;
      mbar_wait(&full[st], (uint32_t)((it / GW_NST) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aH0 = sbase + (uint32_t)st * GW_STAGE, bH0 = aH0 + GW_A_BYTES;
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
      if (elect_one()) umma_commit(&empty[st]);
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
    // ---- producers: fp32 rows -> bf16 hi/lo chunk planes.  warp-item = (tensor k, 4 channels) = two load instructions of two full
    // 256-byte rows each (lane = 4 pixels of one row: every 32-byte sector is used whole).  A 16-byte unit of the operand planes is
    // 8 pixels, so a lane pair exchanges halves: the even lane stores the pair's hi unit, the odd lane the lo unit.
    const int pw = warp < 4 ? warp - 1 : warp - 4;            // 0..7
    const int px4 = lane & 15, rsel = lane >> 4, kc = px4 >> 1;
    const bool odd = (lane & 1) != 0;
    auto load_tile = [&](float4 (&v)[7][2], int tile) {
      const int t = tile / perT, tile_t = tile - t * perT;
      const int b = tile_t / tpi, pix_base = (tile_t - b * tpi) * GW_PX;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int item = pw * 7 + j, k = item >> 3, c0 = (item & 7) * 4 + rsel;
        const float* sp = k == 0 ? d.geb_at(t) : k == 1 ? d.gi_at(t) : k == 2 ? d.a_at(t)
                        : k == 3 ? ((CELL == PTK_CELL_INT) ? d.I_at(t) : (CELL == PTK_CELL_HGRU ? d.A_at(t) : d.E_at(t)))
                        : k == 4 ? d.GeB_at(t) : k == 5 ? d.X_at(t) : d.E_at(t);
        const float* src = sp + ((size_t)b * C + c0) * P + pix_base + px4 * 4;
        v[j][0] = __ldcs(reinterpret_cast<const float4*>(src));                      // channel c0
        v[j][1] = __ldcs(reinterpret_cast<const float4*>(src + (size_t)2 * P));      // channel c0 + 2
      }
    };
    auto store_tile = [&](const float4 (&v)[7][2], int it) {
      const int st = it % GW_NST;
      if (it >= GW_NST) mbar_wait(&empty[st], (uint32_t)(((it / GW_NST) - 1) & 1));
      uint8_t* sa = smem_raw + (size_t)st * GW_STAGE;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int item = pw * 7 + j, k = item >> 3, c0 = (item & 7) * 4 + rsel;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float f[4] = {v[j][h].x, v[j][h].y, v[j][h].z, v[j][h].w};
          uint32_t hi[2], lo[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * e]), h1 = __float2bfloat16_rn(f[2 * e + 1]);
            hi[e] = pack_bf16x2(h0, h1);
            lo[e] = pack_bf16x2(__float2bfloat16_rn(f[2 * e] - __bfloat162float(h0)), __float2bfloat16_rn(f[2 * e + 1] - __bfloat162float(h1)));
          }
          // even lane: keeps hi, receives the partner's hi;  odd lane: keeps lo, receives the partner's lo
          const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? hi[0] : lo[0], 1);
          const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? hi[1] : lo[1], 1);
          const uint4 unit = odd ? make_uint4(r0, r1, lo[0], lo[1]) : make_uint4(hi[0], hi[1], r0, r1);
          const int row = c0 + 2 * h;
          uint8_t* dst = k < 3 ? sa + kc * GW_PLANE_A + (k * 32 + row) * 16 + (odd ? GW_LO_A : 0)
                               : sa + GW_A_BYTES + kc * GW_PLANE_B + ((k - 3) * 32 + row) * 16 + (odd ? GW_LO_B : 0);
          *reinterpret_cast<uint4*>(dst) = unit;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    };
    // One register set: a second set with the next tile's loads in flight during the conversion was measured SLOWER (1.44 vs 1.25 ms
    // at the headline shape); the kernel sits at ~3.0 TB/s of operand traffic like its mma.sync predecessor.
    float4 v0[7][2];
    int it = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++it) {
      load_tile(v0, tile);
      store_tile(v0, it);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// returns the number of accumulator slots written (one per CTA) or a negative error code
template <int CELL>
static cudaError_t launch_gw(const IntDev& d, int nT, int grid, cudaStream_t st) {
  cudaError_t e = ptk_func_smem(gate_wgrad_tc_kernel<CELL>, GW_SMEM);
  if (e == cudaSuccess) e = launch_pdl(true, gate_wgrad_tc_kernel<CELL>, dim3(grid), dim3(GW_THREADS), GW_SMEM, st, d, nT);
  return e;
}
int launch_gate_wgrad_tc(const IntDev& d, int cell, int nT, cudaStream_t st) {
  const long long tiles = (long long)nT * d.B * (d.H * d.W / GW_PX);
  const int grid = tiles < GATE_CTAS ? (int)tiles : GATE_CTAS;
  const cudaError_t e = cell == PTK_CELL_INT ? launch_gw<PTK_CELL_INT>(d, nT, grid, st)
                      : cell == PTK_CELL_HGRU ? launch_gw<PTK_CELL_HGRU>(d, nT, grid, st) : launch_gw<2>(d, nT, grid, st);
  if (e != cudaSuccess) {
    set_error("gate_wgrad_tc_kernel launch failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return grid;
}

}  // namespace ptk
