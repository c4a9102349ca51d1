// Weight gradient of the 7x7 horizontal-connection convolutions, accumulated over ALL time steps in one launch
// (tcgen05, accumulators resident in TMEM for the whole kernel, A operand fed from TMEM).
//
//   dW[oc][ic][ky][kx] = sum_t sum_{b,y,x} dY_t[b,oc,y,x] * In_t[b,ic,y+ky-3,x+kx-3]          (models/InT.py:161,172 backward)
//
// Why a second kernel next to ptk_wgrad_tc.cu.  That kernel runs once per time step inside the reverse loop; its 49 x
// (64 x 64) accumulators do not fit in TMEM, so every launch read-add-writes a 200 KB slot per CTA (51 MB per launch
// through the L2 the point-wise kernels live in), and its M = 64 x N = 64 instruction runs at the shared-memory operand
// floor (4 KB per 32 cycles) while issuing all four hi/lo products.  Here:
//
//  * GEMM roles.  K = pixels of one image row segment (16 per instruction).  M = 128 = 4 kernel rows x 32 oc: lane
//    quadrant `sa` of the A operand holds dY shifted by sa rows.  N = 224 = 7 kernel columns x 32 ic: the In row is
//    stored in shared memory as 7 column-shifted copies (28 planes of 8 channels, MN-major no-swizzle).  One instruction
//    therefore accumulates 28 of the 49 taps' 32 x 32 blocks; two accumulator regions (ky 0-3, ky 4-6 + one unused) of
//    224 columns = 448 TMEM columns hold the complete gradient of this CTA for the whole kernel: no slot traffic.
//  * Three products (dYh*Inh, dYh*Inl, dYl*Inh) accumulate into the SAME columns, so nothing is wasted on lo*lo.
//  * The A operand (dY, bf16 hi / lo, K-major = lane per output channel) is written to TMEM by four warps with
//    tcgen05.st and consumed by `tcgen05.mma [d], [a], b_desc`: shared memory only feeds B (7 KB per 112-cycle
//    instruction = 64 B/clk), so the MMA stream is tensor-bound, not operand-bound.
//  * Persistent CTAs (one per SM) walk (t, sample, tile) work items; the gradient leaves TMEM once, at the end, into
//    the CTA's slot [49][32 oc][32 ic]; `int_finalize_kernel` folds the <= 148 slots in double precision in a fixed
//    order (deterministic).
//
// Pipeline per CTA (512 threads):  loaders (8 warps, two groups alternating rows): global fp32 -> bf16 hi/lo, In row ->
// 7 shifted copies in a 4-slot ring, dY row -> packed [oc][16 px] images in a 12-row ring;  A writers (4 warps, one per
// TMEM lane quadrant): dY ring -> tcgen05.st into a 4-slot ring of 16 TMEM columns;  MMA issuer (1 warp);  3 warps issue
// L2 prefetches for the CTA's next tile.  All hand-offs are mbarriers; the MMA side releases slots with tcgen05.commit.
#include "ptk_common.cuh"
#include "ptk_kernels.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int TS_TH = 16, TS_TW = 32, TS_ROWS = TS_TH + KS - 1;   // tile: 16 x 32 output pixels, 22 input rows
constexpr int TS_PLANE = TS_TW * 16 + 16;      // 528 B: one (kx, 8-channel chunk) plane of an In row; +16 staggers the banks
constexpr int TS_NPL = KS * 4;                 // 28 planes = N / 8
constexpr int TS_IMG = TS_NPL * TS_PLANE;      // one precision (hi or lo) of one In row: 14 784 B
constexpr int TS_SLOT = 2 * TS_IMG;            // hi image + lo image
constexpr int TS_NI = 4;                       // In-row ring
constexpr int TS_DYROW = 2 * 2 * C * 32;       // [hi|lo][half][oc][16 px bf16 = 32 B] = 4096 B
constexpr int TS_ND = 12;                      // dY-row ring (a row is live for 7 consecutive In rows)
constexpr int TS_NA = 4;                       // A-operand ring in TMEM: 4 x (8 hi + 8 lo columns)
constexpr int TS_DCOLS = 2 * KS * C;           // 448 accumulator columns
constexpr int TS_IN_BYTES = TS_NI * TS_SLOT;
constexpr int TS_DY_BYTES = TS_ND * TS_DYROW;
constexpr int TS_NBAR = 2 * TS_NI + 2 * TS_ND + 2 * TS_NA + 1;
constexpr int TS_SMEM = TS_IN_BYTES + TS_DY_BYTES + TS_NBAR * 8 + 32;
constexpr int TS_THREADS = 512;
constexpr int TS_MAX_CTAS = 148;

// kind::f16 instruction descriptor: D = f32, A = B = bf16, A K-major (TMEM), B MN-major, M = 128, N = 224
__host__ __device__ constexpr uint32_t ts_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(TS_NPL * 8 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// the (t, sample, tile) work item list of one CTA and the In rows of a tile that lie inside the image
struct TsWalk {
  int H, W, tilesX, tpi, per_t, NT;
  __device__ __forceinline__ TsWalk(const WgradTsArgs& a)
      : H(a.H), W(a.W), tilesX(a.W / TS_TW), tpi((a.H / TS_TH) * (a.W / TS_TW)), per_t(a.B * tpi), NT(a.T * a.B * tpi) {}
  struct Tile { int t, b, y0, x0, rlo, rhi; };
  __device__ __forceinline__ Tile tile(int w) const {
    Tile q;
    q.t = w / per_t;
    const int r = w - q.t * per_t;
    q.b = r / tpi;
    const int ti = r - q.b * tpi;
    q.y0 = (ti / tilesX) * TS_TH;
    q.x0 = (ti % tilesX) * TS_TW;
    q.rlo = 3 - q.y0 > 0 ? 3 - q.y0 : 0;                             // In row r <-> image row y0 - 3 + r
    q.rhi = H + 2 - q.y0 < TS_ROWS - 1 ? H + 2 - q.y0 : TS_ROWS - 1;
    return q;
  }
};
// accumulator region g (kernel rows 4g .. 4g+3) has work on In row r iff some dY row r - 4g - sa is inside the tile
__device__ __forceinline__ bool ts_region_active(int g, int r) { return g == 0 ? (r <= TS_TH + 2) : (r >= 4); }

template <bool SINGLE>
__global__ void __launch_bounds__(TS_THREADS, 1) wgrad7_ts_kernel(WgradTsArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sIn = smem_raw;
  uint8_t* sDy = smem_raw + TS_IN_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + TS_IN_BYTES + TS_DY_BYTES);
  uint64_t* inFull = bars;
  uint64_t* inEmpty = inFull + TS_NI;
  uint64_t* dyFull = inEmpty + TS_NI;
  uint64_t* dyEmpty = dyFull + TS_ND;
  uint64_t* aFull = dyEmpty + TS_ND;
  uint64_t* aEmpty = aFull + TS_NA;
  uint64_t* doneBar = aEmpty + TS_NA;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(doneBar + 1);
  volatile int* sTileNow = reinterpret_cast<volatile int*>(tmem_slot + 1);   // tile index the MMA warp has reached (prefetch pacing)

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const TsWalk wk(a);
  const int H = a.H, W = a.W;
  const size_t bs = (size_t)C * H * W, plane = (size_t)H * W;

  if (tid == 0) {
    for (int i = 0; i < TS_NI; ++i) { mbar_init(&inFull[i], 4); mbar_init(&inEmpty[i], 1); }
    for (int i = 0; i < TS_ND; ++i) { mbar_init(&dyFull[i], 2); mbar_init(&dyEmpty[i], 4); }
    for (int i = 0; i < TS_NA; ++i) { mbar_init(&aFull[i], 4); mbar_init(&aEmpty[i], 1); }
    mbar_init(doneBar, 1);
    *sTileNow = -1;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the halo positions of the shifted copies are never written for 32-wide frames: they must read as zero
  for (int i = tid; i < TS_IN_BYTES / 16; i += TS_THREADS) reinterpret_cast<uint4*>(sIn)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    // ================================ loaders =====================================================================
    const int grp = (warp - 8) >> 2, q = (warp - 8) & 3;     // group 0/1 (alternating rows), chunk / half index
    int inCount = 0, k = 0;
    for (int w = blockIdx.x; w < wk.NT; w += gridDim.x, ++k) {
      const TsWalk::Tile tl = wk.tile(w);
      const float* inT = a.in + (size_t)tl.t * a.in_ts + (size_t)tl.b * bs;
      const float* dyT = a.dy + (size_t)tl.t * a.dy_ts + (size_t)tl.b * bs;
      for (int r = 0; r < TS_ROWS; ++r) {
        if (r >= tl.rlo && r <= tl.rhi) {
          if ((inCount & 1) == grp) {
            // ---- In row r: channels 8q .. 8q+7 of pixel x0 + lane -> 7 column-shifted copies (hi and lo images)
            const int gy = tl.y0 - 3 + r;
            const float* p0 = inT + ((size_t)(q * 8) * H + gy) * W + tl.x0 + lane;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __ldg(p0 + (size_t)e * plane);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
            }
            const int slot = inCount % TS_NI;
            if (inCount >= TS_NI) mbar_wait(&inEmpty[slot], (uint32_t)(((inCount / TS_NI) - 1) & 1));
            uint8_t* img = sIn + slot * TS_SLOT + q * TS_PLANE;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
              const int pp = lane + 3 - kx;      // plane kx holds In[x0 + pp + kx - 3] at position pp
              if (pp >= 0 && pp < TS_TW) {
                uint8_t* dst = img + kx * 4 * TS_PLANE + pp * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(dst + TS_IMG) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
            if (W > TS_TW && lane < 6) {
              // frames wider than a tile: the 3 + 3 halo columns carry real neighbours (or the zero border)
              const int hc = lane < 3 ? lane - 3 : TS_TW + lane - 3;      // -3..-1, 32..34
              const int gx = tl.x0 + hc;
              uint32_t hh[4] = {0, 0, 0, 0}, ll[4] = {0, 0, 0, 0};
              if (gx >= 0 && gx < W) {
                const float* ph = inT + ((size_t)(q * 8) * H + gy) * W + gx;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  split_bf16x2(__ldg(ph + (size_t)(2 * e) * plane), __ldg(ph + (size_t)(2 * e + 1) * plane), hh[e], ll[e]);
                }
              }
#pragma unroll
              for (int kx = 0; kx < KS; ++kx) {
                const int pp = hc + 3 - kx;
                if (pp >= 0 && pp < TS_TW) {
                  uint8_t* dst = img + kx * 4 * TS_PLANE + pp * 16;
                  *reinterpret_cast<uint4*>(dst) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                  *reinterpret_cast<uint4*>(dst + TS_IMG) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tcgen05.mma operand reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&inFull[slot]);
          }
          ++inCount;
        }
        if (r < TS_TH) {
          const int dseq = k * TS_TH + r;
          if ((dseq & 1) == grp && q < 2) {
            // ---- dY row r, half q: lane = output channel, 16 consecutive pixels -> 8 hi + 8 lo packed words
            const float* p0 = dyT + ((size_t)lane * H + tl.y0 + r) * W + tl.x0 + 16 * q;
            float4 v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __ldg(reinterpret_cast<const float4*>(p0) + e);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split_bf16x2(v[e].x, v[e].y, hi[2 * e], lo[2 * e]);
              split_bf16x2(v[e].z, v[e].w, hi[2 * e + 1], lo[2 * e + 1]);
            }
            const int slot = dseq % TS_ND;
            if (dseq >= TS_ND) mbar_wait(&dyEmpty[slot], (uint32_t)(((dseq / TS_ND) - 1) & 1));
            // 32 B per lane; the two 16-byte halves are swapped for lanes 4-7 of every 8 so that a quarter warp's
            // 16-byte accesses cover all 32 banks
            uint8_t* dst = sDy + slot * TS_DYROW + q * (C * 32) + lane * 32;
            const int sw = ((lane >> 2) & 1) * 16;
            *reinterpret_cast<uint4*>(dst + sw) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(dst + (sw ^ 16)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(dst + 2 * C * 32 + sw) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(dst + 2 * C * 32 + (sw ^ 16)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&dyFull[slot]);
          }
        }
      }
    }
  } else if (warp < 4) {
    // ================================ A writers (TMEM lane quadrant = warp) =======================================
    const int sa = warp;
    const uint32_t a_lane = tmem_base + ((uint32_t)(sa * 32) << 16) + (uint32_t)TS_DCOLS;
    int aCount = 0, k = 0;
    for (int w = blockIdx.x; w < wk.NT; w += gridDim.x, ++k) {
      const TsWalk::Tile tl = wk.tile(w);
      for (int r = 0; r < TS_ROWS; ++r) {
        if (r >= tl.rlo && r <= tl.rhi) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (!ts_region_active(g, r)) continue;
            const int j = r - 4 * g - sa;                       // dY row paired with In row r for kernel row 4g + sa
            const bool jv = j >= 0 && j < TS_TH && 4 * g + sa < KS;
            const int dseq = k * TS_TH + j, dslot = jv ? dseq % TS_ND : 0;
            if (jv) mbar_wait(&dyFull[dslot], (uint32_t)((dseq / TS_ND) & 1));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t hi[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              if (jv) {
                const uint8_t* src = sDy + dslot * TS_DYROW + h * (C * 32) + lane * 32;
                const int sw = ((lane >> 2) & 1) * 16;
                const uint4 h0 = *reinterpret_cast<const uint4*>(src + sw), h1 = *reinterpret_cast<const uint4*>(src + (sw ^ 16));
                hi[0] = h0.x; hi[1] = h0.y; hi[2] = h0.z; hi[3] = h0.w; hi[4] = h1.x; hi[5] = h1.y; hi[6] = h1.z; hi[7] = h1.w;
                if (!SINGLE) {
                  const uint4 l0 = *reinterpret_cast<const uint4*>(src + 2 * C * 32 + sw);
                  const uint4 l1 = *reinterpret_cast<const uint4*>(src + 2 * C * 32 + (sw ^ 16));
                  lo[0] = l0.x; lo[1] = l0.y; lo[2] = l0.z; lo[3] = l0.w; lo[4] = l1.x; lo[5] = l1.y; lo[6] = l1.z; lo[7] = l1.w;
                }
              }
              const int slot = aCount % TS_NA;
              if (aCount >= TS_NA) mbar_wait(&aEmpty[slot], (uint32_t)(((aCount / TS_NA) - 1) & 1));
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              tmem_st8(a_lane + (uint32_t)(slot * 16), hi);
              if (!SINGLE) tmem_st8(a_lane + (uint32_t)(slot * 16 + 8), lo);
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive(&aFull[slot]);
              ++aCount;
            }
          }
        }
        const int jr = r - (KS - 1);          // every use of dY row r - 6 is behind this warp now
        if (jr >= 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&dyEmpty[(k * TS_TH + jr) % TS_ND]);
        }
      }
    }
    // ---- epilogue: the CTA's complete gradient leaves TMEM once -> slot [49][32 oc][32 ic] -----------------------
    mbar_wait(doneBar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* slotp = a.wacc + (size_t)blockIdx.x * TAPS * C * C;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      const int ky = 4 * g + sa;
#pragma unroll 1
      for (int kx = 0; kx < KS; ++kx) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(sa * 32) << 16) + (uint32_t)(g * KS * C + kx * C), v);   // lane = oc, v[ic]
        if (ky < KS) {
          float* dst = slotp + ((size_t)(ky * KS + kx) * C + lane) * C;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  } else if (warp == 4) {
    // ================================ MMA issuer ==================================================================
    constexpr uint32_t IDESC = ts_idesc();
    const uint64_t dB0 = umma_desc(smem_u32(sIn), 128, TS_PLANE);      // LBO: next 8-pixel K group, SBO: next 8-channel plane
    const uint32_t bHiW = (uint32_t)(dB0 >> 32), bLoW = (uint32_t)dB0;
    int inCount = 0, aCount = 0, k = 0;
    uint32_t used0 = 0, used1 = 0;             // has accumulator region g been written yet
    for (int w = blockIdx.x; w < wk.NT; w += gridDim.x, ++k) {
      const TsWalk::Tile tl = wk.tile(w);
      if (lane == 0) *sTileNow = k;
      for (int r = tl.rlo; r <= tl.rhi; ++r) {
        const int islot = inCount % TS_NI;
        mbar_wait(&inFull[islot], (uint32_t)((inCount / TS_NI) & 1));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (!ts_region_active(g, r)) continue;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int aslot = aCount % TS_NA;
            mbar_wait(&aFull[aslot], (uint32_t)((aCount / TS_NA) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dcol = tmem_base + (uint32_t)(g * KS * C);
            const uint32_t acol = tmem_base + (uint32_t)(TS_DCOLS + aslot * 16);
            const uint32_t bOff = bLoW + (uint32_t)((islot * TS_SLOT + h * 256) >> 4);
            const uint64_t dBh = ((uint64_t)bHiW << 32) | bOff;
            const uint64_t dBl = ((uint64_t)bHiW << 32) | (bOff + (uint32_t)(TS_IMG >> 4));
            const uint32_t acc0 = g == 0 ? used0 : used1;
            if (elect_one()) {
              umma_ts_bf16(dcol, acol, dBh, IDESC, acc0);
              if (!SINGLE) {
                umma_ts_bf16(dcol, acol, dBl, IDESC, 1u);
                umma_ts_bf16(dcol, acol + 8, dBh, IDESC, 1u);
              }
              umma_commit(&aEmpty[aslot]);
            }
            __syncwarp();
            if (g == 0) used0 = 1u; else used1 = 1u;
            ++aCount;
          }
        }
        if (elect_one()) umma_commit(&inEmpty[islot]);
        __syncwarp();
        ++inCount;
      }
    }
    if (elect_one()) umma_commit(doneBar);
    __syncwarp();
  } else {
    // ================================ L2 prefetch of the CTA's next tile (warps 5-7) ==============================
    const int pt = tid - 5 * 32;
    int k = 0;
    for (int w = blockIdx.x; w < wk.NT; w += gridDim.x, ++k) {
      const int wn = w + gridDim.x;
      if (wn >= wk.NT) break;
      while (*sTileNow < k) __nanosleep(256);
      const TsWalk::Tile tl = wk.tile(wn);
      const float* inT = a.in + (size_t)tl.t * a.in_ts + (size_t)tl.b * bs;
      const float* dyT = a.dy + (size_t)tl.t * a.dy_ts + (size_t)tl.b * bs;
      const int nin = C * (tl.rhi - tl.rlo + 1);
      for (int i = pt; i < nin + C * TS_TH; i += 96) {
        if (i < nin) {
          const int ch = i / (tl.rhi - tl.rlo + 1), r = tl.rlo + i % (tl.rhi - tl.rlo + 1);
          prefetch_l2(inT + ((size_t)ch * H + tl.y0 - 3 + r) * W + tl.x0);
        } else {
          const int i2 = i - nin, ch = i2 / TS_TH, r = i2 % TS_TH;
          prefetch_l2(dyT + ((size_t)ch * H + tl.y0 + r) * W + tl.x0);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

int wgrad_ts_ctas(int T, int B, int H, int W) {
  const long long nt = (long long)T * B * (H / TS_TH) * (W / TS_TW);
  return (int)(nt < TS_MAX_CTAS ? nt : TS_MAX_CTAS);
}
size_t wgrad_ts_acc_floats(int T, int B, int H, int W) { return (size_t)wgrad_ts_ctas(T, B, H, W) * TAPS * C * C; }

// stand-alone fold of the slots (ptk_wgrad7; the InT backward folds them in int_finalize_kernel): dW[oc][ic][ky][kx]
__global__ void __launch_bounds__(256) wgrad_ts_fold_kernel(const float* __restrict__ wacc, int nslots, float* __restrict__ dw) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C * C * TAPS) return;
  const int oc = j / (C * TAPS), ic = (j / TAPS) % C, tap = j % TAPS;
  double s = 0;
  for (int g = 0; g < nslots; ++g) s += (double)wacc[((size_t)g * TAPS + tap) * C * C + oc * C + ic];
  dw[j] = (float)s;
}
int launch_wgrad_ts_fold(const float* wacc, int nslots, float* dw, cudaStream_t st) {
  wgrad_ts_fold_kernel<<<(C * C * TAPS + 255) / 256, 256, 0, st>>>(wacc, nslots, dw);
  PTK_LAUNCH_OK("wgrad_ts_fold_kernel");
  return PTK_OK;
}

int launch_wgrad7_ts(const WgradTsArgs& a, cudaStream_t st) {
  const dim3 grid(wgrad_ts_ctas(a.T, a.B, a.H, a.W));
  if (a.mode == PTK_CONV_BF16) {
    PTK_CUDA_OK(ptk_func_smem(wgrad7_ts_kernel<true>, TS_SMEM));
    PTK_CUDA_OK(launch_pdl(a.pdl != 0, wgrad7_ts_kernel<true>, grid, dim3(TS_THREADS), TS_SMEM, st, a));
  } else {
    PTK_CUDA_OK(ptk_func_smem(wgrad7_ts_kernel<false>, TS_SMEM));
    PTK_CUDA_OK(launch_pdl(a.pdl != 0, wgrad7_ts_kernel<false>, grid, dim3(TS_THREADS), TS_SMEM, st, a));
  }
  PTK_LAUNCH_OK("wgrad7_ts_kernel");
  return PTK_OK;
}

}  // namespace ptk
