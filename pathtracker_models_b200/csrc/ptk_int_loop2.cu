// Persistent forward time loop, second generation: the six 1x1 gate convolutions run on tcgen05 as well.
//
// Reference being replaced: the loop of models/InT.py:223-235 around rCell.forward (:145-179) / hConvGRUCell.forward
// (models/ffhgru_hierarchy.py:135-173).  ptk_int_loop.cu showed that merely chaining the chain's phases inside one cooperative
// kernel gains nothing (82 us per frame either way: the point-wise phases are bound by their legacy mma.sync / FFMA gate products
// and by one warp-tile's latency chain, not by launches).  This kernel uses what only a resident CTA can do:
//
//  * thread = pixel.  The CTA's 16 x 32-pixel tile is 512 pixels = 512 threads; warp w owns image row w of the tile, which is
//    TMEM lane quadrant w % 4 of M-tile w / 4 -- so a thread reads its own pixel's gate pre-activations straight out of TMEM
//    (tcgen05.ld, 32 channels = 32 registers) and all global accesses are 128-byte coalesced rows.
//  * gate products = tcgen05.mma with the pixels as the A operand (M = 128, K-major bf16 hi/lo chunk planes staged by the pixel's
//    own thread), the gate matrices as B (N = 32 / 64, bf16 hi/lo, 3-pass split like the 7x7 convolutions), fp32 accumulators in
//    the TMEM columns the convolution is not using at that moment.  Per frame 72 + 48 small instructions (< 1 us of tensor time)
//    replace ~2 300 mma.sync per warp.
//  * the new state goes straight into the NEXT convolution's operand image: the thread that computes I_{t+1} (resp. Ge_{t+1}) of a
//    pixel converts it to bf16 hi/lo and stores it at the pixel's place in the zero-padded image, so the convolution only fetches
//    its 3 + 3 halo rows from the neighbouring tile (through L2, after the per-sample barrier) instead of staging 22 rows.
//  * BatchNorm: split arrive / wait grid barriers as in ptk_int_loop.cu; everything that does not need the batch statistics
//    (mid: excitation gate, stem, inhibition-gate pre-activation; post: stem and Wa_w X of the next frame) sits before the wait.
//
// Arithmetic differs from the chain only in the gate products (bf16 hi/lo 3-pass on tcgen05 instead of tf32 hi/lo 3-pass on
// mma.sync): same parity gate (tests/test_loop_gpu.py runs the reference fixtures through this kernel).
#define PTK_LOOP_STAMPS_TU
#include "ptk_int_loop.cuh"

namespace ptk {

constexpr int G_PX = T2_TH * T2_TW;            // 512 pixels of the CTA tile = 4 M-tiles of 128
constexpr int G_PLANE = G_PX * 16;             // one 8-channel chunk plane: [pixel][8 bf16]
constexpr int G_IMG = 8 * G_PLANE;             // hi planes 0-3, lo planes 4-7: 64 KB
// point-wise view of the shared memory (the convolution's view is LP_RING / LP_IMG / LP_PART of ptk_int_loop.cuh)
constexpr int L2_IMG0 = 0;                     // mid: S;  post: X_{t+1}, then E_{t+1}      (= the weight ring: outside the conv image)
constexpr int L2_IMG1 = G_IMG;                 // mid: Ge
constexpr int L2_IMG2 = 2 * G_IMG;             // mid: X
constexpr int L2_GW = 3 * G_IMG;               // gate matrices of the phase (mid 16 KB, post 8 KB), inside the quadrant-tile area
constexpr int L2_FOLD = 0;                     // fold scratch double [8][32][3]: image 0 is dead whenever a fold runs
static_assert(L2_GW >= LP_PART && L2_GW + 16384 <= LP_BARS, "gate matrices must sit outside the conv image, under the barriers");
// gate matrices as B operands, bf16 hi/lo: block = [hi: 4 k-chunks][N rows][8 k] [lo: same]
constexpr int GW_MID_S = 0, GW_MID_GE = 8192, GW_MID_X = 12288, GW_MID_BYTES = 16384;      // S: N = 64 ([We_w ; Wi_u]), Ge: We_u, X: Wi_w
constexpr int GW_POST = 16384, GW_POST_X = 0, GW_POST_E = 4096, GW_POST_BYTES = 8192;        // X: Wa_w, E: Wa_u
constexpr int GW_TOTAL = GW_POST + GW_POST_BYTES;

// gate matrices ([out][in] fp32, parameter order a_w a_u i_w i_u e_w e_u) -> the B-operand blocks above
__global__ void prep_gates_tc_kernel(IntGateSrc g, uint8_t* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (block, n, k)
  if (i >= 6 * C * C) return;
  const int blk = i / (C * C), n = (i / C) % C, k = i % C;
  // block -> (source matrix, byte offset of the hi half, rows N of the stacked operand, first row)
  const int src[6] = {4, 3, 5, 2, 0, 1};
  const int off[6] = {GW_MID_S, GW_MID_S, GW_MID_GE, GW_MID_X, GW_POST + GW_POST_X, GW_POST + GW_POST_E};
  const int N[6] = {64, 64, 32, 32, 32, 32};
  const int n0[6] = {0, 32, 0, 0, 0, 0};
  const float v = g.w[src[blk]][n * C + k];
  const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
  const size_t e = (size_t)off[blk] + ((size_t)(k >> 3) * N[blk] + n0[blk] + n) * 16 + (size_t)(k & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(dst + e) = h;
  *reinterpret_cast<__nv_bfloat16*>(dst + e + 64 * N[blk]) = l;
}

// one pixel's 32 channels -> bf16 hi/lo chunk planes of a gate-GEMM image (A operand, K-major: [chunk][pixel][8 ch])
__device__ __forceinline__ void stage_px(uint8_t* img, const float (&v)[32], int p) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = v[q * 8 + 2 * e], v1 = v[q * 8 + 2 * e + 1];
      split_bf16x2(v0, v1, hi[e], lo[e]);
    }
    *reinterpret_cast<uint4*>(img + q * G_PLANE + p * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + (4 + q) * G_PLANE + p * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
// 8 channels (chunk q) of one pixel -> the convolution's operand image at padded pixel pc
__device__ __forceinline__ void stage_px_conv(uint8_t* sA, const float (&v)[8], int q, int pc) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float v0 = v[2 * e], v1 = v[2 * e + 1];
    split_bf16x2(v0, v1, hi[e], lo[e]);
  }
  *reinterpret_cast<uint4*>(sA + q * T2_PLANE + pc * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(sA + (4 + q) * T2_PLANE + pc * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// D[pixel][n] (+)= img[pixel][k] * Wblock[n][k]^T for the CTA's four M-tiles; 3-pass split (Ah Bh + Ah Bl + Al Bh).
// Called by one converged warp; `first` = the very first instruction into these accumulator columns overwrites.
__device__ __forceinline__ void gate_mma(uint32_t img, uint32_t wblk, int N, uint32_t dcol0, int dstride, bool first) {
  const uint64_t dA0 = umma_desc(0, G_PLANE, 128);            // pixels: LBO = next 8-channel plane, SBO = next 8 pixels
  const uint64_t dB0 = umma_desc(0, (uint32_t)N * 16, 128);   // weights: LBO = next k-chunk (N rows x 16 B), SBO = next 8 rows
  const uint32_t aW = (uint32_t)(dA0 >> 32), bW = (uint32_t)(dB0 >> 32);
  const uint32_t idesc = umma_idesc(N);
#pragma unroll 1
  for (int m = 0; m < 4; ++m) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const uint32_t aH = (uint32_t)dA0 + ((img + (uint32_t)(2 * s) * G_PLANE + (uint32_t)m * 2048) >> 4);
      const uint32_t aL = aH + ((4u * G_PLANE) >> 4);
      const uint32_t bH = (uint32_t)dB0 + ((wblk + (uint32_t)s * (uint32_t)(2 * N * 16)) >> 4);
      const uint32_t bL = bH + (uint32_t)((64 * N) >> 4);
      const uint64_t dAh = ((uint64_t)aW << 32) | aH, dAl = ((uint64_t)aW << 32) | aL;
      const uint64_t dBh = ((uint64_t)bW << 32) | bH, dBl = ((uint64_t)bW << 32) | bL;
      const uint32_t dcol = dcol0 + (uint32_t)(m * dstride);
      if (elect_one()) {
        umma_bf16(dcol, dAh, dBh, idesc, (first && s == 0) ? 0u : 1u);
        umma_bf16(dcol, dAh, dBl, idesc, 1u);
        umma_bf16(dcol, dAl, dBh, idesc, 1u);
      }
    }
  }
  __syncwarp();
}

template <int CELL, int ACT, bool SINGLE>
__global__ void __launch_bounds__(LP_THREADS, 1) int_loop2_kernel(IntDev d, LoopArgs la, const uint8_t* __restrict__ gateB) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sA = smem_raw + LP_IMG;                     // the convolution's operand image
  double (*sFin)[C][3] = reinterpret_cast<double (*)[C][3]>(smem_raw + L2_FOLD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + LP_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T2_NST + 6);
  uint64_t* gbar = bars + 2 * T2_NST + 8;              // gate MMAs complete
  uint64_t* wbar = gbar + 1;                           // gate matrices landed
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sbi[C], sbe[C], sba[C], sAl[C], sMu[C], sKa[C], sGa[C], sWr[C], sScale[C], sShift[C], sMerge[C * 4];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int H = d.H, W = d.W, P = H * W;
  const int x0 = blockIdx.x * T2_TW, y0 = blockIdx.y * T2_TH, b = blockIdx.z;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int nCta = gridDim.x * gridDim.y * gridDim.z;

  if (tid == 0) {
    uint64_t* full = bars;
    uint64_t* empty = bars + T2_NST;
    uint64_t* dFull = bars + 2 * T2_NST;
    uint64_t* dEmpty = dFull + 2;
    for (int i = 0; i < T2_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&dFull[0], 1); mbar_init(&dFull[1], 1);
    mbar_init(&dEmpty[0], T2_EPI_WARPS); mbar_init(&dEmpty[1], T2_EPI_WARPS);
    mbar_init(dEmpty + 2, LP_THREADS); mbar_init(dEmpty + 3, LP_THREADS - 64);
    mbar_init(gbar, 1); mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < C * 3; i += LP_THREADS) sWp[i] = __ldg(d.Wp + i);
  if (tid < C) {
    const int c = tid;
    sbp[c] = __ldg(d.bp + c);
    sba[c] = __ldg(d.gate_b[0] + c) + __ldg(d.gate_b[1] + c);
    sbi[c] = __ldg(d.gate_b[2] + c) + __ldg(d.gate_b[3] + c);
    sbe[c] = __ldg(d.gate_b[4] + c) + __ldg(d.gate_b[5] + c);
    sAl[c] = __ldg(d.alpha + c); sMu[c] = __ldg(d.mu + c);
    sKa[c] = __ldg(d.kappa + c); sGa[c] = __ldg(d.gamma + c);
    sWr[c] = __ldg(d.Wr + c);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // this thread's pixel: row `warp`, column `lane` of the CTA's tile
  const int p = warp * 32 + lane;                                   // pixel index inside the tile (gate-GEMM row)
  const int gy = y0 + warp, gx = x0 + lane;
  const int pix = gy * W + gx;
  const size_t base = (size_t)b * C * P + pix;
  const int pc = (warp + 3) * T2_WP + 3 + lane;                     // the pixel's place in the padded conv image
  const uint32_t tq = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);   // TMEM lane quadrant of this warp
  const int mt = warp >> 2;                                         // M-tile
  const uint32_t simg = smem_u32(smem_raw), sgw = smem_u32(smem_raw + L2_GW);
  const float br = __ldg(d.br);
  uint32_t gph = 0;                                                 // completed phases of gbar
  unsigned gtarget = 0, starget = 0;
  unsigned* sctr = la.sctr + b;
  const unsigned tps = (unsigned)la.tiles_per_sample;
  LP_STAMP_DECL

#pragma unroll 1
  for (int t = 0; t < d.T; ++t) {
    // ================= conv0: Ge_t -> Y0_t ==================================================================================
    loop_conv<SINGLE>(d.Ge, reinterpret_cast<const uint8_t*>(d.w_inh_f), d.Y0_at(t), la.part[0], cta, H, W, x0, y0, b, smem_raw,
                      sMerge, tmem_base, 0u, t > 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    LP_STAMP(0);
    cta_arrive(la.gctr);
    LP_STAMP(1);
    gtarget += (unsigned)nCta;
    // ================= mid, statistics-independent half ===================================================================
    float xin[3];
    {
      if (tid == 0) {
        mbar_expect_tx(wbar, GW_MID_BYTES);
        bulk_g2s(smem_raw + L2_GW, gateB, GW_MID_BYTES, wbar);
      }
      const float* Sg = (CELL == PTK_CELL_INT) ? d.I_at(t) + base : d.A_at(t) + (size_t)b * d.a_bstride + pix;
      float v[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = __ldcg(Sg + (size_t)c * P);
      stage_px(smem_raw + L2_IMG0, v, p);
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = __ldcg(d.Ge + base + (size_t)c * P);
      stage_px(smem_raw + L2_IMG1, v, p);
#pragma unroll
      for (int i = 0; i < 3; ++i) xin[i] = d.fv.at(b, i, t, gy, gx);
#pragma unroll
      for (int c = 0; c < 32; ++c)
        v[c] = actf<ACT>(fmaf(sWp[c * 3 + 2], xin[2], fmaf(sWp[c * 3 + 1], xin[1], fmaf(sWp[c * 3], xin[0], sbp[c]))));
      stage_px(smem_raw + L2_IMG2, v, p);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
      mbar_wait(wbar, 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // columns m*64 + [0,32): We_w S + We_u Ge (excitation gate, eq. 9);  m*64 + [32,64): Wi_u S + Wi_w X (inhibition gate, eq. 7)
      gate_mma(simg + L2_IMG0, sgw + GW_MID_S, 64, tmem_base, 64, true);
      gate_mma(simg + L2_IMG1, sgw + GW_MID_GE, 32, tmem_base, 64, false);
      gate_mma(simg + L2_IMG2, sgw + GW_MID_X, 32, tmem_base + 32, 64, false);
      if (elect_one()) umma_commit(gbar);
      __syncwarp();
    }
    mbar_wait(gbar, gph & 1u);
    ++gph;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      float g[32];
      tmem_ld32(tq + (uint32_t)(mt * 64), g);
      float* Gx = d.Gex_at(t) + base;
#pragma unroll
      for (int c = 0; c < 32; ++c) Gx[(size_t)c * P] = sigm<ACT>(g[c] + sbe[c]);
    }
    float gi[32];
    tmem_ld32(tq + (uint32_t)(mt * 64 + 32), gi);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // the gate images are dead: zero the convolution's operand image, the tail below writes the tile's own rows into it
    for (int i = tid; i < T2_A_BYTES / 16; i += LP_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    LP_STAMP(2);
    // ---- BatchNorm0: all convolutions of the batch are in ----
    cta_wait(la.gctr, gtarget);
    LP_STAMP(3);
    loop_fold(la.part[0], nCta, T2_TH * T2_TW, sFin, sScale, sShift, d.bn_w[0], d.bn_b[0], cta == 0 ? d.stat(t, 0) : nullptr);
    LP_STAMP(4);
    // ================= mid, the BatchNorm0-dependent rest: eqs (3b)-(8) ====================================================
    {
      const float* Sg = (CELL == PTK_CELL_INT) ? d.I_at(t) + base : d.A_at(t) + (size_t)b * d.a_bstride + pix;
      const float* It = d.I_at(t) + base;
      const float* Y0 = d.Y0_at(t) + base;
      float* Gi = d.Gi_at(t) + base;
      float* Iout = d.I_at(t + 1) + base;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float y0v[8], sv[8], iv[8], out[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const size_t o = (size_t)(q * 8 + e) * P;
          y0v[e] = __ldcg(Y0 + o);
          sv[e] = __ldcg(Sg + o);
          if (CELL != PTK_CELL_INT) iv[e] = __ldcg(It + o);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = q * 8 + e;
          const float X = actf<ACT>(fmaf(sWp[c * 3 + 2], xin[2], fmaf(sWp[c * 3 + 1], xin[1], fmaf(sWp[c * 3], xin[0], sbp[c]))));
          const float g = sigm<ACT>(gi[c] + sbi[c]);
          const float U = fmaf(y0v[e], sScale[c], sShift[c]);
          const float Pp = U * fmaf(sAl[c], sv[e], sMu[c]);
          const float Q = actf<ACT>(Pp);
          const float Ihat = actf<ACT>(X - Q);
          const float I = (CELL == PTK_CELL_INT) ? sv[e] : iv[e];
          out[e] = (1.0f - g) * I + g * Ihat;
          Gi[(size_t)c * P] = g;
          Iout[(size_t)c * P] = out[e];
        }
        stage_px_conv(sA, out, q, pc);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    LP_STAMP(5);
    // the other tiles of this sample need the halo rows of I_{t+1}
    cta_arrive(sctr);
    starget += tps;
    cta_wait(sctr, starget);
    LP_STAMP(6);
    // ================= conv1: I_{t+1} -> Y1_t ==============================================================================
    loop_conv<SINGLE>(d.I_at(t + 1), reinterpret_cast<const uint8_t*>(d.w_exc_f), d.Y1_at(t), la.part[1], cta, H, W, x0, y0, b,
                      smem_raw, sMerge, tmem_base, 1u, true);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    LP_STAMP(7);
    cta_arrive(la.gctr);
    LP_STAMP(8);
    gtarget += (unsigned)nCta;
    // ================= post ===================================================================================================
    const int has_next = (t + 1 < d.T);
    if (has_next) {   // statistics-independent: stem of frame t+1 and its share Wa_w X of the attention gate (eq. 1)
      if (tid == 0) {
        mbar_expect_tx(wbar, GW_POST_BYTES);
        bulk_g2s(smem_raw + L2_GW, gateB + GW_POST, GW_POST_BYTES, wbar);
      }
      float v[32];
#pragma unroll
      for (int i = 0; i < 3; ++i) xin[i] = d.fv.at(b, i, t + 1, gy, gx);
#pragma unroll
      for (int c = 0; c < 32; ++c)
        v[c] = actf<ACT>(fmaf(sWp[c * 3 + 2], xin[2], fmaf(sWp[c * 3 + 1], xin[1], fmaf(sWp[c * 3], xin[0], sbp[c]))));
      stage_px(smem_raw + L2_IMG0, v, p);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (warp == 1) {
        mbar_wait(wbar, 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gate_mma(simg + L2_IMG0, sgw + GW_POST_X, 32, tmem_base, 32, true);
        if (elect_one()) umma_commit(gbar);
        __syncwarp();
      }
    }
    // the convolution is through with its operand image: zero it for the Ge_{t+1} rows written below
    for (int i = tid; i < T2_A_BYTES / 16; i += LP_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    if (has_next) {   // image 0 is read by the MMAs above; it is reused for the fold scratch and E_{t+1}
      mbar_wait(gbar, gph & 1u);
      ++gph;
    }
    LP_STAMP(9);
    // ---- BatchNorm1 ----
    cta_wait(la.gctr, gtarget);
    LP_STAMP(10);
    loop_fold(la.part[1], nCta, T2_TH * T2_TW, sFin, sScale, sShift, d.bn_w[1], d.bn_b[1], cta == 0 ? d.stat(t, 1) : nullptr);
    LP_STAMP(11);
    float e2[32];
    {
      const float* Y1 = d.Y1_at(t) + base;
      const float* I2 = d.I_at(t + 1) + base;
      const float* Et = d.E_at(t) + base;
      const float* Gx = d.Gex_at(t) + base;
      float* Eout = d.E_at(t + 1) + base;
      float sread = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float y1[8], i2[8], ev[8], gxv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const size_t o = (size_t)(q * 8 + e) * P;
          y1[e] = __ldcg(Y1 + o); i2[e] = __ldcg(I2 + o); ev[e] = __ldcg(Et + o); gxv[e] = __ldcg(Gx + o);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = q * 8 + e;
          const float V = fmaf(y1[e], sScale[c], sShift[c]);
          const float R = V * fmaf(sKa[c], i2[e], sGa[c]);
          const float Ehat = actf<ACT>(R);
          e2[c] = (1.0f - gxv[e]) * ev[e] + gxv[e] * Ehat;
          sread = fmaf(sWr[c], e2[c], sread);
          Eout[(size_t)c * P] = e2[c];
        }
      }
      if (d.states != nullptr) d.states[((size_t)b * d.T + t) * P + pix] = sread + br;   // testmode: readout_conv(excitation), InT.py:233
    }
    if (has_next) {   // attention gate of frame t+1: A = sigmoid(Wa_w X + Wa_u E + b), Ge = A * E  (eqs 1-2)
      stage_px(smem_raw + L2_IMG0, e2, p);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gate_mma(simg + L2_IMG0, sgw + GW_POST_E, 32, tmem_base, 32, false);
        if (elect_one()) umma_commit(gbar);
        __syncwarp();
      }
      mbar_wait(gbar, gph & 1u);
      ++gph;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float a[32];
      tmem_ld32(tq + (uint32_t)(mt * 32), a);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      float* Aout = d.A_at(t + 1) + (size_t)b * d.a_bstride + pix;
      float* Geout = d.Ge + base;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float ge[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = q * 8 + e;
          const float av = sigm<ACT>(a[c] + sba[c]);
          ge[e] = av * e2[c];
          Aout[(size_t)c * P] = av;
          Geout[(size_t)c * P] = ge[e];
        }
        stage_px_conv(sA, ge, q, pc);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      LP_STAMP(12);
      cta_arrive(sctr);
      starget += tps;
      cta_wait(sctr, starget);
      LP_STAMP(13);
    }
  }
  LP_STAMP_STORE(cta);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

#if PTK_DEBUG_STAMPS
int read_loop_stamps(long long* out) {
  PTK_CUDA_OK(cudaMemcpyFromSymbol(out, g_loop_acc, sizeof(long long) * 148 * 16));
  return PTK_OK;
}
#endif

// ---- host ------------------------------------------------------------------------------------------------------------------------
size_t int_loop2_gate_bytes() { return GW_TOTAL; }

template <int CELL, int ACT, bool SINGLE>
static int launch_loop2(const IntDev& d, const LoopArgs& la, const uint8_t* gateB, cudaStream_t st) {
  auto kern = int_loop2_kernel<CELL, ACT, SINGLE>;
  PTK_CUDA_OK(ptk_func_smem(kern, LP_SMEM));
  dim3 grid(d.W / T2_TW, d.H / T2_TH, d.B);
  IntDev dd = d;
  LoopArgs ll = la;
  const uint8_t* gb = gateB;
  void* args[] = {&dd, &ll, &gb};
  ProfileScope ps(PTK_PROF_CONV_FWD, st);
  PTK_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(LP_THREADS), args, (size_t)LP_SMEM, st));
  return PTK_OK;
}

int run_int_forward_loop2(const IntDev& d, int cell, int act, float* part1, unsigned* sync_words, uint8_t* gateB, cudaStream_t st) {
  LoopArgs la;
  la.part[0] = d.bn_part; la.part[1] = part1;
  la.gctr = sync_words; la.sctr = sync_words + 64;
  la.tiles_per_sample = (d.H / T2_TH) * (d.W / T2_TW);
  PTK_CUDA_OK(cudaMemsetAsync(sync_words, 0, int_loop_sync_words(d.B) * sizeof(unsigned), st));
  IntGateSrc g;
  for (int i = 0; i < 6; ++i) g.w[i] = d.gate[i];
  prep_gates_tc_kernel<<<(6 * C * C + 255) / 256, 256, 0, st>>>(g, gateB);
  PTK_LAUNCH_OK("prep_gates_tc_kernel");
  const bool single = d.conv_mode == PTK_CONV_BF16;
  const bool tanh_ = act == PTK_ACT_TANH;
#define PTK_LOOP_CASE(CELL, ACTV)                                                                      \
  return single ? launch_loop2<CELL, ACTV | ACT_FAST, true>(d, la, gateB, st) : launch_loop2<CELL, ACTV | ACT_FAST, false>(d, la, gateB, st)
  if (cell == PTK_CELL_INT) {
    if (tanh_) { PTK_LOOP_CASE(PTK_CELL_INT, PTK_ACT_TANH); }
    PTK_LOOP_CASE(PTK_CELL_INT, PTK_ACT_SOFTPLUS);
  }
  if (tanh_) { PTK_LOOP_CASE(PTK_CELL_HGRU, PTK_ACT_TANH); }
  PTK_LOOP_CASE(PTK_CELL_HGRU, PTK_ACT_SOFTPLUS);
#undef PTK_LOOP_CASE
}

}  // namespace ptk
