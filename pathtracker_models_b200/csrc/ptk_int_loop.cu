// Persistent forward time loop of the InT / hGRU circuit: ONE cooperative kernel walks all T frames.
//
// Reference being replaced: the Python loop `for t in range(x.shape[2])` of models/InT.py:223-235 (and
// models/ffhgru_hierarchy.py:246-262) around rCell.forward / hConvGRUCell.forward.  The PDL chain of ptk_int_fwd.cu runs
// four launches per frame (conv, mid, conv, post); here one CTA per 16 x 32-pixel tile (the tile of conv7_tc2) stays
// resident for the whole clip and runs the same four phases itself:
//
//   conv0 (tcgen05, Ge_t -> Y0_t + BatchNorm partials)   -- grid barrier: ARRIVE
//   mid, the part that needs no batch statistics (excitation gate (9), stem, inhibition-gate pre-activation (7))
//                                                        -- grid barrier: WAIT, fold the partials (every CTA, same order)
//   mid, the BatchNorm0-dependent rest (eqs 3b-8) -> I_{t+1}        -- per-sample barrier (the 3-row halo of the next conv)
//   conv1 (I_{t+1} -> Y1_t + partials)                   -- grid barrier: ARRIVE
//   post operand loads                                   -- grid barrier: WAIT, fold
//   post (eqs 10b-13 -> E_{t+1}; attention (1)-(2) of frame t+1 -> A, Ge)        -- per-sample barrier
//
// What residency buys (there is no second copy of the arithmetic: the tile bodies are the ones of the chain, ptk_int_fwd_tiles.cuh,
// and the MMA / epilogue schedule is conv7_tc2's, so results are bit-identical to the chain):
//  * TMEM allocation, the per-channel parameter vectors and the kernel prologues happen once, not 4 T times;
//  * the two BatchNorm reductions are split arrive / wait barriers, so the statistics-independent half of `mid` and the
//    operand loads of `post` run while the slowest CTA is still convolving;
//  * all 16 warps stage the operand image (the stand-alone convolution has 8).
// Training and inference: the checkpoints the reverse loop reads are written exactly as by the chain.
//
// Shared memory is one 215 KB region used in turn as the convolution's [weight ring | operand image | quadrant tiles]
// and as the point-wise [gate matrices | 16 x 2 warp tiles | fold scratch]; the image is re-zeroed per convolution.
#include "ptk_int_loop.cuh"

namespace ptk {

template <int CELL, int ACT, bool SINGLE>
__global__ void __launch_bounds__(LP_THREADS, 1) int_loop_kernel(IntDev d, LoopArgs la) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  float* sGW = reinterpret_cast<float*>(smem_raw + LP_PW_W);
  float* tiles = reinterpret_cast<float*>(smem_raw + LP_PW_TILES);
  double (*sFin)[C][3] = reinterpret_cast<double (*)[C][3]>(smem_raw + LP_PW_FOLD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + LP_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T2_NST + 6);
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sbi[C], sbe[C], sba[C], sAl[C], sMu[C], sKa[C], sGa[C], sWr[C], sScale[C], sShift[C], sMerge[C * 4];

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int H = d.H, W = d.W, P = H * W;
  const int x0 = blockIdx.x * T2_TW, y0 = blockIdx.y * T2_TH, b = blockIdx.z;
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int nCta = gridDim.x * gridDim.y * gridDim.z;

  // ---- once per clip: barriers, TMEM, per-channel parameter vectors ------------------------------------------------------
  if (tid == 0) {
    uint64_t* full = bars;
    uint64_t* empty = bars + T2_NST;
    uint64_t* dFull = bars + 2 * T2_NST;
    uint64_t* dEmpty = dFull + 2;
    for (int i = 0; i < T2_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&dFull[0], 1); mbar_init(&dFull[1], 1);
    mbar_init(&dEmpty[0], T2_EPI_WARPS); mbar_init(&dEmpty[1], T2_EPI_WARPS);
    mbar_init(dEmpty + 2, LP_THREADS); mbar_init(dEmpty + 3, LP_THREADS - 64);
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

  const Lane L;
  // this warp's point-wise tile: image row y0 + warp of the CTA's tile = 32 consecutive pixels
  const int pix_base = (y0 + warp) * W + x0, pix = pix_base + L.pxo;
  const size_t base = (size_t)b * C * P + pix;
  float* sT0 = tiles + warp * 2 * TILE_F;
  float* sT1 = sT0 + TILE_F;
  const float br = __ldg(d.br);
  const MidW mw{sGW, sGW + C * C, sGW + 2 * C * C, sGW + 3 * C * C, sWp, sbp, sbi, sbe, sAl, sMu, sScale, sShift};
  const PostW pw{sGW, sGW + C * C, nullptr, nullptr, sWp, sbp, sba, sbe, sKa, sGa, sScale, sShift, sWr};
  unsigned gtarget = 0, starget = 0;
  unsigned* sctr = la.sctr + b;
  const unsigned tps = (unsigned)la.tiles_per_sample;
  LP_STAMP_DECL

#pragma unroll 1
  for (int t = 0; t < d.T; ++t) {
    // ================= conv0: Ge_t -> Y0_t ==================================================================================
    loop_conv<SINGLE>(d.Ge, reinterpret_cast<const uint8_t*>(d.w_inh_f), d.Y0_at(t), la.part[0], cta, H, W, x0, y0, b, smem_raw,
                      sMerge, tmem_base, 0u);
    LP_STAMP(0);                       // conv0: thread 0 (producer warp) through
    cta_arrive(la.gctr);
    LP_STAMP(1);                       // conv0: whole CTA through (epilogue done)
    gtarget += (unsigned)nCta;
    // ================= mid, statistics-independent half ===================================================================
    {
      cta_fill_async(sGW, d.gateTF[2], C * C);
      cta_fill_async(sGW + C * C, d.gateTF[3], C * C);
      cta_fill_async(sGW + 2 * C * C, d.gateTF[4], C * C);
      cta_fill_async(sGW + 3 * C * C, d.gateTF[5], C * C);
      const float* It = d.I_at(t);
      if (CELL == PTK_CELL_INT) warp_tile_load_async(sT0, It, (size_t)C * P, b, pix_base, P);
      else warp_tile_load_async(sT0, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P);
      warp_tile_load_async(sT1, d.Ge, (size_t)C * P, b, pix_base, P);
      cp_async_commit();
      float xin[3][4];
      frame_px(d.fv, b, t, pix, W, xin);
      cp_async_wait<0>();
      __syncthreads();
      const float* Sg = (CELL == PTK_CELL_INT) ? It + base : d.A_at(t) + (size_t)b * d.a_bstride + pix;
      MidIn cur;
      mid_tile_head<CELL, ACT>(d, t, base, Sg, sT0, sT1, mw, L, xin, cur, It, d.Y0_at(t));
      LP_STAMP(2);                     // mid head (warp 0)
      // ---- BatchNorm0: all convolutions of the batch are in ----
      cta_wait(la.gctr, gtarget);
      LP_STAMP(3);                     // grid barrier wait (incl. the other warps' heads)
      loop_fold(la.part[0], nCta, T2_TH * T2_TW, sFin, sScale, sShift, d.bn_w[0], d.bn_b[0], cta == 0 ? d.stat(t, 0) : nullptr);
      LP_STAMP(4);                     // fold
      mid_tile_tail<CELL, ACT>(d, t, base, Sg, sT0, sT1, mw, L, cur, It, d.Y0_at(t), d.I_at(t + 1));
      LP_STAMP(5);                     // mid tail (warp 0)
    }
    // the other tiles of this sample need the halo rows of I_{t+1}
    cta_arrive(sctr);
    starget += tps;
    cta_wait(sctr, starget);
    LP_STAMP(6);                       // sample barrier (incl. the other warps' tails)
    // ================= conv1: I_{t+1} -> Y1_t ==============================================================================
    loop_conv<SINGLE>(d.I_at(t + 1), reinterpret_cast<const uint8_t*>(d.w_exc_f), d.Y1_at(t), la.part[1], cta, H, W, x0, y0, b,
                      smem_raw, sMerge, tmem_base, 1u);
    LP_STAMP(7);
    cta_arrive(la.gctr);
    LP_STAMP(8);                       // conv1 complete
    gtarget += (unsigned)nCta;
    // ================= post ===================================================================================================
    {
      const int has_next = (t + 1 < d.T);
      cta_fill_async(sGW, d.gateTF[0], C * C);
      cta_fill_async(sGW + C * C, d.gateTF[1], C * C);
      cp_async_commit();
      float xin[3][4];
      if (has_next) frame_px(d.fv, b, t + 1, pix, W, xin);
      PostIn cur;
      post_load(d.Y1_at(t), d.I_at(t + 1), d.E_at(t), d.Gex_at(t), base + (size_t)L.oc0 * P, cur);
      cp_async_wait<0>();
      LP_STAMP(9);                     // post operand loads issued
      // ---- BatchNorm1 ----
      cta_wait(la.gctr, gtarget);
      LP_STAMP(10);
      loop_fold(la.part[1], nCta, T2_TH * T2_TW, sFin, sScale, sShift, d.bn_w[1], d.bn_b[1], cta == 0 ? d.stat(t, 1) : nullptr);
      LP_STAMP(11);
      post_tile<ACT, false>(d, t, has_next, b, pix_base, sT0, sT1, pw, br, L, xin, cur);
      LP_STAMP(12);                    // post tile (warp 0)
      if (has_next) {   // the other tiles of this sample need the halo rows of Ge_{t+1}
        cta_arrive(sctr);
        starget += tps;
        cta_wait(sctr, starget);
      }
      LP_STAMP(13);
    }
  }
  LP_STAMP_STORE(cta);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

// ---- host ------------------------------------------------------------------------------------------------------------------------
size_t int_loop_sync_words(int B) { return (size_t)64 + (size_t)B; }   // grid counter (own 256-byte line) + one counter per sample

template <int CELL, int ACT, bool SINGLE>
static int launch_loop(const IntDev& d, const LoopArgs& la, cudaStream_t st) {
  auto kern = int_loop_kernel<CELL, ACT, SINGLE>;
  PTK_CUDA_OK(ptk_func_smem(kern, LP_SMEM));
  dim3 grid(d.W / T2_TW, d.H / T2_TH, d.B);
  IntDev dd = d;
  LoopArgs ll = la;
  void* args[] = {&dd, &ll};
  ProfileScope ps(PTK_PROF_CONV_FWD, st);
  PTK_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(LP_THREADS), args, (size_t)LP_SMEM, st));
  return PTK_OK;
}

// can the whole grid be co-resident?  (one 215 KB CTA per SM)
bool int_loop_fits(int B, int H, int W) {
  static int sms = -1, coop = 0;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return coop && (long long)B * (H / T2_TH) * (W / T2_TW) <= (long long)sms;
}

int run_int_forward_loop(const IntDev& d, int cell, int act, float* part1, unsigned* sync_words, cudaStream_t st) {
  LoopArgs la;
  la.part[0] = d.bn_part; la.part[1] = part1;
  la.gctr = sync_words; la.sctr = sync_words + 64;
  la.tiles_per_sample = (d.H / T2_TH) * (d.W / T2_TW);
  PTK_CUDA_OK(cudaMemsetAsync(sync_words, 0, int_loop_sync_words(d.B) * sizeof(unsigned), st));
  const bool single = d.conv_mode == PTK_CONV_BF16;
  const bool tanh_ = act == PTK_ACT_TANH;
#define PTK_LOOP_CASE(CELL, ACTV)                                                                      \
  return single ? launch_loop<CELL, ACTV | ACT_FAST, true>(d, la, st) : launch_loop<CELL, ACTV | ACT_FAST, false>(d, la, st)
  if (cell == PTK_CELL_INT) {
    if (tanh_) { PTK_LOOP_CASE(PTK_CELL_INT, PTK_ACT_TANH); }
    PTK_LOOP_CASE(PTK_CELL_INT, PTK_ACT_SOFTPLUS);
  }
  if (tanh_) { PTK_LOOP_CASE(PTK_CELL_HGRU, PTK_ACT_TANH); }
  PTK_LOOP_CASE(PTK_CELL_HGRU, PTK_ACT_SOFTPLUS);
#undef PTK_LOOP_CASE
}

}  // namespace ptk
