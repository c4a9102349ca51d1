// Forward time loop of the InT / hGRU circuit: the point-wise (gate / blend / BatchNorm-apply)
// stages between the two 7x7 convolutions of each step, the stem and the readout.
//
// Reference being replaced (per step): models/InT.py:145-179 (rCell.forward) and
// models/ffhgru_hierarchy.py:135-173 (hConvGRUCell.forward); stem models/InT.py:212-213;
// readout models/InT.py:236-241.  Equation numbers (1)-(13) are those of SURVEY.md section 8.
//
// Warp-tile kernels (ptk_pw.cuh): a warp owns 32 pixels x 32 channels, the six 1x1 gate
// convolutions are register-tiled 32x32x32 products against shared-memory tiles; CTAs are
// persistent (weights staged once) and warps stride over the tiles.
#include "ptk_conv_stats.cuh"
#include "ptk_int.h"
#include "ptk_int_fwd_tiles.cuh"
#include "ptk_profile.h"
#include "ptk_pw.cuh"

namespace ptk {


template <int ACT, bool NOINH>
__global__ void __launch_bounds__(128, 4) int_pre_kernel(IntDev d, int t) {
  extern __shared__ __align__(16) float dsm[];
  float* sWaw = dsm;
  float* sWau = sWaw + C * C;
  float* sWew = sWau + C * C;                       // only staged when NOINH
  float* sWeu = sWew + (NOINH ? C * C : 0);
  float* tiles = sWeu + (NOINH ? C * C : 0);
  __shared__ float sWp[C * 3], sbp[C], sba[C], sbe[C];
  const float* const* gsrc = (ACT & ACT_FAST) ? d.gateTF : d.gateT;   // B-fragment order for the mma products
  cta_fill(sWaw, gsrc[0], C * C);
  cta_fill(sWau, gsrc[1], C * C);
  if (NOINH) { cta_fill(sWew, gsrc[4], C * C); cta_fill(sWeu, gsrc[5], C * C); }
  cta_fill(sWp, d.Wp, C * 3);
  cta_fill(sbp, d.bp, C);
  if (threadIdx.x < C) {
    sba[threadIdx.x] = __ldg(d.gate_b[0] + threadIdx.x) + __ldg(d.gate_b[1] + threadIdx.x);
    sbe[threadIdx.x] = __ldg(d.gate_b[4] + threadIdx.x) + __ldg(d.gate_b[5] + threadIdx.x);
  }
  __syncthreads();
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  float* sX = tiles + warp * 2 * TILE_F;
  float* sE = sX + TILE_F;
  for (int wt = blockIdx.x * 4 + warp; wt < nWT; wt += gridDim.x * 4) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    warp_tile_load_async(sE, d.E_at(t), (size_t)C * P, b, pix_base, P);
    cp_async_commit();
    float xin[3][4];
    frame_px(d.fv, b, t, pix_base + L.pxo, d.W, xin);
    cp_async_wait<0>();
    __syncwarp();
    attention_tile<ACT, NOINH>(d, t, b, pix_base, sX, sE, sWaw, sWau, sba, sWp, sbp, sWew, sWeu, sbe, L, xin);
    __syncwarp();
  }
}

// eqs (3b)-(9): BatchNorm0 apply, inhibition drive, inhibition gate + blend, excitation gate.
template <int CELL, int ACT>
__global__ void __launch_bounds__(128, 4) int_mid_kernel(IntDev d, int t) {
  extern __shared__ __align__(16) float dsm[];
  float* sWiw = dsm;
  float* sWiu = sWiw + C * C;
  float* sWew = sWiu + C * C;
  float* sWeu = sWew + C * C;
  float* tiles = sWeu + C * C;
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sbi[C], sbe[C], sAl[C], sMu[C], sScale[C], sShift[C];
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  float* sS = tiles + warp * 2 * TILE_F;
  float* sGX = sS + TILE_F;            // holds Ge for the excitation-gate product, then X
  const float* It = d.I_at(t);
  float* Iout = d.I_at(t + 1);
  const float* Y0 = d.Y0_at(t);
  int wt = blockIdx.x * 4 + warp;
  float xin[3][4];
  // prologue on parameters and the (read-only) input clip only: overlaps the previous kernel's tail (PDL)
  const float* const* gsrc = (ACT & ACT_FAST) ? d.gateTF : d.gateT;   // B-fragment order for the mma products
  cta_fill_async(sWiw, gsrc[2], C * C);
  cta_fill_async(sWiu, gsrc[3], C * C);
  cta_fill_async(sWew, gsrc[4], C * C);
  cta_fill_async(sWeu, gsrc[5], C * C);
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  cta_fill_async(sAl, d.alpha, C);
  cta_fill_async(sMu, d.mu, C);
  cp_async_commit();
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    sbi[c] = __ldg(d.gate_b[2] + c) + __ldg(d.gate_b[3] + c);
    sbe[c] = __ldg(d.gate_b[4] + c) + __ldg(d.gate_b[5] + c);
  }
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    frame_px(d.fv, b, t, pix_base + L.pxo, d.W, xin);
  }
  pdl_wait();
  pdl_launch_dependents();
  if (wt < nWT) {     // the first tile's operands
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    if (CELL == PTK_CELL_INT) warp_tile_load_async(sS, It, (size_t)C * P, b, pix_base, P);
    else warp_tile_load_async(sS, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P);
    warp_tile_load_async(sGX, d.Ge, (size_t)C * P, b, pix_base, P);
  }
  cp_async_commit();
  if (d.defer_stats) {                // the convolution left per-CTA partials: fold them here (every CTA, same order)
    consumer_stats_fold(d.bn_part, d.B * (d.H / 16) * (d.W / 32), 16 * 32, d.bn_w[0], d.bn_b[0], sScale, sShift,
                        blockIdx.x == 0 ? d.stat(t, 0) : nullptr);
  } else if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const float* st = d.stat(t, 0);   // BN0 of step t: mean, rstd (written by the convolution just before)
    const float sc = __ldg(d.bn_w[0] + c) * st[C + c];
    sScale[c] = sc;                   // U = Y0*scale + shift
    sShift[c] = __ldg(d.bn_b[0] + c) - st[c] * sc;
  }
  cp_async_wait<0>();
  __syncthreads();
  const MidW mw{sWiw, sWiu, sWew, sWeu, sWp, sbp, sbi, sbe, sAl, sMu, sScale, sShift};
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * 4, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    const float* Sg = (CELL == PTK_CELL_INT) ? It + base : d.A_at(t) + (size_t)b * d.a_bstride + pix;
    if (!firstTile) {
      if (CELL == PTK_CELL_INT) warp_tile_load_async(sS, It, (size_t)C * P, b, pix_base, P);
      else warp_tile_load_async(sS, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P);
      warp_tile_load_async(sGX, d.Ge, (size_t)C * P, b, pix_base, P);
      cp_async_commit();
      frame_px(d.fv, b, t, pix, d.W, xin);
      cp_async_wait<0>();
      __syncwarp();
    }
    MidIn cur;
    mid_tile_head<CELL, ACT>(d, t, base, Sg, sS, sGX, mw, L, xin, cur, It, Y0);
    mid_tile_tail<CELL, ACT>(d, t, base, Sg, sS, sGX, mw, L, cur, It, Y0, Iout);
    __syncwarp();
  }
}

// eqs (10b)-(13) for step t, then (1)-(2) for step t+1 (if any).
// Latency structure (profiles/r1f_pointwise_full.txt: global-load scoreboard stalls dominated): the first
// tile's operands are requested before the weight fill, the fill itself is cp.async, and the per-channel
// loop prefetches the next channel's operands while the current one is computed.
#if PTK_DEBUG_STAMPS   // tuning instrumentation, not in the shipped library (see ptk_conv_tc.cu)
__device__ unsigned long long g_pw_stamp[4];   // globaltimer at entry / after the PDL wait of int_post_kernel's CTA 0
#define PW_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_pw_stamp[i])); } while (0)
#else
#define PW_STAMP(i) do { } while (0)
#endif
template <int ACT, bool NOINH>
__global__ void __launch_bounds__(128, 4) int_post_kernel(IntDev d, int t, int has_next) {
  PW_STAMP(0);
  extern __shared__ __align__(16) float dsm[];
  float* sWaw = dsm;
  float* sWau = sWaw + C * C;
  float* sWew = sWau + C * C;                       // only staged when NOINH
  float* sWeu = sWew + (NOINH ? C * C : 0);
  float* tiles = sWeu + (NOINH ? C * C : 0);
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sba[C], sbe[C], sKa[C], sGa[C], sScale[C], sShift[C], sWr[C];
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  const float* Et = d.E_at(t);
  const float* I2 = d.I_at(t + 1);
  const float* Y1 = d.Y1_at(t);
  const float* Gx = d.Gex_at(t);
  int wt = blockIdx.x * 4 + warp;
  PostIn cur;
  float xin[3][4];
  // prologue on parameters and the (read-only) input clip only: overlaps the previous kernel's tail (PDL)
  const float* const* gsrc = (ACT & ACT_FAST) ? d.gateTF : d.gateT;   // B-fragment order for the mma products
  cta_fill_async(sWaw, gsrc[0], C * C);
  cta_fill_async(sWau, gsrc[1], C * C);
  if (NOINH) { cta_fill_async(sWew, gsrc[4], C * C); cta_fill_async(sWeu, gsrc[5], C * C); }
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  cta_fill_async(sKa, d.kappa, C);
  cta_fill_async(sGa, d.gamma, C);
  cta_fill_async(sWr, d.Wr, C);
  cp_async_commit();
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    sba[c] = __ldg(d.gate_b[0] + c) + __ldg(d.gate_b[1] + c);
    sbe[c] = __ldg(d.gate_b[4] + c) + __ldg(d.gate_b[5] + c);
  }
  if (wt < nWT && has_next) {
    const int b = wt / tpi, pix = (wt - b * tpi) * WPX + L.pxo;
    frame_px(d.fv, b, t + 1, pix, d.W, xin);
  }
  pdl_wait();
  PW_STAMP(1);
  pdl_launch_dependents();
  if (wt < nWT) {    // first tile: operands of channel 0
    const int b = wt / tpi, pix = (wt - b * tpi) * WPX + L.pxo;
    post_load(Y1, I2, Et, Gx, (size_t)b * C * P + pix + (size_t)L.oc0 * P, cur);
  }
  if (d.defer_stats) {
    consumer_stats_fold(d.bn_part, d.B * (d.H / 16) * (d.W / 32), 16 * 32, d.bn_w[1], d.bn_b[1], sScale, sShift,
                        blockIdx.x == 0 ? d.stat(t, 1) : nullptr);
  } else if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const float* st = d.stat(t, 1);   // BN1 of step t (written by the convolution just before)
    const float sc = __ldg(d.bn_w[1] + c) * st[C + c];
    sScale[c] = sc;
    sShift[c] = __ldg(d.bn_b[1] + c) - st[c] * sc;
  }
  cp_async_wait<0>();
  __syncthreads();
  float* sX = tiles + warp * 2 * TILE_F;
  float* sE = sX + TILE_F;
  const float br = __ldg(d.br);
  const PostW pw{sWaw, sWau, sWew, sWeu, sWp, sbp, sba, sbe, sKa, sGa, sScale, sShift, sWr};
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * 4, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    if (!firstTile) {
      post_load(Y1, I2, Et, Gx, base + (size_t)L.oc0 * P, cur);
      if (has_next) frame_px(d.fv, b, t + 1, pix, d.W, xin);
    }
    post_tile<ACT, NOINH>(d, t, has_next, b, pix_base, sX, sE, pw, br, L, xin, cur);
    __syncwarp();
  }
}

// Readout (models/InT.py:236-241): 1x1 conv C->1, concat goal channel x[:,2,0], 5x5 conv 2->1
// (pad 2), global mean, Linear(1,1).  One CTA per sample.
__global__ void __launch_bounds__(256) int_readout_kernel(IntDev d, const float* __restrict__ E, float* logits) {
  extern __shared__ float sm[];
  const int H = d.H, W = d.W, P = H * W, Wp2 = W + 4, Hp2 = H + 4;
  float* sR = sm;                 // [H+4][W+4] zero-padded readout map
  float* sG = sm + Hp2 * Wp2;     // goal map
  __shared__ float sWt[50], sWr[C], sRed[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 50) sWt[tid] = __ldg(d.Wt + tid);
  if (tid < C) sWr[tid] = __ldg(d.Wr + tid);
  for (int i = tid; i < 2 * Hp2 * Wp2; i += 256) sm[i] = 0.f;
  __syncthreads();
  const float br = __ldg(d.br);
  for (int pix = tid; pix < P; pix += 256) {
    float r = br;
#pragma unroll 8
    for (int c = 0; c < C; ++c) r = fmaf(sWr[c], E[((size_t)b * C + c) * P + pix], r);
    const int y = pix / W, x = pix % W;
    sR[(y + 2) * Wp2 + x + 2] = r;
    sG[(y + 2) * Wp2 + x + 2] = d.fv.at(b, 2, 0, y, x);
  }
  __syncthreads();
  float s = 0.f;
  for (int pix = tid; pix < P; pix += 256) {
    const int y = pix / W, x = pix % W;
    float o = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        o = fmaf(sWt[ky * 5 + kx], sR[(y + ky) * Wp2 + x + kx], o);
        o = fmaf(sWt[25 + ky * 5 + kx], sG[(y + ky) * Wp2 + x + kx], o);
      }
    s += o;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((tid & 31) == 0) sRed[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += sRed[i];
    const float pooled = tot / (float)P + __ldg(d.bt);
    logits[b] = fmaf(__ldg(d.Wd), pooled, __ldg(d.bd));
  }
}

// weight re-layouts done once per call: 7x7 banks -> [ic][tap][oc] (forward) and the flipped,
// in/out-swapped bank for conv_transpose (dgrad); 1x1 gates -> [in][out].
__global__ void prep_weights_kernel(const float* __restrict__ w /*[oc][ic][49]*/, float* wf, float* wd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * C * TAPS) return;
  const int oc = i / (C * TAPS), ic = (i / TAPS) % C, tap = i % TAPS;
  const float v = w[i];
  if (wf) wf[(ic * TAPS + tap) * C + oc] = v;
  if (wd) wd[(oc * TAPS + (TAPS - 1 - tap)) * C + ic] = v;   // conv_transpose2d(dY, w, padding=3)
}
// gateT = [in][out]; gateT + 6*1024: the same matrix in mma B-fragment order (ptk_pw.cuh); gateT + 12*1024: the
// [out][in] matrix (k = out, n = in: the transposed products of the backward kernels) in B-fragment order.
__global__ void prep_gates_kernel(IntGateSrc g, float* gateT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 6 * C * C) return;
  const int gi = i / (C * C), o = (i / C) % C, in = i % C;
  gateT[gi * C * C + in * C + o] = g.w[gi][o * C + in];
  // fragment-order element r of matrix gi: (kt, nt, lane, j) -> k = kt*8 + (lane&3) + 4j, n = nt*8 + (lane>>2)
  const int r = i % (C * C), j = r & 1, lane = (r >> 1) & 31, nt = (r >> 6) & 3, kt = r >> 8;
  const int k = kt * 8 + (lane & 3) + 4 * j, n = nt * 8 + (lane >> 2);
  gateT[(6 + gi) * C * C + r] = g.w[gi][n * C + k];    // gateT[k][n] = w[out = n][in = k]
  gateT[(12 + gi) * C * C + r] = g.w[gi][k * C + n];   // gate[k = out][n = in]
}

// ---- host ----------------------------------------------------------------------------------
int pw_grid(int B, int H, int W) {
  const int nWT = B * H * W / WPX;
  const int g = (nWT + 3) / 4;
  return g < PW_MAX_CTAS ? g : PW_MAX_CTAS;
}

static constexpr size_t SM_PRE = (size_t)(2 * C * C + 4 * 2 * TILE_F) * sizeof(float);
static constexpr size_t SM_MID = (size_t)(4 * C * C + 4 * 2 * TILE_F) * sizeof(float);

constexpr int CELL_NOINH = 2;   // internal: InT(no_inh=True)
static constexpr size_t SM_PRE_NI = SM_PRE + 2 * C * C * sizeof(float);

template <int ACT, bool NOINH>
static int run_forward_act(const IntDev& d_in, int cell, cudaStream_t st) {
  IntDev d = d_in;
  // second-generation tensor-core convolutions leave the BatchNorm partials to their consumers (int_mid / int_post)
  d.defer_stats = (d.conv_mode != PTK_CONV_FP32) && conv_v2_enabled() && stats_deferred();
  const int grid = pw_grid(d.B, d.H, d.W);
  const size_t sm_pre = NOINH ? SM_PRE_NI : SM_PRE;
  PTK_CUDA_OK(ptk_func_smem(int_pre_kernel<ACT, NOINH>, (int)sm_pre));
  PTK_CUDA_OK(ptk_func_smem(int_post_kernel<ACT, NOINH>, (int)sm_pre));
  PTK_CUDA_OK(ptk_func_smem(int_mid_kernel<PTK_CELL_INT, ACT>, (int)SM_MID));
  PTK_CUDA_OK(ptk_func_smem(int_mid_kernel<PTK_CELL_HGRU, ACT>, (int)SM_MID));
  {
    ProfileScope ps(PTK_PROF_POINT_FWD, st);
    int_pre_kernel<ACT, NOINH><<<grid, 128, sm_pre, st>>>(d, 0);
  }
  PTK_LAUNCH_OK("int_pre_kernel");
  // tensor-core modes, whole grid co-resident: ONE persistent cooperative kernel walks all frames (ptk_int_loop.cu)
  const bool loop = !NOINH && (ACT & ACT_FAST) && d.conv_mode != PTK_CONV_FP32 && conv_v2_enabled() && int_loop_mode() > 0 &&
                    int_loop_fits(d.B, d.H, d.W);
  if (loop) {
    const int rc = int_loop_mode() == 1 ? run_int_forward_loop(d, cell, ACT & 1, d.bn_part2, d.loop_sync, st)
                                        : run_int_forward_loop2(d, cell, ACT & 1, d.bn_part2, d.loop_sync, d.loop_gates, st);
    if (rc) return rc;
    PTK_LAUNCH_OK("int_loop_kernel");
  }
  for (int t = 0; t < d.T && !loop; ++t) {
    int rc;
    if (!NOINH) {
      ConvArgs c0{d.Ge, d.w_inh_f, d.Y0_at(t), d.B, d.H, d.W, 0, d.bn_part, d.counter, d.stat(t, 0), d.conv_mode, 1};
    c0.v2 = (d.conv_mode != PTK_CONV_FP32) && conv_v2_enabled();
      c0.defer_stats = d.defer_stats;
      { ProfileScope ps(PTK_PROF_CONV_FWD, st); rc = launch_conv7(c0, st); }
      if (rc) return rc;
      {
        ProfileScope ps(PTK_PROF_POINT_FWD, st);
        if (cell == PTK_CELL_INT) PTK_CUDA_OK(launch_pdl(true, int_mid_kernel<PTK_CELL_INT, ACT>, dim3(grid), dim3(128), SM_MID, st, d, t));
        else PTK_CUDA_OK(launch_pdl(true, int_mid_kernel<PTK_CELL_HGRU, ACT>, dim3(grid), dim3(128), SM_MID, st, d, t));
      }
      PTK_LAUNCH_OK("int_mid_kernel");
    }
    ConvArgs c1{d.I_at(t + 1), d.w_exc_f, d.Y1_at(t), d.B, d.H, d.W, 0, d.bn_part, d.counter, d.stat(t, 1), d.conv_mode, 1};
    c1.v2 = (d.conv_mode != PTK_CONV_FP32) && conv_v2_enabled();
    c1.defer_stats = d.defer_stats;
    { ProfileScope ps(PTK_PROF_CONV_FWD, st); rc = launch_conv7(c1, st); }
    if (rc) return rc;
    {
      ProfileScope ps(PTK_PROF_POINT_FWD, st);
      PTK_CUDA_OK(launch_pdl(true, int_post_kernel<ACT, NOINH>, dim3(grid), dim3(128), sm_pre, st, d, t, (int)(t + 1 < d.T)));
    }
    PTK_LAUNCH_OK("int_post_kernel");
  }
  const size_t sm = (size_t)2 * (d.H + 4) * (d.W + 4) * sizeof(float);
  PTK_CUDA_OK(ptk_func_smem(int_readout_kernel, (int)sm));   // above the 48 KB default from 64x96 frames on
  {
    ProfileScope ps(PTK_PROF_MISC, st);
    int_readout_kernel<<<d.B, 256, sm, st>>>(d, d.E_at(d.T), d.logits);
  }
  PTK_LAUNCH_OK("int_readout_kernel");
  return PTK_OK;
}

template <int ACT>
static int run_forward_ni(const IntDev& d, int cell, cudaStream_t st) {
  return cell == CELL_NOINH ? run_forward_act<ACT, true>(d, cell, st) : run_forward_act<ACT, false>(d, cell, st);
}

#if PTK_DEBUG_STAMPS
int read_pw_stamps(unsigned long long* out) {
  PTK_CUDA_OK(cudaMemcpyFromSymbol(out, g_pw_stamp, sizeof(unsigned long long) * 4));
  return PTK_OK;
}
#endif

int run_int_forward(const IntDev& d, int cell, int act, cudaStream_t st) {
  const bool fast = (d.conv_mode != PTK_CONV_FP32);
  if (act == PTK_ACT_SOFTPLUS) return fast ? run_forward_ni<PTK_ACT_SOFTPLUS | ACT_FAST>(d, cell, st)
                                           : run_forward_ni<PTK_ACT_SOFTPLUS>(d, cell, st);
  return fast ? run_forward_ni<PTK_ACT_TANH | ACT_FAST>(d, cell, st) : run_forward_ni<PTK_ACT_TANH>(d, cell, st);
}

int run_prep_weights(const float* w_inh, const float* w_exc, float* inh_f, float* inh_d, float* exc_f,
                     float* exc_d, const IntGateSrc& g, float* gateT, cudaStream_t st) {
  const int n = C * C * TAPS;
  ProfileScope ps(PTK_PROF_MISC, st, 3);
  if (w_inh) prep_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w_inh, inh_f, inh_d);
  if (w_exc) prep_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w_exc, exc_f, exc_d);
  if (gateT) prep_gates_kernel<<<(6 * C * C + 255) / 256, 256, 0, st>>>(g, gateT);
  PTK_LAUNCH_OK("prep_weights");
  return PTK_OK;
}

}  // namespace ptk
