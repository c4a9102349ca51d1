// Device-side view of one InT / hGRU problem instance (parameters, checkpoints, scratch).
#pragma once
#include "ptk_common.cuh"
#include "ptk_kernels.h"

namespace ptk {

struct IntGateSrc {
  const float* w[6];
};

// per-CTA accumulator layout of the backward point-wise kernels (floats)
enum { ACC_KAPPA = 0, ACC_GAMMA = 32, ACC_ALPHA = 64, ACC_MU = 96, ACC_BP = 128, ACC_WP = 160, ACC_N = 256 };
// per-CTA accumulator layout of the gate weight-gradient kernel
enum {
  GACC_WEW = 0, GACC_WEU = 1024, GACC_WIW = 2048, GACC_WIU = 3072, GACC_WAW = 4096, GACC_WAU = 5120,
  GACC_BE = 6144, GACC_BI = 6176, GACC_BA = 6208, GACC_N = 6240
};
constexpr int HEAD_N = 96;        // per-sample readout-head partials (86 used), see int_readout_bwd_kernel
constexpr int WGRAD_GROUPS = 37;  // 4 ic-chunks x 37 groups = 148 CTAs (one per SM)
constexpr int GATE_CTAS = 148;    // gate weight-gradient kernel: one persistent CTA per SM
constexpr int PW_MAX_CTAS = 148 * 4;   // grid of the point-wise kernels: 4 CTAs/SM so that (cfg 2: 512 CTAs) every warp owns ONE tile
constexpr int PW_EW_CTAS = 148 * 4;    // grid of the purely element-wise bwd_a kernel; also the number of accumulator slots

struct IntDev {
  int B, T, H, W;
  int conv_mode;           // PTK_CONV_*
  FrameView fv;
  // parameters (views into the flat vector) and re-laid-out copies (workspace)
  const float *Wp, *bp, *alpha, *mu, *gamma, *kappa;
  const float* gate[6];    // [out][in], a_w a_u i_w i_u e_w e_u
  const float* gateT[6];   // [in][out]
  const float* gateTF[6];  // gateT in mma B-fragment order (ptk_pw.cuh), tensor-core modes
  const float* gateF[6];   // gate ([out][in], the k-by-n matrix of the transposed products) in B-fragment order
  const float* gate_b[6];
  const float *bn_w[2], *bn_b[2];
  const float *Wr, *br, *Wt, *bt, *Wd, *bd;
  const float *w_inh_f, *w_exc_f, *w_inh_d, *w_exc_d;
  // checkpoints (saved for backward) -- or a small ring when inference-only
  float *ckI, *ckE;        // [T+1 or 2][B,C,P]
  size_t ck_stride;
  int ck_ring;
  float *Y0, *Y1, *Gi, *Gex;   // [T or 1][B,C,P]: pre-BN conv outputs, inhibition / excitation gates
  size_t y_stride;
  float* A;                // attention gate per step: saved [T][B,C,P], ring scratch, or gates output [B,T,C,P]
  long long a_tstride, a_bstride;
  float* stats;            // [T or 1][2 bn][2 (mean,rstd)][C]
  // forward scratch / outputs
  float* Ge;
  float *states, *logits;
  float* bn_part;
  unsigned* counter;
  int defer_stats;         // forward chain: the convolutions leave per-CTA BatchNorm partials in bn_part, int_mid / int_post fold them
  float* bn_part2;         // persistent time-loop kernel (ptk_int_loop.cu): second partial-statistics buffer, barrier counters
  unsigned* loop_sync;
  uint8_t* loop_gates;     // gate matrices as bf16 hi/lo tcgen05 B operands (ptk_int_loop2.cu)
  // backward scratch
  float *gE, *gI, *geb, *dAct, *dGe, *dX, *dS, *giS, *aS, *Xs;
  float *GeB[2], *gebB[2]; // per-step-parity copies of Ge and geb (the fused bwd_d(t)+bwd_a(t-1) kernel writes step t-1's
                           // while gate_wgrad(t) still reads step t's)
  float *cgI, *cGe;        // data-gradient outputs of the two transposed convolutions (added by the consumers)
  // tensor-core modes: the 7x7 weight gradients are taken once, after the reverse loop, over ALL steps
  // (ptk_wgrad_ts.cu), so every step's dY0 / dY1 and Ge are kept: [T][B,C,P] each (dy_tstride = B*C*P, else 0)
  float *dYs[2], *GeT;
  float *gebT, *giT, *aT, *XT;   // likewise the operands of the six gate weight gradients (one launch over all steps at the end)
  size_t dy_tstride;
  int bn_fused;            // the BatchNorm backward (b3 / b9) runs inside the data-gradient convolution's operand staging
                           // (conv7_tc2): bwd_a / bwd_c write dV / dU to the scratch `dAct`, the convolution writes dY per step
  int pw_ctas;             // grid of the backward point-wise kernels (fixed => deterministic slots)
  float* acc;              // [pw_ctas][ACC_N]
  float* gacc;             // [GATE_CTAS][GACC_N]
  float* bn_sums;          // [pw_ctas][2][C] step-local partials
  float* bn_tot;           // [2][C]  (sum dV)/N, (sum dV*xhat)/N of the current BatchNorm
  double* bn_acc;          // [2 bn][2][C] accumulated over steps: d(bn bias), d(bn weight)
  float* head;             // [B][HEAD_N]
  float* dgoal;            // [B][P] gradient of the goal map x[:,2,0] through the readout
  float *wacc_exc, *wacc_inh;
  float* dx;               // optional [B,3,T,H,W]
  float* dparams;
  const float* dlogits;

  __host__ __device__ int slot(int t) const { return ck_ring ? (t & 1) : t; }
  __host__ __device__ int yslot(int t) const { return ck_ring ? 0 : t; }
  __host__ __device__ float* stat(int t, int bn) const { return stats + ((size_t)yslot(t) * 2 + bn) * 2 * C; }
  __host__ __device__ float* I_at(int t) const { return ckI + ck_stride * slot(t); }
  __host__ __device__ float* E_at(int t) const { return ckE + ck_stride * slot(t); }
  __host__ __device__ float* Y0_at(int t) const { return Y0 + y_stride * yslot(t); }
  __host__ __device__ float* Y1_at(int t) const { return Y1 + y_stride * yslot(t); }
  __host__ __device__ float* Gi_at(int t) const { return Gi + y_stride * yslot(t); }
  __host__ __device__ float* Gex_at(int t) const { return Gex + y_stride * yslot(t); }
  __host__ __device__ float* A_at(int t) const { return A + (size_t)t * a_tstride; }
  // gradient wrt the pre-BatchNorm conv output of BatchNorm `bn` at step t (dV / dU before bn_apply, dY after)
  __host__ __device__ float* dAct_at(int t, int bn) const { return dy_tstride ? dYs[bn] + dy_tstride * t : dAct; }
  __host__ __device__ float* dV_at(int t, int bn) const { return bn_fused ? dAct : dAct_at(t, bn); }
  __host__ __device__ float* GeB_at(int t) const { return dy_tstride ? GeT + dy_tstride * t : GeB[t & 1]; }
  __host__ __device__ float* geb_at(int t) const { return dy_tstride ? gebT + dy_tstride * t : gebB[t & 1]; }
  __host__ __device__ float* gi_at(int t) const { return dy_tstride ? giT + dy_tstride * t : giS; }
  __host__ __device__ float* a_at(int t) const { return dy_tstride ? aT + dy_tstride * t : aS; }
  __host__ __device__ float* X_at(int t) const { return dy_tstride ? XT + dy_tstride * t : Xs; }
};

int run_int_forward(const IntDev& d, int cell, int act, cudaStream_t st);
int run_int_backward(const IntDev& d, int cell, int act, cudaStream_t st);
int run_prep_weights(const float* w_inh, const float* w_exc, float* inh_f, float* inh_d, float* exc_f,
                     float* exc_d, const IntGateSrc& g, float* gateT, cudaStream_t st);
int pw_grid(int B, int H, int W);
bool stats_deferred();                 // PTK_STATS_DEFER=1: int_mid / int_post fold the convolutions' BatchNorm partials (default: the
                                       // convolution's last CTA does; measured equal)
// gate weight gradients of all steps on tcgen05 (ptk_gate_wgrad_tc.cu); returns the number of gacc slots written or < 0
int launch_gate_wgrad_tc(const IntDev& d, int cell, int nT, cudaStream_t st);
// persistent forward time loop (ptk_int_loop.cu): one cooperative kernel for all frames
int int_loop_mode();                   // PTK_FWD_LOOP: 0 = per-phase PDL chain, 1 = the chain's phases in one cooperative kernel
                                       // (bit-identical), 2 = second generation (tcgen05 gate products, thread = pixel)
bool int_loop_fits(int B, int H, int W);
size_t int_loop_sync_words(int B);
int run_int_forward_loop(const IntDev& d, int cell, int act, float* part1, unsigned* sync_words, cudaStream_t st);
size_t int_loop2_gate_bytes();
int run_int_forward_loop2(const IntDev& d, int cell, int act, float* part1, unsigned* sync_words, uint8_t* gateB, cudaStream_t st);

}  // namespace ptk
