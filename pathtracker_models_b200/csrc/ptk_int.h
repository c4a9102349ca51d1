// Device-side view of one InT / hGRU problem instance (parameters, checkpoints, scratch).
#pragma once
#include "ptk_common.cuh"
#include "ptk_kernels.h"

namespace ptk {

struct IntGateSrc {
  const float* w[6];
};

// per-tile accumulator layout of the backward point-wise kernels (floats)
enum {
  ACC_KAPPA = 0, ACC_GAMMA = 32, ACC_ALPHA = 64, ACC_MU = 96, ACC_BE = 128, ACC_BI = 160, ACC_BA = 192,
  ACC_BP = 224, ACC_WP = 256, ACC_WEW = 352, ACC_WEU = ACC_WEW + 1024, ACC_WIW = ACC_WEU + 1024,
  ACC_WIU = ACC_WIW + 1024, ACC_WAW = ACC_WIU + 1024, ACC_WAU = ACC_WAW + 1024, ACC_N = ACC_WAU + 1024
};
constexpr int HEAD_N = 96;   // per-sample readout-head partials (86 used), see int_readout_bwd_kernel
constexpr int WGRAD_GROUPS = 37;  // 4 ic-chunks x 37 groups = 148 CTAs (one per SM)

struct IntDev {
  int B, T, H, W;
  FrameView fv;
  // parameters (views into the flat vector) and re-laid-out copies (workspace)
  const float *Wp, *bp, *alpha, *mu, *gamma, *kappa;
  const float* gate[6];    // [out][in], a_w a_u i_w i_u e_w e_u
  const float* gateT[6];   // [in][out]
  const float* gate_b[6];
  const float *bn_w[2], *bn_b[2];
  const float *Wr, *br, *Wt, *bt, *Wd, *bd;
  const float *w_inh_f, *w_exc_f, *w_inh_d, *w_exc_d;
  // checkpoints (saved for backward) -- or a 2-slot ring when inference-only
  float *ckI, *ckE;
  size_t ck_stride;
  int ck_ring;
  float *Y0, *Y1;
  size_t y_stride;
  float* stats;            // [T or 1][2 bn][2 (mean,rstd)][C]
  // forward scratch / outputs
  float* A;                // attention gate: scratch [B,C,P] or gates output [B,T,C,P]
  long long a_tstride, a_bstride;
  float *Ge, *Gex;
  float *states, *logits;
  float* bn_part;
  unsigned* counter;
  // backward scratch
  float *gE, *gI, *geb, *dAct, *dGe, *dX, *dS;
  float* acc;              // [tiles][ACC_N]
  float* bn_sums;          // [tiles][2][C] step-local partials
  float* bn_tot;           // [2][C]  (sum dV)/N, (sum dV*xhat)/N of the current BatchNorm
  double* bn_acc;          // [2 bn][2][C] accumulated over steps: d(bn weight), d(bn bias)
  float* head;             // [B][HEAD_N]
  float* dgoal;            // [B][P] gradient of the goal map x[:,2,0] through the readout
  float *wacc_exc, *wacc_inh;
  float* dx;               // optional [B,3,T,H,W]
  float* dparams;
  const float* dlogits;

  __host__ __device__ int slot(int t) const { return ck_ring ? (t & 1) : t; }
  __host__ __device__ int yslot(int t) const { return ck_ring ? 0 : t; }
  __host__ __device__ const float* stat(int t, int bn) const { return stats + ((size_t)yslot(t) * 2 + bn) * 2 * C; }
  __host__ __device__ float* stat_w(int t, int bn) const { return stats + ((size_t)yslot(t) * 2 + bn) * 2 * C; }
};

int run_int_forward(const IntDev& d, int cell, int act, cudaStream_t st);
int run_int_backward(const IntDev& d, int cell, int act, cudaStream_t st);
int run_prep_weights(const float* w_inh, const float* w_exc, float* inh_f, float* inh_d, float* exc_f,
                     float* exc_d, const IntGateSrc& g, float* gateT, cudaStream_t st);

}  // namespace ptk
