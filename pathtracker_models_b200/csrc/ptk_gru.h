// Device-side view of one ConvGRU problem instance (models/kys.py, `--model gru`).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "ptk_common.cuh"

namespace ptk {

constexpr int GC = 64;   // hidden / stem channels (utils/engine.py:149: dimensions * 2)
constexpr int GRU_STEM_SPLIT = 4;   // CTAs per channel in gru_stem_grad_kernel (partials folded in fixed order at the end)
constexpr int GRU_HEAD_N = GC + 1 + 50 + 1 + 1 + 1;   // dWr, dbr, dWt, dbt, dWd, dbd (flat-vector order)

struct GruDev {
  int B, T, H, W, groups, conv_mode;
  FrameView fv;                                   // x [B,3,T,H,W], arbitrary strides
  // parameters (views into the flat vector): preproc.{w [64,3], b}, conv_reset / conv_update / conv_state_new {w [64,128,7,7], b}
  const float *Wp, *bp, *wr, *br, *wu, *bu, *wn, *bnw;
  long long off_wp, off_bp, off_wr, off_br, off_wu, off_bu, off_wn, off_bn;
  // readout head (models/kys.py:127-133): readout_conv [1,64,1,1], target_conv [1,2,5,5], readout_dense [1,1]
  const float *hWr, *hbr, *hWt, *hbt, *hWd, *hbd;
  long long off_head;      // the six head tensors are contiguous in the flat vector: 64, 1, 50, 1, 1, 1
  // re-laid-out weight banks and the gathered [br|bu] bias (workspace)
  float *wruT, *wnT, *wruD, *wnD, *bru;
  // checkpoints (saved): h [T+1][B,64,P]; reset / update gates and the candidate [T][B,64,P]
  float *hs, *rs, *us, *ns;
  // scratch
  float *Z, *Z2, *RU, *Nb, *dZ, *dh, *dhp, *dXb, *dNsum, *dRUsum;
  float *slots_ru, *slots_n, *stem_acc, *bias_grad;
  // tensor-core modes: 24 (out-block, in-block) pairs of 32x32 channels -- 16 for [reset|update] (pair = j*4 + i), 8 for
  // the candidate (16 + j*4 + i): bf16 hi/lo tap streams (forward / transposed) and per-CTA weight-gradient slots
  // bf16x3 mode: wide tcgen05 kernels shared with the ConvLSTM (ptk_convk_tc.cu / ptk_wgradk_tc.cu)
  uint8_t *wk_ruF, *wk_ruD, *wk_nF, *wk_nD;   // weight streams: [reset|update] 128 -> 128 and candidate 128 -> 64, forward / transposed
  void* dyc_k;                                // gate-gradient tiles pre-split for the weight-gradient CTAs (wgradk_prep_dy)
  float* slots_k;                             // 8 x wgradk_tc_slot_floats(7): input-channel blocks 0-3 of [r|u], then of the candidate
  uint8_t *wtcF, *wtcD;
  float* wacc_tc;
  size_t wacc_pair;      // floats per pair
  // I/O
  float* logits;        // [B]
  const float* dlogits; // [B]
  float *head_part, *dgoal;   // [B][GRU_HEAD_N] per-sample head gradients; [B,P] d x[:,2,0] through the goal channel
  float* dparams;
  float* dx;            // optional [B,3,T,H,W] contiguous

  __host__ __device__ size_t S() const { return (size_t)B * GC * H * W; }
  __host__ __device__ float* h_at(int t) const { return hs + (size_t)t * S(); }            // t = 0..T
  __host__ __device__ float* r_at(int t) const { return rs + (size_t)(t - 1) * S(); }      // t = 1..T
  __host__ __device__ float* u_at(int t) const { return us + (size_t)(t - 1) * S(); }
  __host__ __device__ float* n_at(int t) const { return ns + (size_t)(t - 1) * S(); }
};

int run_gru_forward(const GruDev& d, cudaStream_t st);
int run_gru_backward(const GruDev& d, cudaStream_t st);

}  // namespace ptk
