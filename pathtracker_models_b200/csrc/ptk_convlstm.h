// Device-side view of one ConvLSTM problem instance (models/convlstm.py).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ptk {

struct LstmDev {
  int B, T, H, W, KS, groups, conv_mode;
  int ring;                             // checkpoints are two-slot rings (recurrent back-propagation: only steps T-1, T are read)
  void* dyc_tc;                         // delta of the step, pre-split per tile for the weight-gradient CTAs (wgradk_prep_dy)
  float *slots_h_tc, *slots_x_tc;       // tensor-core weight-gradient slots of Wh / Wx (ptk_wgradk_tc.cu)
  void *whF_tc, *whD_tc, *wxF_tc, *wxD_tc;   // [0] = Wh, [1] = Wx streams                // tensor-core weight streams of the recurrent convolution / its transpose
  const float* x;                       // [B,1,H,W]
  // parameters (views into the flat vector)
  const float *w0, *b0;                 // conv0 [25,1,7,7], [25]
  const float* wx[4];                   // Wxi, Wxf, Wxc, Wxo  [25,25,KS,KS]
  const float* wh[4];                   // Whi, Whf, Whc, Who
  const float* bx;                      // NOT contiguous in the flat vector: gathered copy [100] in workspace
  const float *bn_w, *bn_b, *w6, *b6;
  // re-laid-out weight banks (workspace)
  float *w0T, *wxT, *whT, *whD, *wxD;
  // checkpoints
  float *hs, *cs;                       // [T+1][B,25,P]   (ring: [2])
  float* gates;                         // [T][B,100,P] post-activation (i,f,g,o)   (ring: [2])
  float *u, *xt, *Xg;                   // conv0 output, its square, hoisted input terms
  float* head_stats;                    // [2][25]
  // scratch
  float *G, *delta, *delta_sum, *dh, *dc;
  float *slots_h, *slots_x, *slots_0, *bx_grad, *b0_grad, *head_grads;
  float *jp_rc, *jp_eo, *jp_dcx;        // Jacobian penalty in the graph: d/d jc, the step T-1 output-gate term, the extra d c_{T-1}
  const float* dpen;                    // d(total)/d(penalty) [B,25,H,W] (null: the penalty is a constant, create_graph=False)
  int detached;                         // grad_method='rbp': h_{T-1}, c_{T-1} are detached leaves (models/convlstm.py:130-131)
  float *rbp_g, *rbp_state;             // RBP: Neumann sum g [B,25,P]; state {|v_prev|, terminated, accepted terms}
  double* rbp_part;                     // [RBP_CTAS][2] partial sums of squares
  // I/O
  float* out;                           // [B,2,H,W]
  const float* dout;
  float* penalty;                       // [B,25,H,W]
  float* dparams;
  long long off_w0, off_b0, off_wx, off_bx, off_wh, off_head;

  __host__ __device__ float* h_at(int t) const { return hs + (size_t)(ring ? (t & 1) : t) * B * 25 * H * W; }
  __host__ __device__ float* c_at(int t) const { return cs + (size_t)(ring ? (t & 1) : t) * B * 25 * H * W; }
  __host__ __device__ float* gates_at(int t) const { return gates + (size_t)(ring ? (t & 1) : t - 1) * B * 100 * H * W; }   // t = 1..T
};

int run_convlstm_forward(const LstmDev& d, cudaStream_t st);
int run_convlstm_penalty(const LstmDev& d, cudaStream_t st);
int run_convlstm_backward(const LstmDev& d, cudaStream_t st);
int run_convlstm_dx(const LstmDev& d, float* dx, cudaStream_t st);
int run_convlstm_backward_rbp(const LstmDev& d, int num_iter, cudaStream_t st);

}  // namespace ptk
