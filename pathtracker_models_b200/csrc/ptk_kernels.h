// Internal kernel argument structs + host launchers shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptk {

struct ConvArgs {
  const float* in;     // [B,32,H,W]
  const float* wT;     // [32 ic][49 taps][32 oc]
  float* out;          // [B,32,H,W]
  int B, H, W;
  int accumulate;      // out += conv (dgrad into an existing gradient)
  float* part;         // [nCTA][32][2] per-CTA (mean, M2) or nullptr (no statistics)
  unsigned* counter;   // arrival counter for the last-CTA fold (self-resetting)
  float* stats_out;    // [2][32] mean, rstd
  int mode;            // PTK_CONV_FP32: wT is fp32 [ic][tap][oc]; PTK_CONV_BF16X3: wT is the bf16 hi/lo tap stream
  int pdl = 0;         // launched inside the time loop: programmatic dependent launch allowed (ptk_common.cuh)
  long long in_bs = 0, out_bs = 0;   // tensor-core path: elements between samples (0 = 32*H*W); lets a 32-channel block
                                     // of a wider [B,Cw,H,W] tensor be convolved in place (ConvGRU, ptk_gru.cu)
  int v2 = 0;          // tensor-core path: wT is the ptk_conv_tc2.cu weight stream (launch_prep_weights_tc2) -> conv7_tc2_kernel
  int defer_stats = 0; // conv7_tc2: write the per-CTA partials only; the CONSUMER kernel folds them (consumer_stats_fold,
                       // ptk_conv_stats.cuh) -- no fence / atomic hand-shake / last-CTA fold at the end of the convolution
  // conv7_tc2 only: BatchNorm backward fused into the operand staging (the data-gradient convolutions of the reverse loop).
  // `in` then holds dV, the gradient wrt the BatchNorm OUTPUT; the kernel forms
  //   dY = w*rstd * (dV - mean(dV) - xhat * mean(dV*xhat)),  xhat = (Y - mean) * rstd        (eqs b3 / b9)
  // on the fly, convolves dY and writes the interior rows of dY to bn_out (read later by the deferred weight gradient).
  const float* bn_y = nullptr;      // saved conv output Y of the forward step [B,32,H,W]; non-null enables the fusion
  const float* bn_stat = nullptr;   // forward statistics: mean[32], rstd[32]
  const float* bn_tot = nullptr;    // backward sums of this step: mean(dV)[32], mean(dV*xhat)[32]
  const float* bn_w = nullptr;      // BatchNorm weight [32]
  float* bn_out = nullptr;          // dY [B,32,H,W]; must not alias `in`
};

struct WgradArgs {
  const float* in;     // conv input [B,32,H,W]
  const float* dy;     // gradient wrt conv output [B,32,H,W]
  float* wacc;         // [4*G][49][256] per-CTA accumulators
  int B, H, W;
  int first;           // overwrite instead of accumulate
  int mode;            // PTK_CONV_BF16: single pass (hi parts only)
  long long in_bs = 0, dy_bs = 0;    // elements between samples (0 = 32*H*W), as in ConvArgs
};

// weight gradient of the 7x7 convolutions summed over a whole sequence in ONE launch (ptk_wgrad_ts.cu): accumulators
// resident in TMEM, A operand (dY) fed from TMEM, persistent CTAs; slots [ctas][49][32 oc][32 ic] written once
struct WgradTsArgs {
  const float* in;      // conv inputs  [T][B,32,H,W] (time stride in_ts elements)
  long long in_ts;
  const float* dy;      // gradients wrt the conv outputs [T][B,32,H,W] (time stride dy_ts elements)
  long long dy_ts;
  float* wacc;          // wgrad_ts_acc_floats() floats
  int T, B, H, W;
  int mode;             // PTK_CONV_BF16: single pass (hi parts only)
  int pdl = 0;
};
int wgrad_ts_ctas(int T, int B, int H, int W);
size_t wgrad_ts_acc_floats(int T, int B, int H, int W);
int launch_wgrad7_ts(const WgradTsArgs& a, cudaStream_t st);
int launch_wgrad_ts_fold(const float* wacc, int nslots, float* dw, cudaStream_t st);

// wide recurrent convolutions on tcgen05 (ptk_convk_tc.cu; ConvLSTM): fp32 NCHW in / out, prepared bf16 hi/lo weight stream
struct ConvKArgs {
  const float* in;      // [B, Cin, H, W]
  const void* wT;       // launch_prep_weights_k stream
  const float* bias;    // [Cout] or null
  const float* addend;  // [B, Cout, H, W] or null
  float* out;           // [B, Cout, H, W]
  int B, H, W, Cin, Cout, accumulate;
  const float* skip = nullptr;   // RBP device state: skip[1] != 0 => do nothing
  int dbg = 0;                   // timing experiments (PTK_CONVK_EXP): 1 = no epilogue stores, 2 = no MMAs
  int single = 0;                // 1: conv_mode = bf16, ONE product per k-step (Ih x Wh) instead of the 3-pass split: not fp32-parity
};
template <int KSZ> size_t convk_tc_weight_bytes(int transposed);
template <int KSZ> int launch_prep_weights_k(const float* const* w, int transposed, void* dst, cudaStream_t st);
template <int KSZ> int launch_convk_tc(const ConvKArgs& a, int transposed, cudaStream_t st);
size_t convk_tc_stream_bytes(int ks, int nt, int nblk);
int launch_prep_weights_k_generic(const float* const* w, int nbank, int ocb, int icb, int ks, int nt, int nblk, int transposed,
                                  void* dst, cudaStream_t st);
template <int KSZ, int NT> int launch_convk_tc_n(const ConvKArgs& a, cudaStream_t st);

// weight gradient of the same convolution on tcgen05 (ptk_wgradk_tc.cu): slots [groups][KS*KS][128][32], folded at the end
struct WgradKArgs {
  const float* in;      // [B, Cin <= 32, H, W]
  const float* dy;      // [B, Cout <= 128, H, W]
  float* slots;         // wgradk_tc_slot_floats(KS) floats
  int B, H, W, Cin, Cout, accumulate;
  const float* skip = nullptr;
  long long in_bs = 0;  // elements between samples of `in` (0 = Cin*H*W): a 32-channel block of a wider tensor
  const void* dyc = nullptr;   // dy already split and laid out per tile by launch_wgradk_prep_dy (then `dy` is not read)
  int single = 0;              // 1: conv_mode = bf16, one product per k-step
};
// dy [B,Cout,H,W] fp32 -> per 4x32-pixel tile the kernel's shared-memory image (bf16 hi/lo planes): the weight-gradient
// CTAs of all kernel rows (and input blocks) then fetch a tile with ONE bulk copy instead of converting it again each
size_t wgradk_tc_dyc_bytes(int B, int H, int W);
int launch_wgradk_prep_dy(const float* dy, int B, int H, int W, int Cout, void* dyc, const float* skip, cudaStream_t st);
int wgradk_tc_groups(int ks);
size_t wgradk_tc_slot_floats(int ks);
template <int KSZ> int launch_wgradk_tc(const WgradKArgs& a, cudaStream_t st);
int launch_fold_wgradk(const float* slots, int ks, int nbank, int ocb, int icb, float* dst0, long long stride_between, cudaStream_t st,
                       int ic0 = 0, int ic_total = 0);

int conv_num_ctas(int B, int H, int W);
int launch_conv7(const ConvArgs& a, cudaStream_t st);
int launch_wgrad7(const WgradArgs& a, int groups, cudaStream_t st);
int launch_conv7_tc(const ConvArgs& a, cudaStream_t st);
// second-generation tensor-core convolution (ptk_conv_tc2.cu: weights as the A operand, chunked double-buffered epilogue)
bool conv_v2_enabled();            // PTK_CONV_V1=1 selects the first-generation kernel (A/B measurements)
size_t conv7_tc2_weight_bytes();   // per weight bank and direction
int launch_conv7_tc2(const ConvArgs& a, cudaStream_t st);
int launch_prep_weights_tc2(const float* w, void* wf, void* wd, cudaStream_t st, int oc_stride = 0);
int launch_prep_weights_tc(const float* w, void* wf, void* wd, cudaStream_t st, int oc_stride = 0);   // 0 = 32*49 (dense)
int launch_wgrad7_tc(const WgradArgs& a, cudaStream_t st);   // wacc: [nCTA][49][32 oc][32 ic]
size_t wgrad_tc_acc_floats(int B, int H, int W);
#if PTK_DEBUG_STAMPS   // tuning builds only (-DPTK_DEBUG_STAMPS=1): in-kernel cycle / globaltimer stamps
int read_pw_stamps(unsigned long long* out);
int read_tc_debug(long long* out);
int read_wt_debug(long long* out);
int read_tc_times(unsigned long long* out, int n);
int read_loop_stamps(long long* out);
#endif

}  // namespace ptk
