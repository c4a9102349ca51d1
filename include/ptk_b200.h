/*
 * ptk_b200.h -- C ABI of the B200-native PathTracker recurrent core (libptk_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of pathtracker-code/pathtracker-models:
 * the per-timestep recurrent circuit of models/InT.py (InT), models/ffhgru_hierarchy.py (FFhGRU)
 * and models/convlstm.py (ConvLSTM), forward and BPTT backward.  The reference has no native
 * code and no FFI: the interface being replaced is the *body* of the nn.Module.forward methods
 * (and the autograd graph they build).  Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - plain C, no torch types.  All pointers are CALLER-OWNED DEVICE memory (fp32 unless noted)
 *     on the device that is current when the call is made; `stream` is a cudaStream_t passed as
 *     void* (NULL = legacy default stream).  The library allocates nothing persistent, never
 *     synchronises the device, and is re-entrant per (device, stream).
 *   - return 0 on success, a PTK_E* code otherwise; ptk_last_error() returns a thread-local
 *     human-readable message for the last failure on the calling thread.
 *   - state-sized tensors are [B, C, H, W] fp32, contiguous (NCHW), C = 32.
 *   - parameters travel as ONE flat fp32 vector in the canonical order of ptk_int_param_offsets
 *     (the reference's state_dict order without the dead `unit1.w`); gradients come back in the
 *     same layout, which is also the bucket handed to the NCCL all-reduce.
 */
#ifndef PTK_B200_H
#define PTK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTK_VERSION 100

#if defined(__GNUC__)
#define PTK_API __attribute__((visibility("default")))
#else
#define PTK_API
#endif

enum {
  PTK_OK = 0,
  PTK_EINVAL = 1,      /* bad argument / unsupported shape (message says which)          */
  PTK_ENOMEM = 2,      /* caller-provided workspace / saved buffer too small             */
  PTK_ECUDA = 3,       /* a CUDA runtime call or launch failed                            */
  PTK_EUNSUPPORTED = 4 /* variant not built (e.g. kernel_size other than 7)               */
};

enum { PTK_CELL_INT = 0,   /* models/InT.py:157            gated_inhibition = inhibition  */
       PTK_CELL_HGRU = 1   /* models/ffhgru_hierarchy.py:147  gated_inhibition = att_gate */ };
enum { PTK_ACT_SOFTPLUS = 0, /* F.softplus beta=1 threshold=20 (InT.py:184 default nl)    */
       PTK_ACT_TANH = 1      /* utils/engine.py:138-146 InT_tanh                          */ };
enum { PTK_CONV_FP32 = 0,     /* strict fp32 FFMA convolutions (parity-gated default)     */
       PTK_CONV_BF16X3 = 1,   /* 3-pass split-bf16 tcgen05 tiles, fp32 accumulate         */
       PTK_CONV_BF16 = 2      /* single-pass bf16 tcgen05 tiles (NOT fp32-parity: ~1e-3 ..1e-2 on gradients; reported separately) */ };

typedef struct PtkDims {
  int32_t B, T, H, W;  /* clip batch / frames / frame size.  x is [B,3,T,H,W]              */
  int32_t C;           /* hidden channels ("dimensions"); must be 32 (utils/engine.py:75)  */
  int32_t K;           /* horizontal kernel size; must be 7  (utils/engine.py:75)          */
  int32_t cell;        /* PTK_CELL_*                                                       */
  int32_t act;         /* PTK_ACT_*                                                        */
  int32_t no_inh;      /* InT(no_inh=True), models/InT.py:167-168                          */
  int32_t conv_mode;   /* PTK_CONV_*                                                       */
} PtkDims;

/* Number of parameter tensors in the flat vector and their canonical order. */
#define PTK_INT_NPARAM 30
enum {
  PTK_P_PREPROC_W = 0, PTK_P_PREPROC_B, PTK_P_W_EXC, PTK_P_W_INH, PTK_P_ALPHA, PTK_P_MU,
  PTK_P_GAMMA, PTK_P_KAPPA,
  PTK_P_AW_W, PTK_P_AW_B, PTK_P_AU_W, PTK_P_AU_B, PTK_P_IW_W, PTK_P_IW_B, PTK_P_IU_W, PTK_P_IU_B,
  PTK_P_EW_W, PTK_P_EW_B, PTK_P_EU_W, PTK_P_EU_B,
  PTK_P_BN0_W, PTK_P_BN0_B, PTK_P_BN1_W, PTK_P_BN1_B,
  PTK_P_READOUT_W, PTK_P_READOUT_B, PTK_P_TARGET_W, PTK_P_TARGET_B, PTK_P_DENSE_W, PTK_P_DENSE_B
};

PTK_API int ptk_version(void);
PTK_API const char* ptk_last_error(void);

/* Element offsets (PTK_INT_NPARAM + 1 entries; last = total count) of each parameter tensor in
 * the flat vector.  Shapes are those of rCell.__init__ / InT.__init__ (models/InT.py:63-143,
 * 184-208): preproc [C,3]+[C]; w_exc,w_inh [C,C,K,K]; alpha,mu,gamma,kappa [C]; six 1x1 gates
 * [C,C]+[C]; two BatchNorm affine pairs [C]; readout_conv [C]+[1]; target_conv [2,5,5]+[1];
 * readout_dense [1]+[1]. */
PTK_API int ptk_int_param_offsets(const PtkDims* d, int64_t* offsets);

/* Bytes of the checkpoint buffer forward fills for backward: I_t,E_t (t=0..T), and per step the two
 * pre-BatchNorm conv outputs, the three gates (A, G_i, G_e) and the 4*C batch statistics.  This replaces
 * the autograd tape of models/InT.py:223-235 (21 state-sized tensors/step in the reference). */
PTK_API size_t ptk_int_saved_bytes(const PtkDims* d);
PTK_API size_t ptk_int_workspace_bytes(const PtkDims* d);            /* scratch of ptk_int_forward */
/* scratch of ptk_int_backward: the forward scratch plus, in the tensor-core modes, every step's dY0 / dY1 / Ge
 * ([T][B,C,H,W] each) for the 7x7 weight gradients, which are taken in one launch after the reverse loop */
PTK_API size_t ptk_int_backward_workspace_bytes(const PtkDims* d);

/* Forward of InT.forward / FFhGRU.forward (models/InT.py:210-245,
 * models/ffhgru_hierarchy.py:211-276): stem Conv3d(3->C)+nl, T cell steps (rCell.forward
 * models/InT.py:145-179), readout.  x: [B,3,T,H,W] with arbitrary element strides
 * (utils/engine.py:222-243 hands over a channels_last_3d-strided tensor).
 *   logits  [B]            out
 *   saved   ptk_int_saved_bytes, or NULL for inference (no backward possible)
 *   gates   [B,T,C,H,W]    optional (testmode=True, InT.py:230-233), else NULL
 *   states  [B,T,1,H,W]    optional (testmode=True), else NULL */
PTK_API int ptk_int_forward(const PtkDims* d, const float* x, const int64_t x_strides[5],
                    const float* params, float* logits, void* saved, float* gates, float* states,
                    void* workspace, size_t workspace_bytes, void* stream);

/* BPTT backward (replaces loss.backward() through the graph of InT.forward, mainclean.py:204).
 *   dlogits [B] in;  dparams flat vector out (overwritten, canonical layout);
 *   dx optional [B,3,T,H,W] contiguous (viz_model_att.py:189-196), else NULL. */
PTK_API int ptk_int_backward(const PtkDims* d, const float* x, const int64_t x_strides[5],
                     const float* params, const void* saved, const float* dlogits,
                     float* dparams, float* dx, void* workspace, size_t workspace_bytes,
                     void* stream);

/* On-device synthetic PathTracker clips (stands in for utils/TFRDataset.py:6-53; data contract
 * utils/TFRDataset.py:17-25 + utils/engine.py:222-226: values in {0,1}, x [B,3,T,H,W] fp32
 * contiguous, y [B] fp32 labels).  Counter-based: sample i depends only on (seed, first_index+i),
 * so every rank draws a disjoint deterministic shard. */
PTK_API int ptk_synth_pathtracker(uint64_t seed, uint64_t first_index, int32_t B, int32_t T, int32_t H,
                          int32_t W, int32_t n_distractors, int32_t speed, float* x, float* y,
                          void* stream);

/* ---- The step either side of the path -------------------------------------------------
 * Loss + metrics in one launch: criterion = BCEWithLogitsLoss (mainclean.py:156,189) and acc_scores /
 * metric_scores (utils/misc_functions.py:12-45; note the LOGIT is compared with 0.5 and recall divides by the
 * batch size, exactly as written there).  logits, target: [B] fp32 device.  dlogits (optional, [B]) receives
 * d(mean loss)/d(logit).  out5 (device) = {loss, bal_acc*100, precision, recall, f1}.  No host synchronisation. */
PTK_API int ptk_bce_scores(int32_t B, const float* logits, const float* target, float* dlogits, float* out5,
                   void* stream);
/* torch.optim.Adam(lr) with PyTorch defaults (mainclean.py:157) over ONE flat fp32 vector, in place, plus
 * optimizer.zero_grad() (mainclean.py:205-206) when zero_grad != 0.  step = 1 for the first update. */
PTK_API int ptk_adam_step(int64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, float lr,
                  float beta1, float beta2, float eps, int64_t step, int32_t zero_grad, void* stream);

/* ---- ConvLSTM baseline (models/convlstm.py) --------------------------------------------------------------
 * Flat parameter vector in the reference's state_dict order: conv0.{weight[25,1,7,7],bias}, then for each gate
 * k in (i,f,c,o): unit1.Wx<k>.{weight[25,25,K,K],bias[25]}, unit1.Wh<k>.weight[25,25,K,K]; bn.{weight,bias}[25];
 * conv6.{weight[2,25,1,1],bias[2]}.  ptk_convlstm_param_offsets fills 19 entries (18 tensors + total). */
typedef struct PtkLstmDims {
  int32_t B, T, H, W;  /* x is [B,1,H,W]; T = timesteps (models/convlstm.py:138 loops self.timesteps) */
  int32_t K;           /* filt_size: 15 (default, models/convlstm.py:95) or 7                         */
  int32_t conv_mode;   /* PTK_CONV_FP32: every convolution in strict fp32 on CUDA cores; PTK_CONV_BF16X3: the recurrent
                        * convolution of h and its transpose (the two per-step convolutions of the time loops) as
                        * 3-pass split-bf16 tcgen05 implicit GEMMs, same parity gate; PTK_CONV_BF16: the same kernels with ONE
                        * bf16 product per k-step (the bf16 variant: a third of the MMAs, gradients ~1e-2, not fp32-parity) */
  int32_t ckpt_ring;   /* 0: forward keeps h_t, c_t, gates_t of EVERY step (BPTT needs them all).
                        * 1: constant-memory checkpoints for recurrent back-propagation (models/convlstm.py:124-135 runs the
                        *    first T-1 steps under no_grad): two-slot rings, so only steps T-1 and T survive;
                        *    ptk_convlstm_saved_bytes no longer grows with T.  ptk_convlstm_backward then refuses. */
  int32_t rbp;         /* 1: grad_method='rbp' -- h_{T-1}, c_{T-1} are detached leaves (models/convlstm.py:130-131), so the
                        *    second product of the Jacobian penalty is f_T alone and the penalty's gradient
                        *    (ptk_convlstm_backward_penalty) reaches the parameters through the last step only.
                        * 0: 'bptt' -- they are live graph tensors (:140-142): the product also holds the path
                        *    c_{T-1} -> h_{T-1} -> gates_T -> c_T and the penalty's gradient runs back through every step. */
} PtkLstmDims;
PTK_API int ptk_convlstm_param_offsets(const PtkLstmDims* d, int64_t* offsets);
PTK_API size_t ptk_convlstm_saved_bytes(const PtkLstmDims* d);
PTK_API size_t ptk_convlstm_workspace_bytes(const PtkLstmDims* d);
/* ConvLSTM.forward up to `output = conv6(bn(h_T))` (models/convlstm.py:116-147, bptt branch); the criterion
 * (:148) stays with the caller.  penalty: optional [B,25,H,W] out, the Jacobian penalty of :152-163
 * (training mode; a constant for the caller unless ptk_convlstm_backward_penalty is used); needs T >= 2.  saved: ptk_convlstm_saved_bytes, always required. */
PTK_API int ptk_convlstm_forward(const PtkLstmDims* d, const float* x, const float* params, float* output,
                                 float* penalty, void* saved, void* workspace, size_t workspace_bytes, void* stream);
/* BPTT: doutput [B,2,H,W] -> dparams (flat, overwritten). */
PTK_API int ptk_convlstm_backward(const PtkLstmDims* d, const float* x, const float* params, const void* saved,
                                  const float* doutput, float* dparams, void* workspace, size_t workspace_bytes,
                                  void* stream);

/* Recurrent back-propagation (`grad_method='rbp'`: dummyhgru.backward models/convlstm.py:9-54, forward branch
 * :124-135): Neumann series g = sum_k (J^T)^k dL/dh_T on the last step's Jacobian, truncated after num_iter
 * terms or by the reference's norm tests (|g| > 1, |v_k| > |v_{k-1}|, |v_k| < 1e-9); parameter gradients = one
 * cell-step backward with dh_T := g.  The stop test runs on the device (no synchronisation); terms_out
 * (optional, device, 1 float) receives the number of accepted terms.  Reads only the last two steps of the
 * checkpoints: works with ckpt_ring = 1 (O(1) memory) as well as with the full BPTT buffer. */
PTK_API int ptk_convlstm_backward_rbp(const PtkLstmDims* d, const float* x, const float* params, const void* saved,
                                      const float* doutput, int32_t num_iter, float* dparams, float* terms_out,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* jacobian_penalty=True (models/convlstm.py:152-163 with create_graph=True): gradient of
 *   total = loss + <dpenalty, penalty>      (the training loop uses loss + 10 * penalty.mean(), mainclean.py:192-196)
 * with the penalty INSIDE the graph, i.e. the double back-propagation autograd performs in the reference, in closed form:
 * second-order seeds at the gate pre-activations of steps T and T-1, on c_{T-1} and directly on the Wh weight gradient
 * (two extra forward convolutions, two transposes, two weight-gradient passes), then the unchanged reverse loop.
 * d->rbp selects the loss gradient: 0 = BPTT (as ptk_convlstm_backward), 1 = recurrent back-propagation (as
 * ptk_convlstm_backward_rbp; num_iter / terms_out as there, ignored for BPTT).  dpenalty: [B,25,H,W]. */
PTK_API int ptk_convlstm_backward_penalty(const PtkLstmDims* d, const float* x, const float* params, const void* saved,
                                          const float* doutput, const float* dpenalty, int32_t num_iter, float* dparams,
                                          float* terms_out, void* workspace, size_t workspace_bytes, void* stream);

/* d(loss)/d(x) [B,1,H,W] of the ConvLSTM (the reference is differentiable w.r.t. the image through autograd: x feeds conv0,
 * models/convlstm.py:118): dx = conv_transpose2d(d(loss)/d(u), conv0.weight, padding = 3).  Call it right after
 * ptk_convlstm_backward / ptk_convlstm_backward_rbp on the same stream with the SAME workspace, untouched in between:
 * d(loss)/d(u) is read from where that call left it. */
PTK_API int ptk_convlstm_input_grad(const PtkLstmDims* d, const float* params, float* dx, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* ---- ConvGRU baseline (models/kys.py; `--model gru`, utils/engine.py:147-153: dimensions = 64, kernel 7) ----------
 * Flat parameter vector (14 tensors): preproc.{weight[64,3,1,1,1],bias}, unit1.conv_reset.{weight[64,128,7,7],bias},
 * unit1.conv_update.{weight,bias}, unit1.conv_state_new.{weight,bias}, readout_conv.{weight[1,64,1,1],bias},
 * target_conv.{weight[1,2,5,5],bias}, readout_dense.{weight[1,1],bias}; ptk_gru_param_offsets fills 15 entries.
 * ptk_gru_forward = GRU.forward (:104-135): stem, the T ConvGRUCell steps (:36-48) and the readout -> logits [B].
 * saved (always required) holds h_t (T+1), reset gates, update gates and candidates (T each) as [.][B,64,H,W] blocks
 * in that order: the testmode dumps are read from it.
 * ptk_gru_backward: dlogits [B] -> dparams (flat, overwritten) and optionally dx [B,3,T,H,W] contiguous. */
typedef struct PtkGruDims {
  int32_t B, T, H, W;
  int32_t C;           /* 64  */
  int32_t K;           /* 7   */
  int32_t conv_mode;   /* PTK_CONV_FP32: generic strict-fp32 CUDA-core convolutions (ptk_convg.cuh).
                        * PTK_CONV_BF16X3 (default): the stacked 128->128 [r|u] convolution, the 128->64 candidate
                        *   convolution, their transposes and weight gradients on the wide tcgen05 kernels shared with the
                        *   ConvLSTM (ptk_convk_tc.cu, ptk_wgradk_tc.cu), 3-pass split-bf16, same parity gate.
                        * PTK_CONV_BF16: single-pass bf16 as 32x32-channel block launches on the first-generation InT
                        *   kernels (ptk_conv_tc.cu, ptk_wgrad_tc.cu); not an fp32-parity mode                          */
} PtkGruDims;
PTK_API int ptk_gru_param_offsets(const PtkGruDims* d, int64_t* offsets);
PTK_API size_t ptk_gru_saved_bytes(const PtkGruDims* d);
PTK_API size_t ptk_gru_workspace_bytes(const PtkGruDims* d);
PTK_API int ptk_gru_forward(const PtkGruDims* d, const float* x, const int64_t x_strides[5], const float* params,
                            float* logits, void* saved, void* workspace, size_t workspace_bytes, void* stream);
PTK_API int ptk_gru_backward(const PtkGruDims* d, const float* x, const int64_t x_strides[5], const float* params,
                             const void* saved, const float* dlogits, float* dparams, float* dx, void* workspace,
                             size_t workspace_bytes, void* stream);

/* One horizontal-connection convolution on its own (test / measurement entry; the time loop calls the
 * same kernels internally): out (+)= conv2d(in, weight, padding=3) (transpose=0, models/InT.py:161,172)
 * or conv_transpose2d(in, weight, padding=3) (transpose=1, its data gradient).  in/out [B,32,H,W],
 * weight [32,32,7,7] as stored by torch; stats optional [2][32] out: batch mean and 1/sqrt(var+1e-3)
 * of the convolution result (the BatchNorm that follows it). */
/* One wide operator of the ConvLSTM time loops on its own, on the tcgen05 kernels (test / measurement entry; the time
 * loops call the same launchers).  weights: the four recurrent banks Wh{i,f,c,o} (models/convlstm.py:66-79) contiguous,
 * [4][25][25][K][K].  op 0: out [B,100,H,W] = stacked conv(in [B,25,H,W]) (+ aux [B,100,H,W] if given);
 * op 1: out [B,25,H,W] = transposed conv(in [B,100,H,W]);  op 2: out [4][25][25][K][K] = weight gradient for
 * in = h [B,25,H,W], aux = dY [B,100,H,W]. */
PTK_API size_t ptk_lstm_wide_op_workspace_bytes(int32_t K);
PTK_API int ptk_lstm_wide_op(int32_t B, int32_t H, int32_t W, int32_t K, int32_t op, const float* in, const float* aux,
                             const float* weights, float* out, void* workspace, size_t workspace_bytes, void* stream);
PTK_API size_t ptk_conv7_workspace_bytes(int32_t B, int32_t H, int32_t W);
PTK_API int ptk_conv7(int32_t B, int32_t H, int32_t W, int32_t conv_mode, int32_t transpose, int32_t accumulate,
                      const float* in, const float* weight, float* out, float* stats, void* workspace,
                      size_t workspace_bytes, void* stream);
/* The 7x7 weight gradient on its own (test / measurement entry), summed over T (input, output-gradient) pairs:
 *   dweight[oc][ic][ky][kx] = sum_t sum_{b,y,x} dy[t][b,oc,y,x] * in[t][b,ic,y+ky-3,x+kx-3]
 * i.e. the gradient autograd accumulates into w_inh / w_exc through F.conv2d(..., padding=3) (models/InT.py:161,172) over
 * the time loop.  in, dy: [T][B,32,H,W] contiguous.  Tensor-core modes only (ptk_wgrad_ts.cu). */
PTK_API size_t ptk_wgrad7_workspace_bytes(int32_t T, int32_t B, int32_t H, int32_t W);
PTK_API int ptk_wgrad7(int32_t T, int32_t B, int32_t H, int32_t W, int32_t conv_mode, const float* in, const float* dy,
                       float* dweight, void* workspace, size_t workspace_bytes, void* stream);


/* Measurement hooks (used by bench.py; off by default, no effect on results).
 * ptk_profile_start(): from now on every kernel-class launch is bracketed by cudaEvents on its stream.
 * ptk_profile_stop(ms, count): synchronises those events; fills ms[k] / count[k] for the 8 classes
 *   0 conv fwd, 1 conv dgrad, 2 wgrad, 3 point-wise fwd, 4 point-wise bwd, 5 bn-apply, 6 misc, 7 gate wgrad.
 * ptk_launch_count(): kernels launched by this library in this process so far. */
PTK_API void ptk_profile_start(void);
PTK_API int ptk_profile_stop(double* ms, long long* count);
PTK_API long long ptk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PTK_B200_H */
