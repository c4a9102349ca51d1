// C ABI of libptk_b200 (see include/ptk_b200.h): argument checking, workspace carving, dispatch.
#include <stdarg.h>
#include <string.h>

#include "ptk_common.cuh"
#include "ptk_convlstm.h"
#include "ptk_gru.h"
#include "ptk_int.h"

#include <stdlib.h>
namespace ptk {
bool conv_v2_enabled() {
  static const bool on = [] { const char* e = getenv("PTK_CONV_V1"); return !(e && e[0] == '1'); }();
  return on;
}
// floats per prepared 7x7 weight bank and direction: fp32 re-layout, or the larger of the two bf16 hi/lo streams
static size_t wbank_floats() {
  const size_t a = (size_t)C * C * TAPS, b = (conv7_tc2_weight_bytes() + 3) / 4;
  return a > b ? a : b;
}
int int_loop_mode() {
  static const int mode = [] { const char* e = getenv("PTK_FWD_LOOP"); return e ? atoi(e) : 0; }();
  return mode;
}
bool stats_deferred() {
  static const bool on = [] { const char* e = getenv("PTK_STATS_DEFER"); return e && e[0] == '1'; }();
  return on;
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("PTK_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}


static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int check_dims(const PtkDims* d, const char* who) {
  if (!d) { set_error("%s: dims is NULL", who); return PTK_EINVAL; }
  if (d->C != C || d->K != KS) {
    set_error("%s: only dimensions=32, kernel_size=7 are built (got C=%d K=%d); utils/engine.py:75 never uses others",
              who, d->C, d->K);
    return PTK_EUNSUPPORTED;
  }
  if (d->B < 1 || d->T < 1 || d->H < 16 || d->W < 32 || d->H % 16 || d->W % 32) {
    set_error("%s: unsupported shape B=%d T=%d H=%d W=%d (need B,T >= 1, H %% 16 == 0, W %% 32 == 0)", who, d->B,
              d->T, d->H, d->W);
    return PTK_EINVAL;
  }
  // the readout kernels keep two zero-padded maps (+ one plain map in backward) of a whole frame in shared memory
  if (((size_t)2 * (d->H + 4) * (d->W + 4) + (size_t)d->H * d->W) * sizeof(float) > (size_t)227 * 1024) {
    set_error("%s: frame %dx%d too large: the readout tile (2 (H+4)(W+4) + H W floats) exceeds the 227 KB of shared memory per CTA",
              who, d->H, d->W);
    return PTK_EUNSUPPORTED;
  }
  if (d->cell != PTK_CELL_INT && d->cell != PTK_CELL_HGRU) { set_error("%s: bad cell %d", who, d->cell); return PTK_EINVAL; }
  if (d->act != PTK_ACT_SOFTPLUS && d->act != PTK_ACT_TANH) { set_error("%s: bad act %d", who, d->act); return PTK_EINVAL; }
  if (d->no_inh && d->cell != PTK_CELL_INT) { set_error("%s: no_inh exists for the InT cell only (models/InT.py:167)", who); return PTK_EINVAL; }
  if (d->conv_mode != PTK_CONV_FP32 && d->conv_mode != PTK_CONV_BF16X3 && d->conv_mode != PTK_CONV_BF16) { set_error("%s: bad conv_mode %d", who, d->conv_mode); return PTK_EINVAL; }
  return PTK_OK;
}

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base((char*)p) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* r = base ? (T*)(base + off) : nullptr;
    off += n * sizeof(T);
    return r;
  }
};

static size_t state_elems(const PtkDims* d) { return (size_t)d->B * C * d->H * d->W; }

// carve the workspace into the IntDev scratch pointers; returns bytes needed
static size_t carve_workspace(const PtkDims* dm, void* ws, IntDev* d, bool with_bwd) {
  Carver cv(ws);
  const size_t S = state_elems(dm);
  const int pw = pw_grid(dm->B, dm->H, dm->W);
  float* w = cv.take<float>(4 * wbank_floats());
  float* gt = cv.take<float>(3 * 6 * C * C);   // gateT, gateTF, gateF
  float* Ge = cv.take<float>(S);
  float* big = cv.take<float>(14 * S);   // inference ring (9S) or backward scratch (14S)
  float* ring_stats = cv.take<float>(4 * C);
  float* bn_part = cv.take<float>((size_t)conv_num_ctas(dm->B, dm->H, dm->W) * C * 2);
  unsigned* counter = cv.take<unsigned>(4);
  float* bn_part2 = cv.take<float>((size_t)conv_num_ctas(dm->B, dm->H, dm->W) * C * 2);
  unsigned* loop_sync = cv.take<unsigned>(int_loop_sync_words(dm->B));
  uint8_t* loop_gates = cv.take<uint8_t>(int_loop2_gate_bytes());
  float* acc = cv.take<float>((size_t)PW_EW_CTAS * ACC_N);
  float* gacc = cv.take<float>((size_t)2 * GATE_CTAS * GACC_N);   // tensor-core modes: two slots per CTA (pixel halves)
  const size_t bn_sums_n = (size_t)PW_EW_CTAS * 2 * C > (size_t)dm->B * C * 4 ? (size_t)PW_EW_CTAS * 2 * C : (size_t)dm->B * C * 4;
  float* bn_sums = cv.take<float>(bn_sums_n);   // per-CTA partials of the warp-tile kernels, or [B*C][4] row sums of int_bwd_a_stream_kernel
  float* bn_tot = cv.take<float>(2 * C);
  double* bn_acc = cv.take<double>(4 * C);
  float* head = cv.take<float>((size_t)dm->B * HEAD_N);
  float* dgoal = cv.take<float>(S / C);
  size_t wacc_n = (size_t)4 * WGRAD_GROUPS * TAPS * 256;
  if (dm->conv_mode != PTK_CONV_FP32) wacc_n = wgrad_ts_acc_floats(dm->T, dm->B, dm->H, dm->W);
  float* wacc_exc = cv.take<float>(wacc_n);
  float* wacc_inh = cv.take<float>(wacc_n);
  // tensor-core modes, backward only: every step's dY0, dY1 and Ge stay live for the deferred weight gradient
  const bool defer = dm->conv_mode != PTK_CONV_FP32 && with_bwd;
  float* dys = cv.take<float>(defer ? 7 * (size_t)dm->T * S : 0);
  if (d) {
    d->Ge = Ge;
    d->bn_part = bn_part; d->counter = counter; d->pw_ctas = pw;
    d->bn_part2 = bn_part2; d->loop_sync = loop_sync; d->loop_gates = loop_gates;
    d->acc = acc; d->gacc = gacc; d->bn_sums = bn_sums; d->bn_tot = bn_tot; d->bn_acc = bn_acc;
    d->head = head; d->dgoal = dgoal;
    d->wacc_exc = wacc_exc; d->wacc_inh = wacc_inh;
    // backward scratch view of `big`
    d->gE = big; d->gI = big + S; d->geb = big + 2 * S; d->dAct = big + 3 * S; d->dGe = big + 4 * S;
    d->dX = big + 5 * S; d->dS = big + 6 * S; d->giS = big + 7 * S; d->aS = big + 8 * S; d->Xs = big + 9 * S;
    d->cgI = big + 10 * S; d->cGe = big + 11 * S;
    d->GeB[0] = Ge; d->GeB[1] = big + 12 * S; d->gebB[0] = d->geb; d->gebB[1] = big + 13 * S;
    d->dy_tstride = defer ? S : 0;
    d->bn_fused = defer && conv_v2_enabled();
    const size_t TS = defer ? (size_t)dm->T * S : 0;
    d->dYs[0] = dys; d->dYs[1] = dys + TS; d->GeT = dys + 2 * TS;
    d->gebT = dys + 3 * TS; d->giT = dys + 4 * TS; d->aT = dys + 5 * TS; d->XT = dys + 6 * TS;
    // inference ring view of `big` (only used when saved == NULL)
    if (d->ck_ring) {
      d->ckI = big; d->ckE = big + 2 * S; d->ck_stride = S;
      d->Y0 = big + 4 * S; d->Y1 = big + 5 * S; d->Gi = big + 6 * S; d->Gex = big + 7 * S; d->y_stride = 0;
      d->A = big + 8 * S; d->a_tstride = 0; d->a_bstride = (long long)C * dm->H * dm->W;
      d->stats = ring_stats;
    }
    d->w_inh_f = w; d->w_inh_d = w + wbank_floats();
    d->w_exc_f = w + 2 * wbank_floats(); d->w_exc_d = w + 3 * wbank_floats();
    for (int i = 0; i < 6; ++i) {
      d->gateT[i] = gt + i * C * C;
      d->gateTF[i] = gt + (6 + i) * C * C;
      d->gateF[i] = gt + (12 + i) * C * C;
    }
  }
  return (cv.off + 255) & ~(size_t)255;
}

static void bind_params(IntDev* d, const float* p) {
  const ParamOffsets po = param_offsets(C, KS);
  d->Wp = p + po.o[PTK_P_PREPROC_W]; d->bp = p + po.o[PTK_P_PREPROC_B];
  d->alpha = p + po.o[PTK_P_ALPHA]; d->mu = p + po.o[PTK_P_MU];
  d->gamma = p + po.o[PTK_P_GAMMA]; d->kappa = p + po.o[PTK_P_KAPPA];
  for (int i = 0; i < 6; ++i) {
    d->gate[i] = p + po.o[PTK_P_AW_W + 2 * i];
    d->gate_b[i] = p + po.o[PTK_P_AW_B + 2 * i];
  }
  d->bn_w[0] = p + po.o[PTK_P_BN0_W]; d->bn_b[0] = p + po.o[PTK_P_BN0_B];
  d->bn_w[1] = p + po.o[PTK_P_BN1_W]; d->bn_b[1] = p + po.o[PTK_P_BN1_B];
  d->Wr = p + po.o[PTK_P_READOUT_W]; d->br = p + po.o[PTK_P_READOUT_B];
  d->Wt = p + po.o[PTK_P_TARGET_W]; d->bt = p + po.o[PTK_P_TARGET_B];
  d->Wd = p + po.o[PTK_P_DENSE_W]; d->bd = p + po.o[PTK_P_DENSE_B];
}

static void bind_saved(IntDev* d, const PtkDims* dm, void* saved) {
  const size_t S = state_elems(dm);
  const size_t T = (size_t)dm->T;
  float* s = (float*)saved;
  d->ck_ring = 0;
  d->ckI = s; d->ckE = s + (T + 1) * S; d->ck_stride = S;
  d->Y0 = s + 2 * (T + 1) * S; d->Y1 = d->Y0 + T * S; d->A = d->Y1 + T * S; d->Gi = d->A + T * S;
  d->Gex = d->Gi + T * S; d->y_stride = S;
  d->a_tstride = (long long)S; d->a_bstride = (long long)C * dm->H * dm->W;
  d->stats = d->Gex + T * S;
}

static int prep(IntDev* d, const float* params, cudaStream_t st) {
  const ParamOffsets po = param_offsets(C, KS);
  IntGateSrc g;
  for (int i = 0; i < 6; ++i) g.w[i] = d->gate[i];
  const float* w_inh = params + po.o[PTK_P_W_INH];
  const float* w_exc = params + po.o[PTK_P_W_EXC];
  if (d->conv_mode != PTK_CONV_FP32) {
    // the four banks hold the bf16 hi/lo tap streams instead of fp32 re-layouts
    const bool v2 = conv_v2_enabled();
    int rc = v2 ? launch_prep_weights_tc2(w_inh, (void*)d->w_inh_f, (void*)d->w_inh_d, st)
                : launch_prep_weights_tc(w_inh, (void*)d->w_inh_f, (void*)d->w_inh_d, st);
    if (rc) return rc;
    if ((rc = v2 ? launch_prep_weights_tc2(w_exc, (void*)d->w_exc_f, (void*)d->w_exc_d, st)
                 : launch_prep_weights_tc(w_exc, (void*)d->w_exc_f, (void*)d->w_exc_d, st))) return rc;
    return run_prep_weights(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, g, (float*)d->gateT[0], st);
  }
  return run_prep_weights(w_inh, w_exc, (float*)d->w_inh_f, (float*)d->w_inh_d, (float*)d->w_exc_f,
                          (float*)d->w_exc_d, g, (float*)d->gateT[0], st);
}

}  // namespace ptk

using namespace ptk;

extern "C" {

int ptk_version(void) { return PTK_VERSION; }
const char* ptk_last_error(void) { return g_err; }

int ptk_int_param_offsets(const PtkDims* d, int64_t* offsets) {
  if (!d || !offsets) { set_error("ptk_int_param_offsets: NULL argument"); return PTK_EINVAL; }
  if (d->C < 1 || d->K < 1) { set_error("ptk_int_param_offsets: bad C/K"); return PTK_EINVAL; }
  const ParamOffsets po = param_offsets(d->C, d->K);
  memcpy(offsets, po.o, sizeof(po.o));
  return PTK_OK;
}

size_t ptk_int_saved_bytes(const PtkDims* d) {
  if (check_dims(d, "ptk_int_saved_bytes")) return 0;
  const size_t S = state_elems(d);
  return ((size_t)(2 * (d->T + 1) + 5 * d->T) * S + (size_t)d->T * 4 * C) * sizeof(float);
}

size_t ptk_int_workspace_bytes(const PtkDims* d) {
  if (check_dims(d, "ptk_int_workspace_bytes")) return 0;
  return carve_workspace(d, nullptr, nullptr, false);
}
size_t ptk_int_backward_workspace_bytes(const PtkDims* d) {
  if (check_dims(d, "ptk_int_backward_workspace_bytes")) return 0;
  return carve_workspace(d, nullptr, nullptr, true);
}

int ptk_int_forward(const PtkDims* dm, const float* x, const int64_t x_strides[5], const float* params,
                    float* logits, void* saved, float* gates, float* states, void* workspace,
                    size_t workspace_bytes, void* stream) {
  int rc = check_dims(dm, "ptk_int_forward");
  if (rc) return rc;
  if (!x || !x_strides || !params || !logits || !workspace) { set_error("ptk_int_forward: NULL argument"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_int_workspace_bytes(dm)) {
    set_error("ptk_int_forward: workspace too small (%zu < %zu)", workspace_bytes, ptk_int_workspace_bytes(dm));
    return PTK_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  IntDev d;
  memset(&d, 0, sizeof(d));
  d.B = dm->B; d.T = dm->T; d.H = dm->H; d.W = dm->W; d.conv_mode = dm->conv_mode;
  d.fv = FrameView{x, x_strides[0], x_strides[1], x_strides[2], x_strides[3], x_strides[4]};
  bind_params(&d, params);
  d.ck_ring = (saved == nullptr);
  if (saved) bind_saved(&d, dm, saved);
  carve_workspace(dm, workspace, &d, false);
  const size_t S = state_elems(dm);
  if (gates) {   // testmode: the attention gate of every step goes straight into the [B,T,C,H,W] output
    d.A = gates; d.a_tstride = (long long)C * d.H * d.W; d.a_bstride = (long long)d.T * C * d.H * d.W;
  }
  d.states = states;
  d.logits = logits;
  PTK_CUDA_OK(cudaMemsetAsync(d.counter, 0, 4 * sizeof(unsigned), st));
  PTK_CUDA_OK(cudaMemsetAsync(d.ckI, 0, S * sizeof(float), st));   // I_0 = 0 (models/InT.py:218)
  PTK_CUDA_OK(cudaMemsetAsync(d.ckE, 0, S * sizeof(float), st));   // E_0 = 0 (models/InT.py:217)
  if ((rc = prep(&d, params, st))) return rc;
  return run_int_forward(d, dm->no_inh ? 2 : dm->cell, dm->act, st);
}

int ptk_int_backward(const PtkDims* dm, const float* x, const int64_t x_strides[5], const float* params,
                     const void* saved, const float* dlogits, float* dparams, float* dx, void* workspace,
                     size_t workspace_bytes, void* stream) {
  int rc = check_dims(dm, "ptk_int_backward");
  if (rc) return rc;
  if (!x || !x_strides || !params || !saved || !dlogits || !dparams || !workspace) {
    set_error("ptk_int_backward: NULL argument (forward must have been given a `saved` buffer)");
    return PTK_EINVAL;
  }
  if (workspace_bytes < ptk_int_backward_workspace_bytes(dm)) {
    set_error("ptk_int_backward: workspace too small (%zu < %zu = ptk_int_backward_workspace_bytes)", workspace_bytes,
              ptk_int_backward_workspace_bytes(dm));
    return PTK_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  IntDev d;
  memset(&d, 0, sizeof(d));
  d.B = dm->B; d.T = dm->T; d.H = dm->H; d.W = dm->W; d.conv_mode = dm->conv_mode;
  d.fv = FrameView{x, x_strides[0], x_strides[1], x_strides[2], x_strides[3], x_strides[4]};
  bind_params(&d, params);
  bind_saved(&d, dm, const_cast<void*>(saved));
  carve_workspace(dm, workspace, &d, true);
  d.dlogits = dlogits; d.dparams = dparams; d.dx = dx;
  PTK_CUDA_OK(cudaMemsetAsync(d.counter, 0, 4 * sizeof(unsigned), st));
  if ((rc = prep(&d, params, st))) return rc;   // the workspace may have been reused since forward
  return run_int_backward(d, dm->no_inh ? 2 : dm->cell, dm->act, st);
}


int ptk_conv7(int32_t B, int32_t H, int32_t W, int32_t conv_mode, int32_t transpose, int32_t accumulate,
              const float* in, const float* weight, float* out, float* stats, void* workspace,
              size_t workspace_bytes, void* stream) {
  PtkDims dm = {B, 1, H, W, C, KS, 0, 0, 0, conv_mode};
  int rc = check_dims(&dm, "ptk_conv7");
  if (rc) return rc;
  if (!in || !weight || !out || !workspace) { set_error("ptk_conv7: NULL argument"); return PTK_EINVAL; }
  const size_t need = ptk_conv7_workspace_bytes(B, H, W);
  if (workspace_bytes < need) { set_error("ptk_conv7: workspace too small (%zu < %zu)", workspace_bytes, need); return PTK_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace);
  float* wf = cv.take<float>(wbank_floats());
  float* wd = cv.take<float>(wbank_floats());
  float* part = cv.take<float>((size_t)conv_num_ctas(B, H, W) * C * 2);
  unsigned* counter = cv.take<unsigned>(4);
  PTK_CUDA_OK(cudaMemsetAsync(counter, 0, 4 * sizeof(unsigned), st));
  const bool v2 = conv_mode != PTK_CONV_FP32 && conv_v2_enabled();
  if (conv_mode != PTK_CONV_FP32) rc = v2 ? launch_prep_weights_tc2(weight, wf, wd, st) : launch_prep_weights_tc(weight, wf, wd, st);
  else {
    IntGateSrc g = {};
    rc = run_prep_weights(weight, nullptr, wf, wd, nullptr, nullptr, g, nullptr, st);
  }
  if (rc) return rc;
  ConvArgs a{in, transpose ? wd : wf, out, B, H, W, accumulate, stats ? part : nullptr, counter, stats, conv_mode};
  a.v2 = v2;
  return launch_conv7(a, st);
}

#if PTK_DEBUG_STAMPS   // tuning builds only (tools/): not declared in include/ptk_b200.h, not in the shipped library
PTK_API int ptk_debug_tc_stamps(long long* out) { return read_tc_debug(out); }
PTK_API int ptk_debug_wt_stamps(long long* out) { return read_wt_debug(out); }
PTK_API int ptk_debug_tc_times(unsigned long long* out, int n) { return read_tc_times(out, n); }
PTK_API int ptk_debug_pw_stamps(unsigned long long* out) { return read_pw_stamps(out); }
PTK_API int ptk_debug_loop_stamps(long long* out) { return read_loop_stamps(out); }
#endif

size_t ptk_lstm_wide_op_workspace_bytes(int32_t K) {
  if (K != 7 && K != 15) return 0;
  Carver cv(nullptr);
  cv.take<uint8_t>(K == 15 ? convk_tc_weight_bytes<15>(1) : convk_tc_weight_bytes<7>(1));   // >= the forward stream
  cv.take<float>(wgradk_tc_slot_floats(K));
  return (cv.off + 255) & ~(size_t)255;
}
int ptk_lstm_wide_op(int32_t B, int32_t H, int32_t W, int32_t K, int32_t op, const float* in, const float* aux,
                     const float* weights, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (K != 7 && K != 15) { set_error("ptk_lstm_wide_op: filt_size %d not built (7 and 15 are)", K); return PTK_EUNSUPPORTED; }
  if (B < 1 || H < 8 || W < 32 || H % 8 || W % 32 || op < 0 || op > 2) { set_error("ptk_lstm_wide_op: unsupported B=%d H=%d W=%d op=%d", B, H, W, op); return PTK_EINVAL; }
  if (!in || !weights || !out || !workspace || (op == 2 && !aux)) { set_error("ptk_lstm_wide_op: NULL argument"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_lstm_wide_op_workspace_bytes(K)) { set_error("ptk_lstm_wide_op: workspace too small"); return PTK_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(workspace);
  uint8_t* wstream = cv.take<uint8_t>(K == 15 ? convk_tc_weight_bytes<15>(1) : convk_tc_weight_bytes<7>(1));
  float* slots = cv.take<float>(wgradk_tc_slot_floats(K));
  const size_t per = (size_t)25 * 25 * K * K;
  const float* banks[4] = {weights, weights + per, weights + 2 * per, weights + 3 * per};
  int rc;
  if (op == 2) {
    WgradKArgs w{in, aux, slots, B, H, W, 25, 100, 0, nullptr};
    if ((rc = K == 15 ? launch_wgradk_tc<15>(w, st) : launch_wgradk_tc<7>(w, st))) return rc;
    return launch_fold_wgradk(slots, K, 4, 25, 25, out, (long long)per, st);
  }
  if ((rc = K == 15 ? launch_prep_weights_k<15>(banks, op, wstream, st) : launch_prep_weights_k<7>(banks, op, wstream, st))) return rc;
  ConvKArgs k{in, wstream, nullptr, op == 0 ? aux : nullptr, out, B, H, W, op == 0 ? 25 : 100, op == 0 ? 100 : 25, 0, nullptr};
  return K == 15 ? launch_convk_tc<15>(k, op, st) : launch_convk_tc<7>(k, op, st);
}

size_t ptk_wgrad7_workspace_bytes(int32_t T, int32_t B, int32_t H, int32_t W) {
  if (T < 1 || B < 1 || H < 16 || W < 32 || H % 16 || W % 32) return 0;
  return (wgrad_ts_acc_floats(T, B, H, W) * sizeof(float) + 255) & ~(size_t)255;
}
int ptk_wgrad7(int32_t T, int32_t B, int32_t H, int32_t W, int32_t conv_mode, const float* in, const float* dy,
               float* dweight, void* workspace, size_t workspace_bytes, void* stream) {
  PtkDims dm = {B, T, H, W, C, KS, 0, 0, 0, conv_mode};
  int rc = check_dims(&dm, "ptk_wgrad7");
  if (rc) return rc;
  if (conv_mode == PTK_CONV_FP32) { set_error("ptk_wgrad7: tensor-core modes only (bf16x3, bf16)"); return PTK_EUNSUPPORTED; }
  if (!in || !dy || !dweight || !workspace) { set_error("ptk_wgrad7: NULL argument"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_wgrad7_workspace_bytes(T, B, H, W)) { set_error("ptk_wgrad7: workspace too small"); return PTK_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  const long long S = (long long)B * C * H * W;
  WgradTsArgs a{in, S, dy, S, (float*)workspace, T, B, H, W, conv_mode, 0};
  if ((rc = launch_wgrad7_ts(a, st))) return rc;
  return launch_wgrad_ts_fold((const float*)workspace, wgrad_ts_ctas(T, B, H, W), dweight, st);
}

size_t ptk_conv7_workspace_bytes(int32_t B, int32_t H, int32_t W) {
  if (B < 1 || H < 16 || W < 32 || H % 16 || W % 32) return 0;
  Carver cv(nullptr);
  cv.take<float>(wbank_floats());
  cv.take<float>(wbank_floats());
  cv.take<float>((size_t)conv_num_ctas(B, H, W) * C * 2);
  cv.take<unsigned>(4);
  return (cv.off + 255) & ~(size_t)255;
}


// ------------------------------------------------------------------------------------------------------------
// ConvLSTM (models/convlstm.py)
static int lstm_check(const PtkLstmDims* d, const char* who) {
  if (!d) { set_error("%s: dims is NULL", who); return PTK_EINVAL; }
  if (d->K != 7 && d->K != 15) { set_error("%s: filt_size %d not built (7 and 15 are)", who, d->K); return PTK_EUNSUPPORTED; }
  if (d->conv_mode != PTK_CONV_FP32 && d->conv_mode != PTK_CONV_BF16X3 && d->conv_mode != PTK_CONV_BF16) {
    set_error("%s: unknown conv_mode %d", who, d->conv_mode);
    return PTK_EUNSUPPORTED;
  }
  if (d->B < 1 || d->T < 1 || d->H < 16 || d->W < 32 || d->H % 16 || d->W % 32) {
    set_error("%s: unsupported shape B=%d T=%d H=%d W=%d", who, d->B, d->T, d->H, d->W);
    return PTK_EINVAL;
  }
  return PTK_OK;
}
static void lstm_offsets(int K, int64_t* o) {   // conv0.w, conv0.b, 4 x (Wx.w, Wx.b, Wh.w), bn.w, bn.b, conv6.w, conv6.b
  const int64_t n = (int64_t)25 * 25 * K * K;
  int i = 0;
  int64_t a = 0;
  o[i++] = a; a += 25 * 49;
  o[i++] = a; a += 25;
  for (int k = 0; k < 4; ++k) { o[i++] = a; a += n; o[i++] = a; a += 25; o[i++] = a; a += n; }
  o[i++] = a; a += 25;
  o[i++] = a; a += 25;
  o[i++] = a; a += 50;
  o[i++] = a; a += 2;
  o[i] = a;
}
static size_t lstm_carve(const PtkLstmDims* dm, void* saved, void* ws, LstmDev* d, size_t* saved_bytes) {
  const size_t P = (size_t)dm->H * dm->W, S = (size_t)dm->B * 25 * P, SG = 4 * S, taps = (size_t)dm->K * dm->K;
  const int groups = 7;   // (4 oc-blocks x 25 ic) x 7 = 700 CTAs: 5 per SM at most vs 4.73 on average (wave quantisation)
  Carver sv(saved);
  const bool ring = dm->ckpt_ring != 0;
  float* hs = sv.take<float>((size_t)(ring ? 2 : dm->T + 1) * S);
  float* cs = sv.take<float>((size_t)(ring ? 2 : dm->T + 1) * S);
  float* gates = sv.take<float>((size_t)(ring ? 2 : dm->T) * SG);
  float* u = sv.take<float>(S);
  float* xt = sv.take<float>(S);
  float* hstats = sv.take<float>(64);
  if (saved_bytes) *saved_bytes = (sv.off + 255) & ~(size_t)255;
  Carver cv(ws);
  float* w0T = cv.take<float>(49 * 32);
  float* wxT = cv.take<float>(3200 * taps);
  float* whT = cv.take<float>(3200 * taps);
  float* whD = cv.take<float>(3200 * taps);
  float* wxD = cv.take<float>(3200 * taps);
  float* bx = cv.take<float>(128);
  float* Xg = cv.take<float>(SG);
  float* G = cv.take<float>(SG);
  float* delta = cv.take<float>(SG);
  float* dsum = cv.take<float>(SG);
  float* dh = cv.take<float>(S);
  float* dc = cv.take<float>(S);
  float* slots_h = cv.take<float>((size_t)groups * 2500 * taps);
  float* slots_x = cv.take<float>((size_t)groups * 2500 * taps);
  float* slots_0 = cv.take<float>((size_t)groups * 25 * 49);
  float* bxg = cv.take<float>(128);
  float* b0g = cv.take<float>(32);
  float* hg = cv.take<float>(128);
  float* rbp_g = cv.take<float>(S);
  float* jp_rc = cv.take<float>(S);
  float* jp_eo = cv.take<float>(S);
  float* jp_dcx = cv.take<float>(S);
  float* rbp_state = cv.take<float>(64);
  double* rbp_part = cv.take<double>(2 * 148 * 2);
  uint8_t *whF_tc = nullptr, *whD_tc = nullptr, *wxF_tc = nullptr, *wxD_tc = nullptr;
  float *slots_h_tc = nullptr, *slots_x_tc = nullptr;
  uint8_t* dyc_tc = nullptr;
  if (dm->conv_mode != PTK_CONV_FP32) {
    slots_h_tc = cv.take<float>(wgradk_tc_slot_floats(dm->K));
    slots_x_tc = cv.take<float>(wgradk_tc_slot_floats(dm->K));
    dyc_tc = cv.take<uint8_t>(wgradk_tc_dyc_bytes(dm->B, dm->H, dm->W));
    wxF_tc = cv.take<uint8_t>(dm->K == 15 ? convk_tc_weight_bytes<15>(0) : convk_tc_weight_bytes<7>(0));
    wxD_tc = cv.take<uint8_t>(dm->K == 15 ? convk_tc_weight_bytes<15>(1) : convk_tc_weight_bytes<7>(1));
    whF_tc = cv.take<uint8_t>(dm->K == 15 ? convk_tc_weight_bytes<15>(0) : convk_tc_weight_bytes<7>(0));
    whD_tc = cv.take<uint8_t>(dm->K == 15 ? convk_tc_weight_bytes<15>(1) : convk_tc_weight_bytes<7>(1));
  }
  if (d) {
    d->whF_tc = whF_tc; d->whD_tc = whD_tc; d->conv_mode = dm->conv_mode; d->slots_h_tc = slots_h_tc;
    d->slots_x_tc = slots_x_tc; d->wxF_tc = wxF_tc; d->wxD_tc = wxD_tc; d->dyc_tc = dyc_tc;
    d->rbp_g = rbp_g; d->rbp_state = rbp_state; d->rbp_part = rbp_part;
    d->jp_rc = jp_rc; d->jp_eo = jp_eo; d->jp_dcx = jp_dcx;
    d->groups = groups;
    d->hs = hs; d->cs = cs; d->gates = gates; d->u = u; d->xt = xt; d->head_stats = hstats;
    d->w0T = w0T; d->wxT = wxT; d->whT = whT; d->whD = whD; d->wxD = wxD; d->bx = bx; d->Xg = Xg; d->G = G;
    d->delta = delta; d->delta_sum = dsum; d->dh = dh; d->dc = dc; d->slots_h = slots_h; d->slots_x = slots_x;
    d->slots_0 = slots_0; d->bx_grad = bxg; d->b0_grad = b0g; d->head_grads = hg;
  }
  return (cv.off + 255) & ~(size_t)255;
}
static int lstm_bind(LstmDev* d, const PtkLstmDims* dm, const float* x, const float* params, void* saved, void* ws,
                     cudaStream_t st) {
  memset(d, 0, sizeof(*d));
  d->B = dm->B; d->T = dm->T; d->H = dm->H; d->W = dm->W; d->KS = dm->K;
  d->ring = dm->ckpt_ring != 0;
  d->detached = dm->rbp != 0;
  d->x = x;
  int64_t o[20];
  lstm_offsets(dm->K, o);
  d->w0 = params + o[0]; d->b0 = params + o[1];
  for (int k = 0; k < 4; ++k) { d->wx[k] = params + o[2 + 3 * k]; d->wh[k] = params + o[4 + 3 * k]; }
  d->bn_w = params + o[14]; d->bn_b = params + o[15]; d->w6 = params + o[16]; d->b6 = params + o[17];
  d->off_w0 = o[0]; d->off_b0 = o[1]; d->off_wx = o[2]; d->off_bx = o[3]; d->off_wh = o[4]; d->off_head = o[14];
  lstm_carve(dm, saved, ws, d, nullptr);
  for (int k = 0; k < 4; ++k)   // gather the four Wx biases into one [100] vector
    PTK_CUDA_OK(cudaMemcpyAsync((float*)d->bx + 25 * k, params + o[3 + 3 * k], 25 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return PTK_OK;
}

int ptk_convlstm_param_offsets(const PtkLstmDims* d, int64_t* offsets) {
  if (!d || !offsets) { set_error("ptk_convlstm_param_offsets: NULL argument"); return PTK_EINVAL; }
  lstm_offsets(d->K, offsets);
  return PTK_OK;
}
size_t ptk_convlstm_saved_bytes(const PtkLstmDims* d) {
  if (lstm_check(d, "ptk_convlstm_saved_bytes")) return 0;
  size_t sb = 0;
  lstm_carve(d, nullptr, nullptr, nullptr, &sb);
  return sb;
}
size_t ptk_convlstm_workspace_bytes(const PtkLstmDims* d) {
  if (lstm_check(d, "ptk_convlstm_workspace_bytes")) return 0;
  return lstm_carve(d, nullptr, nullptr, nullptr, nullptr);
}
int ptk_convlstm_forward(const PtkLstmDims* dm, const float* x, const float* params, float* output, float* penalty,
                         void* saved, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = lstm_check(dm, "ptk_convlstm_forward");
  if (rc) return rc;
  if (!x || !params || !output || !saved || !workspace) { set_error("ptk_convlstm_forward: NULL argument"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_convlstm_workspace_bytes(dm)) { set_error("ptk_convlstm_forward: workspace too small"); return PTK_ENOMEM; }
  if (penalty && dm->T < 2) { set_error("ptk_convlstm_forward: the Jacobian penalty needs timesteps >= 2 (models/convlstm.py:140-144)"); return PTK_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  LstmDev d;
  if ((rc = lstm_bind(&d, dm, x, params, saved, workspace, st))) return rc;
  d.out = output; d.penalty = penalty;
  if ((rc = run_convlstm_forward(d, st))) return rc;
  if (penalty) rc = run_convlstm_penalty(d, st);
  return rc;
}
int ptk_convlstm_backward(const PtkLstmDims* dm, const float* x, const float* params, const void* saved,
                          const float* doutput, float* dparams, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = lstm_check(dm, "ptk_convlstm_backward");
  if (rc) return rc;
  if (!x || !params || !saved || !doutput || !dparams || !workspace) { set_error("ptk_convlstm_backward: NULL argument"); return PTK_EINVAL; }
  if (dm->ckpt_ring) { set_error("ptk_convlstm_backward: BPTT needs the checkpoints of every step; the forward ran with ckpt_ring = 1"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_convlstm_workspace_bytes(dm)) { set_error("ptk_convlstm_backward: workspace too small"); return PTK_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  LstmDev d;
  if ((rc = lstm_bind(&d, dm, x, params, const_cast<void*>(saved), workspace, st))) return rc;
  d.dout = doutput; d.dparams = dparams;
  return run_convlstm_backward(d, st);
}

int ptk_convlstm_backward_rbp(const PtkLstmDims* dm, const float* x, const float* params, const void* saved,
                              const float* doutput, int32_t num_iter, float* dparams, float* terms_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  int rc = lstm_check(dm, "ptk_convlstm_backward_rbp");
  if (rc) return rc;
  if (!x || !params || !saved || !doutput || !dparams || !workspace) { set_error("ptk_convlstm_backward_rbp: NULL argument"); return PTK_EINVAL; }
  if (num_iter < 0) { set_error("ptk_convlstm_backward_rbp: num_iter %d < 0", num_iter); return PTK_EINVAL; }
  if (workspace_bytes < ptk_convlstm_workspace_bytes(dm)) { set_error("ptk_convlstm_backward_rbp: workspace too small"); return PTK_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  LstmDev d;
  if ((rc = lstm_bind(&d, dm, x, params, const_cast<void*>(saved), workspace, st))) return rc;
  d.dout = doutput; d.dparams = dparams;
  if ((rc = run_convlstm_backward_rbp(d, num_iter, st))) return rc;
  if (terms_out) PTK_CUDA_OK(cudaMemcpyAsync(terms_out, d.rbp_state + 2, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return PTK_OK;
}

int ptk_convlstm_backward_penalty(const PtkLstmDims* dm, const float* x, const float* params, const void* saved,
                                  const float* doutput, const float* dpenalty, int32_t num_iter, float* dparams, float* terms_out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  int rc = lstm_check(dm, "ptk_convlstm_backward_penalty");
  if (rc) return rc;
  if (!x || !params || !saved || !doutput || !dpenalty || !dparams || !workspace) { set_error("ptk_convlstm_backward_penalty: NULL argument"); return PTK_EINVAL; }
  if (dm->T < 2) { set_error("ptk_convlstm_backward_penalty: the Jacobian penalty needs timesteps >= 2 (models/convlstm.py:140-144)"); return PTK_EINVAL; }
  if (!dm->rbp && dm->ckpt_ring) { set_error("ptk_convlstm_backward_penalty: BPTT needs the checkpoints of every step; the forward ran with ckpt_ring = 1"); return PTK_EINVAL; }
  if (dm->rbp && num_iter < 0) { set_error("ptk_convlstm_backward_penalty: num_iter %d < 0", num_iter); return PTK_EINVAL; }
  if (workspace_bytes < ptk_convlstm_workspace_bytes(dm)) { set_error("ptk_convlstm_backward_penalty: workspace too small"); return PTK_ENOMEM; }
  cudaStream_t st = (cudaStream_t)stream;
  LstmDev d;
  if ((rc = lstm_bind(&d, dm, x, params, const_cast<void*>(saved), workspace, st))) return rc;
  d.dout = doutput; d.dparams = dparams; d.dpen = dpenalty;
  if (!dm->rbp) return run_convlstm_backward(d, st);
  if ((rc = run_convlstm_backward_rbp(d, num_iter, st))) return rc;
  if (terms_out) PTK_CUDA_OK(cudaMemcpyAsync(terms_out, d.rbp_state + 2, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return PTK_OK;
}

int ptk_convlstm_input_grad(const PtkLstmDims* dm, const float* params, float* dx, void* workspace, size_t workspace_bytes,
                            void* stream) {
  int rc = lstm_check(dm, "ptk_convlstm_input_grad");
  if (rc) return rc;
  if (!params || !dx || !workspace) { set_error("ptk_convlstm_input_grad: NULL argument"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_convlstm_workspace_bytes(dm)) { set_error("ptk_convlstm_input_grad: workspace too small"); return PTK_ENOMEM; }
  LstmDev d;
  memset(&d, 0, sizeof(d));
  d.B = dm->B; d.T = dm->T; d.H = dm->H; d.W = dm->W; d.KS = dm->K;
  int64_t o[20];
  lstm_offsets(dm->K, o);
  d.w0 = params + o[0];
  lstm_carve(dm, nullptr, workspace, &d, nullptr);     // only the workspace view is used: `dh` still holds d(loss)/d(u)
  return run_convlstm_dx(d, dx, (cudaStream_t)stream);
}

// -----------------------------------------------------------------------------------------------------------------
// ConvGRU (models/kys.py)
static int gru_check(const PtkGruDims* d, const char* who) {
  if (!d) { set_error("%s: dims is NULL", who); return PTK_EINVAL; }
  if (d->C != GC || d->K != 7) {
    set_error("%s: only dimensions=64, kernel_size=7 are built (utils/engine.py:147-153 passes exactly these); got C=%d K=%d",
              who, d->C, d->K);
    return PTK_EUNSUPPORTED;
  }
  if (d->B < 1 || d->T < 1 || d->H < 16 || d->W < 32 || d->H % 16 || d->W % 32) {
    set_error("%s: unsupported shape B=%d T=%d H=%d W=%d (need B,T >= 1, H %% 16 == 0, W %% 32 == 0)", who, d->B, d->T, d->H, d->W);
    return PTK_EINVAL;
  }
  if (d->conv_mode != PTK_CONV_FP32 && d->conv_mode != PTK_CONV_BF16X3 && d->conv_mode != PTK_CONV_BF16) {
    set_error("%s: unknown conv_mode %d", who, d->conv_mode);
    return PTK_EINVAL;
  }
  return PTK_OK;
}
static size_t gru_carve(const PtkGruDims* dm, void* saved, void* ws, GruDev* d, size_t* saved_bytes) {
  const size_t P = (size_t)dm->H * dm->W, S = (size_t)dm->B * GC * P, taps = 49;
  const int groups = 7;
  Carver sv(saved);     // layout relied upon by pathtracker_models_b200/kys.py (testmode dumps): h, r, u, n
  float* hs = sv.take<float>((size_t)(dm->T + 1) * S);
  float* rs = sv.take<float>((size_t)dm->T * S);
  float* us = sv.take<float>((size_t)dm->T * S);
  float* ns = sv.take<float>((size_t)dm->T * S);
  if (saved_bytes) *saved_bytes = (sv.off + 255) & ~(size_t)255;
  Carver cv(ws);
  float* wruT = cv.take<float>(4 * 2 * GC * taps * 32);
  float* wnT = cv.take<float>(2 * 2 * GC * taps * 32);
  float* wruD = cv.take<float>(4 * 2 * GC * taps * 32);
  float* wnD = cv.take<float>(4 * GC * taps * 32);
  float* bru = cv.take<float>(2 * GC);
  float* Z = cv.take<float>(2 * S);
  float* Z2 = cv.take<float>(2 * S);
  float* RU = cv.take<float>(2 * S);
  float* dZ = cv.take<float>(2 * S);
  float* dRUsum = cv.take<float>(2 * S);
  float* Nb = cv.take<float>(S);
  float* dh = cv.take<float>(S);
  float* dhp = cv.take<float>(S);
  float* dXb = cv.take<float>(S);
  float* dNsum = cv.take<float>(S);
  float* slots_ru = cv.take<float>((size_t)groups * 2 * GC * 2 * GC * taps);
  float* slots_n = cv.take<float>((size_t)groups * GC * 2 * GC * taps);
  float* stem_acc = cv.take<float>(GRU_STEM_SPLIT * 4 * GC);
  float* bias_grad = cv.take<float>(3 * GC);
  float* head_part = cv.take<float>((size_t)dm->B * GRU_HEAD_N);
  float* dgoal = cv.take<float>((size_t)dm->B * P);
  uint8_t *wtcF = nullptr, *wtcD = nullptr;
  float* wacc_tc = nullptr;
  const size_t wacc_pair = wgrad_tc_acc_floats(dm->B, dm->H, dm->W);
  uint8_t *wk_ruF = nullptr, *wk_ruD = nullptr, *wk_nF = nullptr, *wk_nD = nullptr;
  float* slots_k = nullptr;
  uint8_t* dyc_k = nullptr;
  if (dm->conv_mode == PTK_CONV_BF16) {      // single-pass mode: 32x32-channel blocks on the InT kernels
    wtcF = cv.take<uint8_t>((size_t)24 * 4 * C * C * TAPS);      // bf16 hi + lo per pair = 4 bytes per weight
    wtcD = cv.take<uint8_t>((size_t)24 * 4 * C * C * TAPS);
    wacc_tc = cv.take<float>(24 * wacc_pair);
  } else if (dm->conv_mode == PTK_CONV_BF16X3) {
    wk_ruF = cv.take<uint8_t>(convk_tc_stream_bytes(7, 128, 4));
    wk_ruD = cv.take<uint8_t>(convk_tc_stream_bytes(7, 128, 4));
    wk_nF = cv.take<uint8_t>(convk_tc_stream_bytes(7, 64, 4));
    wk_nD = cv.take<uint8_t>(convk_tc_stream_bytes(7, 128, 2));
    slots_k = cv.take<float>(8 * wgradk_tc_slot_floats(7));
    dyc_k = cv.take<uint8_t>(wgradk_tc_dyc_bytes(dm->B, dm->H, dm->W));
  }
  if (d) {
    d->wk_ruF = wk_ruF; d->wk_ruD = wk_ruD; d->wk_nF = wk_nF; d->wk_nD = wk_nD; d->slots_k = slots_k; d->dyc_k = dyc_k;
    d->groups = groups; d->conv_mode = dm->conv_mode; d->head_part = head_part; d->dgoal = dgoal;
    d->wtcF = wtcF; d->wtcD = wtcD; d->wacc_tc = wacc_tc; d->wacc_pair = wacc_pair;
    d->hs = hs; d->rs = rs; d->us = us; d->ns = ns;
    d->wruT = wruT; d->wnT = wnT; d->wruD = wruD; d->wnD = wnD; d->bru = bru;
    d->Z = Z; d->Z2 = Z2; d->RU = RU; d->dZ = dZ; d->dRUsum = dRUsum; d->Nb = Nb; d->dh = dh; d->dhp = dhp; d->dXb = dXb;
    d->dNsum = dNsum; d->slots_ru = slots_ru; d->slots_n = slots_n; d->stem_acc = stem_acc; d->bias_grad = bias_grad;
  }
  return (cv.off + 255) & ~(size_t)255;
}
static void gru_offsets(int64_t* o) {   // preproc.{w,b}, conv_reset.{w,b}, conv_update.{w,b}, conv_state_new.{w,b},
                                        // readout_conv.{w,b}, target_conv.{w,b}, readout_dense.{w,b}, total
  const int64_t nw = (int64_t)GC * 2 * GC * 49;
  const int64_t sizes[14] = {GC * 3, GC, nw, GC, nw, GC, nw, GC, GC, 1, 50, 1, 1, 1};
  int64_t a = 0;
  for (int i = 0; i < 14; ++i) { o[i] = a; a += sizes[i]; }
  o[14] = a;
}
static int gru_bind(GruDev* d, const PtkGruDims* dm, const float* x, const int64_t* xs, const float* params, void* saved,
                    void* ws) {
  memset(d, 0, sizeof(*d));
  d->B = dm->B; d->T = dm->T; d->H = dm->H; d->W = dm->W;
  d->fv.x = x; d->fv.sb = xs[0]; d->fv.sc = xs[1]; d->fv.st = xs[2]; d->fv.sy = xs[3]; d->fv.sx = xs[4];
  int64_t o[15];
  gru_offsets(o);
  d->Wp = params + o[0]; d->bp = params + o[1]; d->wr = params + o[2]; d->br = params + o[3];
  d->wu = params + o[4]; d->bu = params + o[5]; d->wn = params + o[6]; d->bnw = params + o[7];
  d->off_wp = o[0]; d->off_bp = o[1]; d->off_wr = o[2]; d->off_br = o[3]; d->off_wu = o[4]; d->off_bu = o[5];
  d->off_wn = o[6]; d->off_bn = o[7];
  d->hWr = params + o[8]; d->hbr = params + o[9]; d->hWt = params + o[10]; d->hbt = params + o[11];
  d->hWd = params + o[12]; d->hbd = params + o[13]; d->off_head = o[8];
  gru_carve(dm, saved, ws, d, nullptr);
  return PTK_OK;
}

int ptk_gru_param_offsets(const PtkGruDims* d, int64_t* offsets) {
  if (!d || !offsets) { set_error("ptk_gru_param_offsets: NULL argument"); return PTK_EINVAL; }
  if (int rc = gru_check(d, "ptk_gru_param_offsets")) return rc;
  gru_offsets(offsets);
  return PTK_OK;
}
size_t ptk_gru_saved_bytes(const PtkGruDims* d) {
  if (gru_check(d, "ptk_gru_saved_bytes")) return 0;
  size_t sb = 0;
  gru_carve(d, nullptr, nullptr, nullptr, &sb);
  return sb;
}
size_t ptk_gru_workspace_bytes(const PtkGruDims* d) {
  if (gru_check(d, "ptk_gru_workspace_bytes")) return 0;
  return gru_carve(d, nullptr, nullptr, nullptr, nullptr);
}
int ptk_gru_forward(const PtkGruDims* dm, const float* x, const int64_t x_strides[5], const float* params, float* logits,
                    void* saved, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = gru_check(dm, "ptk_gru_forward");
  if (rc) return rc;
  if (!x || !x_strides || !params || !logits || !saved || !workspace) { set_error("ptk_gru_forward: NULL argument"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_gru_workspace_bytes(dm)) { set_error("ptk_gru_forward: workspace too small"); return PTK_ENOMEM; }
  GruDev d;
  if ((rc = gru_bind(&d, dm, x, x_strides, params, saved, workspace))) return rc;
  d.logits = logits;
  return run_gru_forward(d, (cudaStream_t)stream);
}
int ptk_gru_backward(const PtkGruDims* dm, const float* x, const int64_t x_strides[5], const float* params, const void* saved,
                     const float* dlogits, float* dparams, float* dx, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = gru_check(dm, "ptk_gru_backward");
  if (rc) return rc;
  if (!x || !x_strides || !params || !saved || !dlogits || !dparams || !workspace) { set_error("ptk_gru_backward: NULL argument"); return PTK_EINVAL; }
  if (workspace_bytes < ptk_gru_workspace_bytes(dm)) { set_error("ptk_gru_backward: workspace too small"); return PTK_ENOMEM; }
  GruDev d;
  if ((rc = gru_bind(&d, dm, x, x_strides, params, const_cast<void*>(saved), workspace))) return rc;
  d.dlogits = dlogits; d.dparams = dparams; d.dx = dx;
  return run_gru_backward(d, (cudaStream_t)stream);
}

}  // extern "C"
