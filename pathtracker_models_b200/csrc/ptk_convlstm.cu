// ConvLSTM baseline (models/convlstm.py) on the GPU: strict-fp32 CUDA-core kernels.
//
//   forward  (ConvLSTM.forward :116-148, bptt branch; ConvLSTMCell.forward :84-90)
//       u = conv0(x) (7x7 Gabor, 1 -> 25);  xt = u^2
//       Xg = [Wxi|Wxf|Wxc|Wxo] * xt + b          (hoisted: the reference recomputes it every step)
//       for t in 1..T:  G = Xg + [Whi|Whf|Whc|Who] * h_{t-1};  i,f,o = sigmoid, g = tanh
//                       c_t = f c_{t-1} + i g;  h_t = o tanh(c_t)
//       out = conv6(BatchNorm_batchstats(h_T))    (:146-147)
//   backward (autograd of the above, SURVEY section 8 "ConvLSTM step"):
//       per step  delta = gate pre-activation gradients;  dh_{t-1} = delta *^T Wh;  dWh += h_{t-1} (star) delta
//       once      dWx = xt (star) sum_t delta,  db = sum delta,  d xt = (sum_t delta) *^T Wx,  d conv0
//   penalty  (:152-163): the two vector-Jacobian products of the last step.
//
// Gate channel stacking everywhere: channel = gate * 25 + c with gate order (i, f, c/g, o).
// Convolutions: one generic direct kernel (16x32 pixel tile x 32 output channels per CTA, 8 px x 8 oc
// register tiles, input-channel chunks staged by cp.async) templated on the kernel size; the weight
// gradient keeps all taps of one (oc, ic, ky) triple in registers across a persistent CTA's tiles.
#include "ptk_common.cuh"
#include "ptk_convg.cuh"
#include "ptk_convlstm.h"
#include "ptk_kernels.h"
#include "ptk_profile.h"

namespace ptk {

constexpr int LC = 25;            // hidden channels of the ConvLSTM (models/convlstm.py:108)
constexpr int LG = 4 * LC;        // stacked gate channels
// ---- point-wise LSTM cell -------------------------------------------------------------------------------------
__global__ void lstm_fwd_kernel(const float* __restrict__ G, const float* __restrict__ c_prev, float* gates,
                                float* c_out, float* h_out, int B, int P) {
  const size_t n = (size_t)B * LC * P;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % P), c = (int)((e / P) % LC), b = (int)(e / ((size_t)P * LC));
    const size_t g0 = ((size_t)b * LG + c) * P + pix, gs = (size_t)LC * P;
    const float i = sigmoid_precise(G[g0]), f = sigmoid_precise(G[g0 + gs]), g = tanhf(G[g0 + 2 * gs]),
                o = sigmoid_precise(G[g0 + 3 * gs]);
    const float c2 = f * c_prev[e] + i * g;
    gates[g0] = i; gates[g0 + gs] = f; gates[g0 + 2 * gs] = g; gates[g0 + 3 * gs] = o;
    c_out[e] = c2;
    h_out[e] = o * tanhf(c2);
  }
}
// delta of the gate pre-activations from (dh, dc); dh/dc may be constants (penalty seeds) when null.
__global__ void lstm_bwd_kernel(const float* __restrict__ dh, float dh_const, const float* __restrict__ dc, float dc_const,
                                const float* __restrict__ gates, const float* __restrict__ c_prev,
                                const float* __restrict__ c_cur, float* delta, float* dc_prev, float* delta_sum,
                                int first, int B, int P, const float* skip = nullptr) {
  if (skip != nullptr && skip[1] != 0.f) return;
  const size_t n = (size_t)B * LC * P;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % P), c = (int)((e / P) % LC), b = (int)(e / ((size_t)P * LC));
    const size_t g0 = ((size_t)b * LG + c) * P + pix, gs = (size_t)LC * P;
    const float i = gates[g0], f = gates[g0 + gs], g = gates[g0 + 2 * gs], o = gates[g0 + 3 * gs];
    const float tc = tanhf(c_cur[e]);
    const float dhv = dh ? dh[e] : dh_const;
    const float dcv = (dc ? dc[e] : dc_const) + dhv * o * (1.0f - tc * tc);
    const float d_i = dcv * g * i * (1.0f - i), d_f = dcv * c_prev[e] * f * (1.0f - f), d_g = dcv * i * (1.0f - g * g),
                d_o = dhv * tc * o * (1.0f - o);
    delta[g0] = d_i; delta[g0 + gs] = d_f; delta[g0 + 2 * gs] = d_g; delta[g0 + 3 * gs] = d_o;
    if (dc_prev) dc_prev[e] = dcv * f;
    if (delta_sum) {
      if (first) { delta_sum[g0] = d_i; delta_sum[g0 + gs] = d_f; delta_sum[g0 + 2 * gs] = d_g; delta_sum[g0 + 3 * gs] = d_o; }
      else { delta_sum[g0] += d_i; delta_sum[g0 + gs] += d_f; delta_sum[g0 + 2 * gs] += d_g; delta_sum[g0 + 3 * gs] += d_o; }
    }
  }
}
__global__ void square_kernel(const float* __restrict__ u, float* xt, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) xt[e] = u[e] * u[e];
}
__global__ void dsquare_kernel(const float* __restrict__ u, float* g, size_t n) {   // g <- 2 u g
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) g[e] = 2.0f * u[e] * g[e];
}
// jc = f_T + v * o_{T-1} * (1 - tanh^2 c_{T-1});  pen = (jh - .9)+^2 + (jc - .9)+^2   (convlstm.py:152-163)
__global__ void penalty_kernel(const float* __restrict__ jh, const float* __restrict__ v, const float* __restrict__ gates_T,
                               const float* __restrict__ gates_Tm1, const float* __restrict__ c_Tm1, float* pen, int B, int P) {
  const size_t n = (size_t)B * LC * P;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % P), c = (int)((e / P) % LC), b = (int)(e / ((size_t)P * LC));
    const size_t g0 = ((size_t)b * LG + c) * P + pix, gs = (size_t)LC * P;
    const float fT = gates_T[g0 + gs];
    float jc = fT;
    if (gates_Tm1) {   // T >= 2: path c_{T-1} -> h_{T-1} -> gates_T -> c_T
      const float tc = tanhf(c_Tm1[e]);
      jc += v[e] * gates_Tm1[g0 + 3 * gs] * (1.0f - tc * tc);
    }
    const float a = fmaxf(jh[e] - 0.9f, 0.f), bq = fmaxf(jc - 0.9f, 0.f);
    pen[e] = a * a + bq * bq;
  }
}

// ---- the Jacobian penalty INSIDE the graph (jacobian_penalty=True: create_graph=True at models/convlstm.py:157-160) ---------
// total = loss + <dpen, pen>,  pen = (jh - .9)+^2 + (jc - .9)+^2,  jh = delta^h *^T Wh,  jc = f_T + v m,  v = delta^c *^T Wh,
// m = o_{T-1} (1 - tanh^2 c_{T-1})  (m = 0 when h_{T-1}, c_{T-1} are detached leaves: grad_method='rbp').
// delta^h / delta^c are the gate deltas of step T for the seeds (dh, dc) = (1, 0) / (0, 1).  Differentiating once more:
//   r_h = d total / d jh,  r_c = d total / d jc,  r_v = r_c m;   rho^h = r_h * Wh,  rho^c = r_v * Wh  (forward convolutions);
//   d Wh += r_h (star) delta^h + r_v (star) delta^c  (the weight-gradient operator with the roles of input / delta as named);
//   eps = d total / d (gate pre-activations of step T) through delta^h, delta^c and f_T  (jp_eps_kernel), an extra d c_{T-1},
//   and through m an extra output-gate delta at step T-1 and one more d c_{T-1} term (jp_seed_kernel).
// These are added to the ordinary deltas of steps T and T-1, so the rest is the unchanged reverse loop.
// Checked against double back-propagation in fp64 (oracle/convlstm_oracle.py:jp_total_grads) and the reference's own gradients.
__global__ void jp_seed_kernel(float* jh_rh, float* v_rv, const float* __restrict__ gates_T, const float* __restrict__ gates_Tm1,
                               const float* __restrict__ c_Tm1, const float* __restrict__ dpen, float* rc_out, float* eo, float* dcx,
                               int B, int P) {
  const size_t n = (size_t)B * LC * P;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % P), c = (int)((e / P) % LC), b = (int)(e / ((size_t)P * LC));
    const size_t g0 = ((size_t)b * LG + c) * P + pix, gs = (size_t)LC * P;
    float m = 0.f, o1 = 0.f, tc1 = 0.f, v = 0.f;
    if (gates_Tm1) { o1 = gates_Tm1[g0 + 3 * gs]; tc1 = tanhf(c_Tm1[e]); m = o1 * (1.0f - tc1 * tc1); v = v_rv[e]; }
    const float jc = gates_T[g0 + gs] + v * m;
    const float w2 = 2.0f * dpen[e];
    const float rh = w2 * fmaxf(jh_rh[e] - 0.9f, 0.f), rc = w2 * fmaxf(jc - 0.9f, 0.f), kk = rc * v;
    jh_rh[e] = rh;
    if (v_rv) v_rv[e] = rc * m;
    rc_out[e] = rc;
    eo[e] = kk * (1.0f - tc1 * tc1) * o1 * (1.0f - o1);
    dcx[e] = kk * o1 * (-2.0f * tc1 * (1.0f - tc1 * tc1));
  }
}
// rho^h (in place -> eps), rho^c (null when detached), r_c;  dcx += the step-T part
__global__ void jp_eps_kernel(float* rhoH_eps, const float* __restrict__ rhoC, const float* __restrict__ rcv,
                              const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_cur,
                              float* dcx, int B, int P) {
  const size_t n = (size_t)B * LC * P;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % P), c = (int)((e / P) % LC), b = (int)(e / ((size_t)P * LC));
    const size_t g0 = ((size_t)b * LG + c) * P + pix, gs = (size_t)LC * P;
    const float i = gates[g0], f = gates[g0 + gs], g = gates[g0 + 2 * gs], o = gates[g0 + 3 * gs];
    const float cp = c_prev[e], s = tanhf(c_cur[e]), D = 1.0f - s * s, kap = o * D;
    const float ip = i * (1.0f - i), fp = f * (1.0f - f), op = o * (1.0f - o), gp = 1.0f - g * g;
    const float ipp = ip * (1.0f - 2.0f * i), fpp = fp * (1.0f - 2.0f * f), opp = op * (1.0f - 2.0f * o), gpp = -2.0f * g * gp;
    const float rHi = rhoH_eps[g0], rHf = rhoH_eps[g0 + gs], rHc = rhoH_eps[g0 + 2 * gs], rHo = rhoH_eps[g0 + 3 * gs];
    float rCi = 0.f, rCf = 0.f, rCc = 0.f;
    if (rhoC) { rCi = rhoC[g0]; rCf = rhoC[g0 + gs]; rCc = rhoC[g0 + 2 * gs]; }
    const float Q = rHi * g * ip + rHf * cp * fp + rHc * i * gp;
    const float lam = rHo * op * D + Q * o * (-2.0f * s * D);          // d total / d c_T through tanh(c_T)
    const float ai = rHi * kap + rCi, af = rHf * kap + rCf, ac = rHc * kap + rCc;   // d total / d (delta-without-kappa) per gate
    rhoH_eps[g0] = lam * g * ip + ai * g * ipp + ac * ip * gp;
    rhoH_eps[g0 + gs] = lam * cp * fp + af * cp * fpp + rcv[e] * fp;
    rhoH_eps[g0 + 2 * gs] = lam * i * gp + ai * ip * gp + ac * i * gpp;
    rhoH_eps[g0 + 3 * gs] = rHo * s * opp + Q * D * op;
    dcx[e] += lam * f + af * fp;
  }
}
// delta[:, gate_lo : gate_lo + n_gates] += extra (same for delta_sum);  dc += dcx
__global__ void lstm_add_kernel(float* delta, float* delta_sum, const float* __restrict__ extra, int gate_lo, int n_gates,
                                float* dc, const float* __restrict__ dcx, int B, int P) {
  const size_t n = (size_t)B * LC * P;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % P), c = (int)((e / P) % LC), b = (int)(e / ((size_t)P * LC));
    const size_t g0 = ((size_t)b * LG + c) * P + pix, gs = (size_t)LC * P;
    const size_t x0 = ((size_t)b * n_gates * LC + c) * P + pix;
    for (int k = 0; k < n_gates; ++k) {
      const float x = extra[x0 + k * gs];
      delta[g0 + (gate_lo + k) * gs] += x;
      delta_sum[g0 + (gate_lo + k) * gs] += x;
    }
    if (dc) dc[e] += dcx[e];
  }
}

// ---- head: BatchNorm (batch statistics) + 1x1 conv 25 -> 2, forward and backward; one CTA per channel ----------
__global__ void __launch_bounds__(512) head_stats_kernel(const float* __restrict__ h, float* stats, int B, int P) {
  __shared__ double sh[16];
  const int c = blockIdx.x;
  const size_t N = (size_t)B * P;
  double s = 0;
  for (size_t i = threadIdx.x; i < N; i += blockDim.x) s += h[((i / P) * LC + c) * P + i % P];
  const double mean = block_sum_d(s, sh) / (double)N;
  double q = 0;
  for (size_t i = threadIdx.x; i < N; i += blockDim.x) {
    const double dlt = h[((i / P) * LC + c) * P + i % P] - mean;
    q += dlt * dlt;
  }
  const double var = block_sum_d(q, sh) / (double)N;
  if (threadIdx.x == 0) { stats[c] = (float)mean; stats[LC + c] = (float)(1.0 / sqrt(var + (double)BN_EPS)); }
}
__global__ void head_fwd_kernel(const float* __restrict__ h, const float* __restrict__ stats, const float* __restrict__ bn_w,
                                const float* __restrict__ bn_b, const float* __restrict__ w6, const float* __restrict__ b6,
                                float* out, int B, int P) {
  const size_t n = (size_t)B * P;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(e % P), b = (int)(e / P);
    float o0 = __ldg(b6), o1 = __ldg(b6 + 1);
    for (int c = 0; c < LC; ++c) {
      const float hn = (h[((size_t)b * LC + c) * P + pix] - stats[c]) * stats[LC + c] * __ldg(bn_w + c) + __ldg(bn_b + c);
      o0 = fmaf(__ldg(w6 + c), hn, o0);
      o1 = fmaf(__ldg(w6 + LC + c), hn, o1);
    }
    out[((size_t)b * 2) * P + pix] = o0;
    out[((size_t)b * 2 + 1) * P + pix] = o1;
  }
}
// per channel c: sums for dbn_w, dbn_b, dw6[:, c]; then dh for that channel.  grads: [bn_w 25][bn_b 25][w6 50][b6 2]
__global__ void __launch_bounds__(512) head_bwd_kernel(const float* __restrict__ h, const float* __restrict__ stats,
                                                       const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                                       const float* __restrict__ w6, const float* __restrict__ dout, float* dh,
                                                       float* grads, int B, int P) {
  __shared__ double sh[16];
  const int c = blockIdx.x;
  const size_t N = (size_t)B * P;
  const float mean = stats[c], rstd = stats[LC + c], gw = __ldg(bn_w + c), gb = __ldg(bn_b + c);
  const float w0 = __ldg(w6 + c), w1 = __ldg(w6 + LC + c);
  double s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
  for (size_t i = threadIdx.x; i < N; i += blockDim.x) {
    const size_t b = i / P, pix = i % P;
    const float xh = (h[(b * LC + c) * P + pix] - mean) * rstd;
    const float d0 = dout[(b * 2) * P + pix], d1 = dout[(b * 2 + 1) * P + pix];
    const float dhn = w0 * d0 + w1 * d1;
    const float hn = xh * gw + gb;
    s1 += dhn; s2 += (double)dhn * xh; s3 += (double)d0 * hn; s4 += (double)d1 * hn; s5 += d0; s6 += d1;
  }
  s1 = block_sum_d(s1, sh); s2 = block_sum_d(s2, sh); s3 = block_sum_d(s3, sh); s4 = block_sum_d(s4, sh);
  s5 = block_sum_d(s5, sh); s6 = block_sum_d(s6, sh);
  if (threadIdx.x == 0) {
    grads[c] = (float)s2; grads[LC + c] = (float)s1; grads[2 * LC + c] = (float)s3; grads[3 * LC + c] = (float)s4;
    if (c == 0) { grads[4 * LC] = (float)s5; grads[4 * LC + 1] = (float)s6; }
  }
  const float m1 = (float)(s1 / (double)N), m2 = (float)(s2 / (double)N);
  for (size_t i = threadIdx.x; i < N; i += blockDim.x) {
    const size_t b = i / P, pix = i % P, idx = (b * LC + c) * P + pix;
    const float xh = (h[idx] - mean) * rstd;
    const float dhn = w0 * dout[(b * 2) * P + pix] + w1 * dout[(b * 2 + 1) * P + pix];
    dh[idx] = gw * rstd * (dhn - m1 - xh * m2);   // BatchNorm backward with batch statistics
  }
}
// the wide convolutions (25 -> 100 stacked gates, its transpose, its weight gradient) of the recurrent path (Wh: every
// step of the time loops) and of the hoisted input path (Wx: once per sequence).
// conv_mode bf16x3: tcgen05 implicit GEMMs (ptk_convk_tc.cu, ptk_wgradk_tc.cu) with the 3-pass split; bf16: the same kernels issuing
// the leading product only (a third of the MMAs, not fp32-parity); fp32: the generic direct kernels.
template <int KS>
static int wide_conv(const LstmDev& d, const ConvGArgs& g, bool transposed, bool xpath, cudaStream_t st) {
  if (d.conv_mode == PTK_CONV_FP32) return launch_convg<KS>(g, st);
  const void* w = xpath ? (transposed ? d.wxD_tc : d.wxF_tc) : (transposed ? d.whD_tc : d.whF_tc);
  ConvKArgs k{g.in, w, g.bias, g.addend, g.out, g.B, g.H, g.W, g.Cin, g.Cout, g.accumulate, g.skip};
  k.single = (d.conv_mode == PTK_CONV_BF16);
  ProfileScope ps(transposed ? PTK_PROF_CONV_DGRAD : PTK_PROF_CONV_FWD, st);
  return launch_convk_tc<KS>(k, transposed ? 1 : 0, st);
}
template <int KS>
static int recurrent_conv(const LstmDev& d, const ConvGArgs& g, bool transposed, cudaStream_t st) {
  return wide_conv<KS>(d, g, transposed, false, st);
}
template <int KS>
static int wide_wgrad(const LstmDev& d, const WgradGArgs& g, bool xpath, cudaStream_t st) {
  if (d.conv_mode == PTK_CONV_FP32) return launch_wgradg<KS>(g, st);
  // the KS kernel-row CTAs all read the same delta tiles: split them into bf16 hi/lo once, fetch by bulk copy
  int rc = launch_wgradk_prep_dy(g.dy, g.B, g.H, g.W, g.Cout, d.dyc_tc, nullptr, st);
  if (rc) return rc;
  WgradKArgs k{g.in, g.dy, xpath ? d.slots_x_tc : d.slots_h_tc, g.B, g.H, g.W, g.Cin, g.Cout, g.accumulate, nullptr, 0, d.dyc_tc};
  k.single = (d.conv_mode == PTK_CONV_BF16);
  ProfileScope ps(PTK_PROF_WGRAD, st);
  return launch_wgradk_tc<KS>(k, st);
}
template <int KS>
static int recurrent_wgrad(const LstmDev& d, const WgradGArgs& g, cudaStream_t st) { return wide_wgrad<KS>(d, g, false, st); }
template <int KS>
static int recurrent_prep(const LstmDev& d, bool transposed, cudaStream_t st) {   // both banks
  if (d.conv_mode == PTK_CONV_FP32) return PTK_OK;
  int rc = launch_prep_weights_k<KS>(d.wh, transposed ? 1 : 0, transposed ? d.whD_tc : d.whF_tc, st);
  if (rc) return rc;
  return launch_prep_weights_k<KS>(d.wx, transposed ? 1 : 0, transposed ? d.wxD_tc : d.wxF_tc, st);
}

template <int KS>
static int convlstm_forward_t(const LstmDev& d, cudaStream_t st) {
  const int P = d.H * d.W, taps = KS * KS;
  const size_t S = (size_t)d.B * LC * P, SG = (size_t)d.B * LG * P;
  int rc;
  // weights: conv0 [1 block][1][49][32]; Wx / Wh stacked (i,f,c,o): [4 blocks][25][taps][32]
  const float* w0[1] = {d.w0};
  if ((rc = launch_prepg(w0, 1, LC, 1, 49, 0, d.w0T, (size_t)49 * 32, st))) return rc;
  if ((rc = launch_prepg(d.wx, 4, LC, LC, taps, 0, d.wxT, (size_t)4 * LC * taps * 32, st))) return rc;
  if ((rc = launch_prepg(d.wh, 4, LC, LC, taps, 0, d.whT, (size_t)4 * LC * taps * 32, st))) return rc;
  if ((rc = recurrent_prep<KS>(d, false, st))) return rc;
  ConvGArgs c0{d.x, d.w0T, d.b0, nullptr, d.u, d.B, d.H, d.W, 1, LC, 0};
  if ((rc = launch_convg<7>(c0, st))) return rc;
  square_kernel<<<148 * 4, 256, 0, st>>>(d.u, d.xt, S);
  ConvGArgs cx{d.xt, d.wxT, d.bx, nullptr, d.Xg, d.B, d.H, d.W, LC, LG, 0};
  // the hoisted input term stays in strict fp32: xt = u^2 is all-positive and large, so the 16 mantissa bits of the
  // hi/lo split are not enough against the cancellation in Wx * xt (measured: output error 1.5e-4 of max vs 6e-7)
  if ((rc = launch_convg<KS>(cx, st))) return rc;
  PTK_CUDA_OK(cudaMemsetAsync(d.h_at(0), 0, S * sizeof(float), st));
  PTK_CUDA_OK(cudaMemsetAsync(d.c_at(0), 0, S * sizeof(float), st));
  for (int t = 1; t <= d.T; ++t) {
    ConvGArgs ch{d.h_at(t - 1), d.whT, nullptr, d.Xg, d.G, d.B, d.H, d.W, LC, LG, 0};
    if ((rc = recurrent_conv<KS>(d, ch, false, st))) return rc;
    ProfileScope ps(PTK_PROF_POINT_FWD, st);
    lstm_fwd_kernel<<<148 * 4, 256, 0, st>>>(d.G, d.c_at(t - 1), d.gates_at(t), d.c_at(t), d.h_at(t), d.B, P);
    PTK_LAUNCH_OK("lstm_fwd_kernel");
  }
  head_stats_kernel<<<LC, 512, 0, st>>>(d.h_at(d.T), d.head_stats, d.B, P);
  head_fwd_kernel<<<148 * 2, 256, 0, st>>>(d.h_at(d.T), d.head_stats, d.bn_w, d.bn_b, d.w6, d.b6, d.out, d.B, P);
  PTK_LAUNCH_OK("head_fwd_kernel");
  (void)SG;
  return PTK_OK;
}

template <int KS>
static int convlstm_penalty_t(const LstmDev& d, cudaStream_t st) {
  // needs the transposed Wh bank; scratch: delta (SG), jh (S), v (S)
  const int P = d.H * d.W, taps = KS * KS;
  int rc;
  if ((rc = launch_prepg(d.wh, 4, LC, LC, taps, 1, d.whD, (size_t)LG * taps * 32, st))) return rc;
  if ((rc = recurrent_prep<KS>(d, true, st))) return rc;
  const int T = d.T;
  // seed dh = 1, dc = 0  -> jh
  lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(nullptr, 1.0f, nullptr, 0.0f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                           nullptr, nullptr, 0, d.B, P);
  ConvGArgs c1{d.delta, d.whD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0};
  if ((rc = recurrent_conv<KS>(d, c1, true, st))) return rc;
  // seed dh = 0, dc = 1  -> v = d c_T / d h_{T-1}
  lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(nullptr, 0.0f, nullptr, 1.0f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                           nullptr, nullptr, 0, d.B, P);
  ConvGArgs c2{d.delta, d.whD, nullptr, nullptr, d.dc, d.B, d.H, d.W, LG, LC, 0};
  if ((rc = recurrent_conv<KS>(d, c2, true, st))) return rc;
  penalty_kernel<<<148 * 4, 256, 0, st>>>(d.dh, d.dc, d.gates_at(T), (T >= 2 && !d.detached) ? d.gates_at(T - 1) : nullptr,
                                          d.c_at(T - 1), d.penalty, d.B, P);
  PTK_LAUNCH_OK("penalty_kernel");
  return PTK_OK;
}

// second-order seeds of the Jacobian penalty (see jp_seed_kernel): leaves eps (4 gates) in d.G, the extra d c_{T-1} in
// d.jp_dcx, the step T-1 output-gate term in d.jp_eo and the direct d Wh terms in the weight-gradient slots (overwritten:
// every later recurrent weight gradient must accumulate).  Uses d.dh, d.dc, d.delta, d.Xg as scratch: call it first.
template <int KS>
static int convlstm_penalty_seed(const LstmDev& d, cudaStream_t st) {
  const int P = d.H * d.W, taps = KS * KS, T = d.T;
  const bool tail = !d.detached && T >= 2;      // the path c_{T-1} -> h_{T-1} -> gates_T -> c_T exists in the graph
  int rc;
  if ((rc = launch_prepg(d.wh, 4, LC, LC, taps, 0, d.whT, (size_t)4 * LC * taps * 32, st))) return rc;
  if ((rc = recurrent_prep<KS>(d, false, st))) return rc;
  lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(nullptr, 1.0f, nullptr, 0.0f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                           nullptr, nullptr, 0, d.B, P);
  ConvGArgs c1{d.delta, d.whD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0};
  if ((rc = recurrent_conv<KS>(d, c1, true, st))) return rc;                       // jh
  if (tail) {
    lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(nullptr, 0.0f, nullptr, 1.0f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.Xg,
                                             nullptr, nullptr, 0, d.B, P);
    ConvGArgs c2{d.Xg, d.whD, nullptr, nullptr, d.dc, d.B, d.H, d.W, LG, LC, 0};
    if ((rc = recurrent_conv<KS>(d, c2, true, st))) return rc;                     // v
  }
  jp_seed_kernel<<<148 * 4, 256, 0, st>>>(d.dh, tail ? d.dc : nullptr, d.gates_at(T), tail ? d.gates_at(T - 1) : nullptr,
                                          d.c_at(T - 1), d.dpen, d.jp_rc, d.jp_eo, d.jp_dcx, d.B, P);
  PTK_LAUNCH_OK("jp_seed_kernel");
  WgradGArgs w1{d.dh, d.delta, d.slots_h, d.B, d.H, d.W, LC, LG, d.groups, 0};
  if ((rc = recurrent_wgrad<KS>(d, w1, st))) return rc;                            // r_h (star) delta^h
  ConvGArgs f1{d.dh, d.whT, nullptr, nullptr, d.G, d.B, d.H, d.W, LC, LG, 0};
  if ((rc = recurrent_conv<KS>(d, f1, false, st))) return rc;                      // rho^h
  if (tail) {
    WgradGArgs w2{d.dc, d.Xg, d.slots_h, d.B, d.H, d.W, LC, LG, d.groups, 1};
    if ((rc = recurrent_wgrad<KS>(d, w2, st))) return rc;                          // r_v (star) delta^c
    ConvGArgs f2{d.dc, d.whT, nullptr, nullptr, d.delta, d.B, d.H, d.W, LC, LG, 0};
    if ((rc = recurrent_conv<KS>(d, f2, false, st))) return rc;                    // rho^c
  }
  jp_eps_kernel<<<148 * 4, 256, 0, st>>>(d.G, tail ? d.delta : nullptr, d.jp_rc, d.gates_at(T), d.c_at(T - 1), d.c_at(T),
                                         d.jp_dcx, d.B, P);
  PTK_LAUNCH_OK("jp_eps_kernel");
  return PTK_OK;
}

template <int KS>
static int convlstm_grad_tail(const LstmDev& d, cudaStream_t st);

template <int KS>
static int convlstm_backward_t(const LstmDev& d, cudaStream_t st) {
  const int P = d.H * d.W, taps = KS * KS;
  const size_t S = (size_t)d.B * LC * P;
  int rc;
  if ((rc = launch_prepg(d.wh, 4, LC, LC, taps, 1, d.whD, (size_t)LG * taps * 32, st))) return rc;
  if ((rc = recurrent_prep<KS>(d, true, st))) return rc;
  if ((rc = launch_prepg(d.wx, 4, LC, LC, taps, 1, d.wxD, (size_t)LG * taps * 32, st))) return rc;
  const bool jp = d.dpen != nullptr;      // the penalty is part of the graph: its second-order seeds join steps T and T-1
  if (jp && (rc = convlstm_penalty_seed<KS>(d, st))) return rc;
  // head: d out -> dh_T, d bn / conv6
  head_bwd_kernel<<<LC, 512, 0, st>>>(d.h_at(d.T), d.head_stats, d.bn_w, d.bn_b, d.w6, d.dout, d.dh, d.head_grads, d.B, P);
  PTK_LAUNCH_OK("head_bwd_kernel");
  PTK_CUDA_OK(cudaMemsetAsync(d.dc, 0, S * sizeof(float), st));
  for (int t = d.T; t >= 1; --t) {
    const int first = (t == d.T);
    {
      ProfileScope ps(PTK_PROF_POINT_BWD, st);
      lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(d.dh, 0.f, d.dc, 0.f, d.gates_at(t), d.c_at(t - 1), d.c_at(t), d.delta, d.dc,
                                               d.delta_sum, first, d.B, P);
      if (jp && first) lstm_add_kernel<<<148 * 4, 256, 0, st>>>(d.delta, d.delta_sum, d.G, 0, 4, d.dc, d.jp_dcx, d.B, P);
      if (jp && t == d.T - 1) lstm_add_kernel<<<148 * 4, 256, 0, st>>>(d.delta, d.delta_sum, d.jp_eo, 3, 1, nullptr, nullptr, d.B, P);
      PTK_LAUNCH_OK("lstm_bwd_kernel");
    }
    WgradGArgs w{d.h_at(t - 1), d.delta, d.slots_h, d.B, d.H, d.W, LC, LG, d.groups, jp || !first};
    if ((rc = recurrent_wgrad<KS>(d, w, st))) return rc;
    if (t > 1) {
      ConvGArgs c{d.delta, d.whD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0};
      if ((rc = recurrent_conv<KS>(d, c, true, st))) return rc;
    }
  }
  // input path, once: dWx = xt (star) sum delta; db = sum; d xt = sum delta *^T Wx; conv0
  return convlstm_grad_tail<KS>(d, st);
}

// ---- recurrent back-propagation (models/convlstm.py:9-54 `dummyhgru`, :124-135) ----------------------------
// Neumann series on the LAST step's Jacobian J = d h_T / d h_{T-1} (c_{T-1} held fixed):
//   g_0 = v_0 = dL/dh_T;  v_k = J^T v_{k-1};  g_k = g_{k-1} + v_k;   stop (keeping g_{k-1}) when |g_k| > 1,
//   |v_k| > |v_{k-1}| or |v_k| < 1e-9, or after num_iter terms.  Parameter gradients = one cell-step backward
//   with dh_T := g.  The stop test is data dependent; instead of a host round trip per term the state lives on
//   the device (state[0] = |v_{k-1}|, state[1] = terminated, state[2] = accepted terms) and every kernel of a
//   term returns immediately once state[1] is set, so the call never synchronises.
constexpr int RBP_CTAS = 148 * 2;
__global__ void __launch_bounds__(256) rbp_sumsq_kernel(const float* __restrict__ v, const float* __restrict__ g, size_t n,
                                                        double* part, const float* state) {
  if (state != nullptr && state[1] != 0.f) return;
  double sv = 0, sg = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float a = v[i], b = g ? g[i] + a : a;
    sv += (double)a * a;
    sg += (double)b * b;
  }
  __shared__ double s0[256], s1[256];
  s0[threadIdx.x] = sv; s1[threadIdx.x] = sg;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { s0[threadIdx.x] += s0[threadIdx.x + s]; s1[threadIdx.x] += s1[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = s0[0]; part[2 * blockIdx.x + 1] = s1[0]; }
}
__global__ void rbp_decide_kernel(const double* __restrict__ part, int nPart, float* state, int init) {
  if (threadIdx.x != 0 || (!init && state[1] != 0.f)) return;
  double sv = 0, sg = 0;
  for (int i = 0; i < nPart; ++i) { sv += part[2 * i]; sg += part[2 * i + 1]; }
  const float normv = (float)sqrt(sv), normg = (float)sqrt(sg);
  if (init) { state[0] = normv; state[1] = 0.f; state[2] = 0.f; return; }     // normsv[0] = |grad| (:31-34)
  if (normg > 1.0f || normv > state[0] || normv < 1e-9f) state[1] = 1.f;      // :42-46 (g stays g_prev)
  else { state[0] = normv; state[2] += 1.f; }
}
__global__ void rbp_commit_kernel(float* g, float* v_prev, const float* __restrict__ v_new, size_t n, const float* state) {
  if (state[1] != 0.f) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = v_new[i];
    g[i] += v;
    v_prev[i] = v;
  }
}

// input path (time-invariant terms), conv0 and the fold into the flat gradient vector: shared by BPTT and RBP
template <int KS>
static int convlstm_grad_tail(const LstmDev& d, cudaStream_t st) {
  const int P = d.H * d.W, taps = KS * KS;
  const size_t S = (size_t)d.B * LC * P;
  int rc;
  WgradGArgs wx{d.xt, d.delta_sum, d.slots_x, d.B, d.H, d.W, LC, LG, d.groups, 0};
  if ((rc = wide_wgrad<KS>(d, wx, true, st))) return rc;
  chan_sum_kernel<<<LG, 512, 0, st>>>(d.delta_sum, d.bx_grad, d.B, LG, P);
  ConvGArgs cx{d.delta_sum, d.wxD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0};   // dh reused as d xt
  if ((rc = wide_conv<KS>(d, cx, true, true, st))) return rc;
  dsquare_kernel<<<148 * 4, 256, 0, st>>>(d.u, d.dh, S);                                 // d u = 2 u d xt
  WgradGArgs w0{d.x, d.dh, d.slots_0, d.B, d.H, d.W, 1, LC, d.groups, 0};
  if ((rc = launch_wgradg<7>(w0, st))) return rc;
  chan_sum_kernel<<<LC, 512, 0, st>>>(d.dh, d.b0_grad, d.B, LC, P);
  // fold everything into the flat gradient vector (layout: ptk_convlstm_param_offsets)
  const long long per_gate = (long long)LC * LC * taps + LC + (long long)LC * LC * taps;   // Wx.w, Wx.b, Wh.w
  fold_slots_kernel<<<(LC * 49 + 255) / 256, 256, 0, st>>>(d.slots_0, d.groups, LC, 1, 49, d.dparams + d.off_w0, 0, LC);
  if (d.conv_mode == PTK_CONV_FP32)
    fold_slots_kernel<<<(LG * LC * taps + 255) / 256, 256, 0, st>>>(d.slots_x, d.groups, LG, LC, taps, d.dparams + d.off_wx, per_gate, LC);
  else if ((rc = launch_fold_wgradk(d.slots_x_tc, KS, 4, LC, LC, d.dparams + d.off_wx, per_gate, st))) return rc;
  if (d.conv_mode == PTK_CONV_FP32)
    fold_slots_kernel<<<(LG * LC * taps + 255) / 256, 256, 0, st>>>(d.slots_h, d.groups, LG, LC, taps, d.dparams + d.off_wh, per_gate, LC);
  else if ((rc = launch_fold_wgradk(d.slots_h_tc, KS, 4, LC, LC, d.dparams + d.off_wh, per_gate, st))) return rc;
  PTK_CUDA_OK(cudaMemcpyAsync(d.dparams + d.off_b0, d.b0_grad, LC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int k = 0; k < 4; ++k)
    PTK_CUDA_OK(cudaMemcpyAsync(d.dparams + d.off_bx + k * per_gate, d.bx_grad + k * LC, LC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PTK_CUDA_OK(cudaMemcpyAsync(d.dparams + d.off_head, d.head_grads, (4 * LC + 2) * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PTK_LAUNCH_OK("convlstm fold");
  return PTK_OK;
}

template <int KS>
static int convlstm_backward_rbp_t(const LstmDev& d, int num_iter, cudaStream_t st) {
  const int P = d.H * d.W, taps = KS * KS, T = d.T;
  const size_t S = (size_t)d.B * LC * P;
  int rc;
  if ((rc = launch_prepg(d.wh, 4, LC, LC, taps, 1, d.whD, (size_t)LG * taps * 32, st))) return rc;
  if ((rc = recurrent_prep<KS>(d, true, st))) return rc;
  if ((rc = launch_prepg(d.wx, 4, LC, LC, taps, 1, d.wxD, (size_t)LG * taps * 32, st))) return rc;
  const bool jp = d.dpen != nullptr;      // penalty in the graph: one-step seeds (the leaves h_{T-1}, c_{T-1} are detached)
  if (jp && (rc = convlstm_penalty_seed<KS>(d, st))) return rc;
  // head: d out -> dh_T (= grad of dummyhgru.backward), d bn / conv6.  v_prev lives in d.dc, v_new in d.dh.
  head_bwd_kernel<<<LC, 512, 0, st>>>(d.h_at(T), d.head_stats, d.bn_w, d.bn_b, d.w6, d.dout, d.dc, d.head_grads, d.B, P);
  PTK_LAUNCH_OK("head_bwd_kernel");
  PTK_CUDA_OK(cudaMemcpyAsync(d.rbp_g, d.dc, S * sizeof(float), cudaMemcpyDeviceToDevice, st));
  rbp_sumsq_kernel<<<RBP_CTAS, 256, 0, st>>>(d.dc, nullptr, S, d.rbp_part, nullptr);
  rbp_decide_kernel<<<1, 32, 0, st>>>(d.rbp_part, RBP_CTAS, d.rbp_state, 1);
  for (int ii = 0; ii < num_iter; ++ii) {
    lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(d.dc, 0.f, nullptr, 0.f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                             nullptr, nullptr, 0, d.B, P, d.rbp_state);
    ConvGArgs c{d.delta, d.whD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0, d.rbp_state};
    if ((rc = recurrent_conv<KS>(d, c, true, st))) return rc;
    rbp_sumsq_kernel<<<RBP_CTAS, 256, 0, st>>>(d.dh, d.rbp_g, S, d.rbp_part, d.rbp_state);
    rbp_decide_kernel<<<1, 32, 0, st>>>(d.rbp_part, RBP_CTAS, d.rbp_state, 0);
    rbp_commit_kernel<<<148 * 4, 256, 0, st>>>(d.rbp_g, d.dc, d.dh, S, d.rbp_state);
  }
  PTK_LAUNCH_OK("rbp neumann series");
  // one cell-step backward with dh_T := g (dc_T = 0): deltas of the last step only
  lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(d.rbp_g, 0.f, nullptr, 0.f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                           nullptr, d.delta_sum, 1, d.B, P);
  if (jp) lstm_add_kernel<<<148 * 4, 256, 0, st>>>(d.delta, d.delta_sum, d.G, 0, 4, nullptr, nullptr, d.B, P);
  WgradGArgs w{d.h_at(T - 1), d.delta, d.slots_h, d.B, d.H, d.W, LC, LG, d.groups, jp ? 1 : 0};
  if ((rc = recurrent_wgrad<KS>(d, w, st))) return rc;
  return convlstm_grad_tail<KS>(d, st);
}

// d(loss)/d(x): x feeds only conv0 (u = conv0(x), models/convlstm.py:118), so dx = conv_transpose2d(du, conv0.weight, padding = 3)
// with du = 2 u d(xt) as left in `dh` by convlstm_grad_tail.  One thread per pixel, the 25 x 49 taps in shared memory.
__global__ void __launch_bounds__(256) lstm_dx_kernel(const float* __restrict__ du, const float* __restrict__ w0, float* __restrict__ dx,
                                                      int B, int H, int W) {
  __shared__ float sW[LC * 49];
  for (int i = threadIdx.x; i < LC * 49; i += blockDim.x) sW[i] = __ldg(w0 + i);
  __syncthreads();
  const int P = H * W;
  const long long n = (long long)B * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / P), pix = (int)(i - (long long)b * P), y = pix / W, x = pix - y * W;
    float acc = 0.f;
    for (int c = 0; c < LC; ++c) {
      const float* dc = du + ((size_t)b * LC + c) * P;
      const float* wc = sW + c * 49;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = y + 3 - ky;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int xx = x + 3 - kx;
          if (xx >= 0 && xx < W) acc = fmaf(__ldg(dc + (size_t)yy * W + xx), wc[ky * 7 + kx], acc);
        }
      }
    }
    dx[i] = acc;
  }
}
int run_convlstm_dx(const LstmDev& d, float* dx, cudaStream_t st) {
  const long long n = (long long)d.B * d.H * d.W;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  lstm_dx_kernel<<<grid, 256, 0, st>>>(d.dh, d.w0, dx, d.B, d.H, d.W);
  PTK_LAUNCH_OK("lstm_dx_kernel");
  return PTK_OK;
}

int run_convlstm_forward(const LstmDev& d, cudaStream_t st) {
  return d.KS == 15 ? convlstm_forward_t<15>(d, st) : convlstm_forward_t<7>(d, st);
}
int run_convlstm_penalty(const LstmDev& d, cudaStream_t st) {
  return d.KS == 15 ? convlstm_penalty_t<15>(d, st) : convlstm_penalty_t<7>(d, st);
}
int run_convlstm_backward(const LstmDev& d, cudaStream_t st) {
  return d.KS == 15 ? convlstm_backward_t<15>(d, st) : convlstm_backward_t<7>(d, st);
}
int run_convlstm_backward_rbp(const LstmDev& d, int num_iter, cudaStream_t st) {
  return d.KS == 15 ? convlstm_backward_rbp_t<15>(d, num_iter, st) : convlstm_backward_rbp_t<7>(d, num_iter, st);
}

}  // namespace ptk
