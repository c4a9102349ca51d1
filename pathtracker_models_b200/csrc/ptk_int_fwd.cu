// Forward time loop of the InT / hGRU circuit: the point-wise (gate / blend / BatchNorm-apply)
// stages between the two 7x7 convolutions of each step, the stem and the readout.
//
// Reference being replaced (per step): models/InT.py:145-179 (rCell.forward) and
// models/ffhgru_hierarchy.py:135-173 (hConvGRUCell.forward); stem models/InT.py:212-213;
// readout models/InT.py:236-241.  Equation numbers (1)-(13) are those of SURVEY.md section 8.
//
// One thread = one pixel holding all 32 channels in registers; the 1x1 gate convolutions are
// 32x32 mat-vecs against shared-memory weights ([in][out], broadcast float4 reads).
#include "ptk_common.cuh"
#include "ptk_int.h"
#include "ptk_profile.h"

namespace ptk {

__device__ __forceinline__ void smem_fill(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
}

// stem of one pixel: X = nl(Wp x_t + bp)   (models/InT.py:192,212-213)
template <int ACT>
__device__ __forceinline__ void stem_pixel(const FrameView& fv, int b, int t, int y, int x,
                                           const float* sWp, const float* sbp, float (&X)[C]) {
  const float x0 = fv.at(b, 0, t, y, x), x1 = fv.at(b, 1, t, y, x), x2 = fv.at(b, 2, t, y, x);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float pre = fmaf(sWp[c * 3 + 2], x2, fmaf(sWp[c * 3 + 1], x1, fmaf(sWp[c * 3 + 0], x0, sbp[c])));
    X[c] = actf<ACT>(pre);
  }
}

// eqs (1)-(2) for frame t given E_t: A = sigmoid(Wa_w X + Wa_u E), Ge = A*E
template <int ACT>
__device__ __forceinline__ void attention_pixel(const IntDev& d, const float* sWaw, const float* sWau,
                                                const float* sba, const float (&X)[C], const float (&E)[C],
                                                int b, int pix, int P, float* Aout, long long a_bstride) {
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = sba[c];
  matvec32(acc, sWaw, X);
  matvec32(acc, sWau, E);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float A = sigmoidf_(acc[c]);
    Aout[b * a_bstride + (size_t)c * P + pix] = A;
    d.Ge[((size_t)b * C + c) * P + pix] = A * E[c];
  }
}

template <int ACT>
__global__ void __launch_bounds__(PT) int_pre_kernel(IntDev d, int t) {
  __shared__ __align__(16) float sWaw[C * C], sWau[C * C];
  __shared__ float sWp[C * 3], sbp[C], sba[C];
  smem_fill(sWaw, d.gateT[0], C * C);
  smem_fill(sWau, d.gateT[1], C * C);
  smem_fill(sWp, d.Wp, C * 3);
  smem_fill(sbp, d.bp, C);
  if (threadIdx.x < C) sba[threadIdx.x] = __ldg(d.gate_b[0] + threadIdx.x) + __ldg(d.gate_b[1] + threadIdx.x);
  __syncthreads();
  const int P = d.H * d.W, tilesPer = P / PT;
  const int b = blockIdx.x / tilesPer, pix = (blockIdx.x % tilesPer) * PT + threadIdx.x;
  float X[C], E[C];
  stem_pixel<ACT>(d.fv, b, t, pix / d.W, pix % d.W, sWp, sbp, X);
  const float* Et = d.ckE + d.ck_stride * d.slot(t);
#pragma unroll
  for (int c = 0; c < C; ++c) E[c] = Et[((size_t)b * C + c) * P + pix];
  attention_pixel<ACT>(d, sWaw, sWau, sba, X, E, b, pix, P, d.A + (size_t)t * d.a_tstride, d.a_bstride);
}

// eqs (3b)-(9): BatchNorm0 apply, inhibition drive, inhibition gate + blend, excitation gate.
template <int CELL, int ACT>
__global__ void __launch_bounds__(PT) int_mid_kernel(IntDev d, int t) {
  __shared__ __align__(16) float sWiw[C * C], sWiu[C * C], sWew[C * C], sWeu[C * C];
  __shared__ float sWp[C * 3], sbp[C], sbi[C], sbe[C], sAl[C], sMu[C], sScale[C], sShift[C];
  smem_fill(sWiw, d.gateT[2], C * C);
  smem_fill(sWiu, d.gateT[3], C * C);
  smem_fill(sWew, d.gateT[4], C * C);
  smem_fill(sWeu, d.gateT[5], C * C);
  smem_fill(sWp, d.Wp, C * 3);
  smem_fill(sbp, d.bp, C);
  smem_fill(sAl, d.alpha, C);
  smem_fill(sMu, d.mu, C);
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    sbi[c] = __ldg(d.gate_b[2] + c) + __ldg(d.gate_b[3] + c);
    sbe[c] = __ldg(d.gate_b[4] + c) + __ldg(d.gate_b[5] + c);
    const float* st = d.stat(t, 0);   // BN0 of step t: mean, rstd
    const float sc = __ldg(d.bn_w[0] + c) * st[C + c];
    sScale[c] = sc;                                            // U = Y0*scale + shift
    sShift[c] = __ldg(d.bn_b[0] + c) - st[c] * sc;
  }
  __syncthreads();
  const int P = d.H * d.W, tilesPer = P / PT;
  const int b = blockIdx.x / tilesPer, pix = (blockIdx.x % tilesPer) * PT + threadIdx.x;
  const size_t base = (size_t)b * C * P + pix;
  const int slot_in = d.slot(t), slot_out = d.slot(t + 1);
  const float* It = d.ckI + d.ck_stride * slot_in;
  float* Iout = d.ckI + d.ck_stride * slot_out;
  const float* At = d.A + (size_t)t * d.a_tstride + (size_t)b * d.a_bstride + pix;

  float X[C], I[C], S[C];
  stem_pixel<ACT>(d.fv, b, t, pix / d.W, pix % d.W, sWp, sbp, X);
#pragma unroll
  for (int c = 0; c < C; ++c) I[c] = It[base + (size_t)c * P];
#pragma unroll
  for (int c = 0; c < C; ++c) S[c] = (CELL == PTK_CELL_INT) ? I[c] : At[(size_t)c * P];
  {
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = sbi[c];
    matvec32(acc, sWiw, X);
    matvec32(acc, sWiu, S);
    const float* Y0 = d.Y0 + d.y_stride * d.yslot(t);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float Gi = sigmoidf_(acc[c]);
      const float U = fmaf(Y0[base + (size_t)c * P], sScale[c], sShift[c]);
      const float Pp = U * fmaf(sAl[c], S[c], sMu[c]);
      const float Q = actf<ACT>(Pp);
      const float Ihat = actf<ACT>(X[c] - Q);
      Iout[base + (size_t)c * P] = (1.0f - Gi) * I[c] + Gi * Ihat;
    }
  }
  {
    float acc[C], Ge[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = sbe[c];
    matvec32(acc, sWew, S);
#pragma unroll
    for (int c = 0; c < C; ++c) Ge[c] = d.Ge[base + (size_t)c * P];
    matvec32(acc, sWeu, Ge);
#pragma unroll
    for (int c = 0; c < C; ++c) d.Gex[base + (size_t)c * P] = sigmoidf_(acc[c]);
  }
}

// eqs (10b)-(13) for step t, then (1)-(2) for step t+1 (if any).
template <int ACT>
__global__ void __launch_bounds__(PT) int_post_kernel(IntDev d, int t, int has_next) {
  __shared__ __align__(16) float sWaw[C * C], sWau[C * C];
  __shared__ float sWp[C * 3], sbp[C], sba[C], sKa[C], sGa[C], sScale[C], sShift[C], sWr[C];
  smem_fill(sWaw, d.gateT[0], C * C);
  smem_fill(sWau, d.gateT[1], C * C);
  smem_fill(sWp, d.Wp, C * 3);
  smem_fill(sbp, d.bp, C);
  smem_fill(sKa, d.kappa, C);
  smem_fill(sGa, d.gamma, C);
  smem_fill(sWr, d.Wr, C);
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    sba[c] = __ldg(d.gate_b[0] + c) + __ldg(d.gate_b[1] + c);
    const float* st = d.stat(t, 1);   // BN1 of step t
    const float sc = __ldg(d.bn_w[1] + c) * st[C + c];
    sScale[c] = sc;
    sShift[c] = __ldg(d.bn_b[1] + c) - st[c] * sc;
  }
  __syncthreads();
  const int P = d.H * d.W, tilesPer = P / PT;
  const int b = blockIdx.x / tilesPer, pix = (blockIdx.x % tilesPer) * PT + threadIdx.x;
  const size_t base = (size_t)b * C * P + pix;
  const int slot_in = d.slot(t), slot_out = d.slot(t + 1);
  const float* Et = d.ckE + d.ck_stride * slot_in;
  const float* I2 = d.ckI + d.ck_stride * slot_out;
  float* Eout = d.ckE + d.ck_stride * slot_out;
  const float* Y1 = d.Y1 + d.y_stride * d.yslot(t);

  float E2[C];
  float sread = __ldg(d.br);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float V = fmaf(Y1[base + (size_t)c * P], sScale[c], sShift[c]);
    const float R = V * fmaf(sKa[c], I2[base + (size_t)c * P], sGa[c]);
    const float Ehat = actf<ACT>(R);
    const float Gx = d.Gex[base + (size_t)c * P];
    const float e = (1.0f - Gx) * Et[base + (size_t)c * P] + Gx * Ehat;
    E2[c] = e;
    Eout[base + (size_t)c * P] = e;
    sread = fmaf(sWr[c], e, sread);
  }
  if (d.states != nullptr)  // testmode: states.append(readout_conv(excitation)), InT.py:233
    d.states[((size_t)b * d.T + t) * P + pix] = sread;
  if (has_next) {
    float X[C];
    stem_pixel<ACT>(d.fv, b, t + 1, pix / d.W, pix % d.W, sWp, sbp, X);
    attention_pixel<ACT>(d, sWaw, sWau, sba, X, E2, b, pix, P, d.A + (size_t)(t + 1) * d.a_tstride, d.a_bstride);
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
__global__ void prep_gates_kernel(IntGateSrc g, float* gateT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 6 * C * C) return;
  const int gi = i / (C * C), o = (i / C) % C, in = i % C;
  gateT[gi * C * C + in * C + o] = g.w[gi][o * C + in];
}

// ---- host ----------------------------------------------------------------------------------
template <int ACT>
static int run_forward_act(const IntDev& d, int cell, cudaStream_t st) {
  const int P = d.H * d.W;
  const int tiles = d.B * P / PT;
  { ProfileScope ps(PTK_PROF_POINT_FWD, st);
  int_pre_kernel<ACT><<<tiles, PT, 0, st>>>(d, 0); }
  PTK_LAUNCH_OK("int_pre_kernel");
  for (int t = 0; t < d.T; ++t) {
    const int slot_in = d.slot(t), slot_out = d.slot(t + 1);
    ConvArgs c0{d.Ge, d.w_inh_f, d.Y0 + d.y_stride * d.yslot(t), d.B, d.H, d.W, 0,
                d.bn_part, d.counter, d.stat_w(t, 0)};
    int rc;
    { ProfileScope ps(PTK_PROF_CONV_FWD, st); rc = launch_conv7(c0, st); }
    if (rc) return rc;
    { ProfileScope ps(PTK_PROF_POINT_FWD, st);
    if (cell == PTK_CELL_INT) int_mid_kernel<PTK_CELL_INT, ACT><<<tiles, PT, 0, st>>>(d, t);
    else int_mid_kernel<PTK_CELL_HGRU, ACT><<<tiles, PT, 0, st>>>(d, t); }
    PTK_LAUNCH_OK("int_mid_kernel");
    ConvArgs c1{d.ckI + d.ck_stride * slot_out, d.w_exc_f, d.Y1 + d.y_stride * d.yslot(t), d.B, d.H,
                d.W, 0, d.bn_part, d.counter, d.stat_w(t, 1)};
    { ProfileScope ps(PTK_PROF_CONV_FWD, st); rc = launch_conv7(c1, st); }
    if (rc) return rc;
    { ProfileScope ps(PTK_PROF_POINT_FWD, st);
    int_post_kernel<ACT><<<tiles, PT, 0, st>>>(d, t, t + 1 < d.T); }
    PTK_LAUNCH_OK("int_post_kernel");
    (void)slot_in;
  }
  const int slot_T = d.slot(d.T);
  const size_t sm = (size_t)2 * (d.H + 4) * (d.W + 4) * sizeof(float);
  { ProfileScope ps(PTK_PROF_MISC, st);
  int_readout_kernel<<<d.B, 256, sm, st>>>(d, d.ckE + d.ck_stride * slot_T, d.logits); }
  PTK_LAUNCH_OK("int_readout_kernel");
  return PTK_OK;
}

int run_int_forward(const IntDev& d, int cell, int act, cudaStream_t st) {
  if (act == PTK_ACT_SOFTPLUS) return run_forward_act<PTK_ACT_SOFTPLUS>(d, cell, st);
  return run_forward_act<PTK_ACT_TANH>(d, cell, st);
}

int run_prep_weights(const float* w_inh, const float* w_exc, float* inh_f, float* inh_d, float* exc_f,
                     float* exc_d, const IntGateSrc& g, float* gateT, cudaStream_t st) {
  const int n = C * C * TAPS;
  ProfileScope ps(PTK_PROF_MISC, st, 3);
  prep_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w_inh, inh_f, inh_d);
  prep_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w_exc, exc_f, exc_d);
  prep_gates_kernel<<<(6 * C * C + 255) / 256, 256, 0, st>>>(g, gateT);
  PTK_LAUNCH_OK("prep_weights");
  return PTK_OK;
}

}  // namespace ptk
