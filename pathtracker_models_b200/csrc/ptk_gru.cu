// ConvGRU baseline (models/kys.py; `--model gru`, utils/engine.py:147-153) on the GPU: first version, strict fp32.
//
//   forward  (GRU.forward :104-135, ConvGRUCell.forward :36-48)
//       X_t = softplus(preproc(x_t))                                   (:106-108; never materialised for all t)
//       for t: [r|u] = sigmoid([Wr|Wu] * cat(X_t, h) + b)               (:37-40, one stacked 128 -> 128 convolution)
//              n     = tanh(Wn * cat(X_t, r h) + b)                     (:42-43)
//              h     = (1 - u) h + u n                                  (:45)
//   backward (autograd of the above): per step two transposed convolutions and two weight-gradient kernels.
//   The readout head (:127-133) is outside the time loop and stays with the caller.
// Convolutions: the generic direct kernels of ptk_convg.cuh (shared with the ConvLSTM path).
#include "ptk_convg.cuh"
#include "ptk_gru.h"
#include "ptk_kernels.h"

namespace ptk {

#define GRU_FOR_EACH(e, n) for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (n); e += (size_t)gridDim.x * blockDim.x)

__device__ __forceinline__ float gru_stem_pre(const GruDev& d, int b, int c, int frame, int pix) {
  const int y = pix / d.W, x = pix - y * d.W;
  return fmaf(__ldg(d.Wp + c * 3 + 2), d.fv.at(b, 2, frame, y, x),
              fmaf(__ldg(d.Wp + c * 3 + 1), d.fv.at(b, 1, frame, y, x),
                   fmaf(__ldg(d.Wp + c * 3), d.fv.at(b, 0, frame, y, x), __ldg(d.bp + c))));
}

// Z = cat(X_t, h_{t-1}),  Z2[:, :64] = X_t      (step t = 1..T reads frame t-1)
__global__ void gru_stem_cat_kernel(GruDev d, int t, int with_rh) {
  const int P = d.H * d.W;
  const float* hp = d.h_at(t - 1);
  GRU_FOR_EACH(e, d.S()) {
    const int pix = (int)(e % P), c = (int)((e / P) % GC), b = (int)(e / ((size_t)P * GC));
    const float X = softplus_precise(gru_stem_pre(d, b, c, t - 1, pix));
    const size_t z = ((size_t)b * 2 * GC + c) * P + pix;
    d.Z[z] = X;
    d.Z2[z] = X;
    const float h = hp[e];
    d.Z[z + (size_t)GC * P] = h;
    if (with_rh) d.Z2[z + (size_t)GC * P] = d.r_at(t)[e] * h;       // backward: rebuild cat(X_t, r h) from the saved gate
  }
}
__global__ void gru_gate_kernel(GruDev d, int t, int add_bias) {
  const int P = d.H * d.W;
  const float* hp = d.h_at(t - 1);
  GRU_FOR_EACH(e, d.S()) {
    const int pix = (int)(e % P), c = (int)((e / P) % GC), b = (int)(e / ((size_t)P * GC));
    const size_t z = ((size_t)b * 2 * GC + c) * P + pix;
    const float br = add_bias ? __ldg(d.br + c) : 0.f, bu = add_bias ? __ldg(d.bu + c) : 0.f;
    const float r = sigmoid_precise(d.RU[z] + br), u = sigmoid_precise(d.RU[z + (size_t)GC * P] + bu);
    d.r_at(t)[e] = r;
    d.u_at(t)[e] = u;
    d.Z2[z + (size_t)GC * P] = r * hp[e];
  }
}
__global__ void gru_update_kernel(GruDev d, int t, int add_bias) {
  const int P = d.H * d.W;
  const float* hp = d.h_at(t - 1);
  float* ho = d.h_at(t);
  GRU_FOR_EACH(e, d.S()) {
    const float bn = add_bias ? __ldg(d.bnw + (int)((e / P) % GC)) : 0.f;
    const float n = tanh_precise(d.Nb[e] + bn), u = d.u_at(t)[e];
    d.n_at(t)[e] = n;
    ho[e] = (1.0f - u) * hp[e] + u * n;
  }
}

// ---- backward -------------------------------------------------------------------------------------------------
// from dh_t: dN (into Nb), the update half of dRU, dh_{t-1} partial
__global__ void gru_bwd1_kernel(GruDev d, int t, int first) {
  const int P = d.H * d.W;
  const float* hp = d.h_at(t - 1);
  GRU_FOR_EACH(e, d.S()) {
    const int pix = (int)(e % P), c = (int)((e / P) % GC), b = (int)(e / ((size_t)P * GC));
    const size_t z = ((size_t)b * 2 * GC + c) * P + pix;
    const float dh = d.dh[e], u = d.u_at(t)[e], n = d.n_at(t)[e];
    const float dN = dh * u * (1.0f - n * n);
    d.Nb[e] = dN;
    d.dNsum[e] = first ? dN : d.dNsum[e] + dN;
    const float dU = dh * (n - hp[e]) * u * (1.0f - u);
    d.RU[z + (size_t)GC * P] = dU;
    d.dRUsum[z + (size_t)GC * P] = first ? dU : d.dRUsum[z + (size_t)GC * P] + dU;
    d.dhp[e] = dh * (1.0f - u);
  }
}
// from dZ2 = dN *^T Wn: dX partial, the reset half of dRU, dh_{t-1} += d(r h) r
__global__ void gru_bwd2_kernel(GruDev d, int t, int first) {
  const int P = d.H * d.W;
  const float* hp = d.h_at(t - 1);
  GRU_FOR_EACH(e, d.S()) {
    const int pix = (int)(e % P), c = (int)((e / P) % GC), b = (int)(e / ((size_t)P * GC));
    const size_t z = ((size_t)b * 2 * GC + c) * P + pix;
    d.dXb[e] = d.dZ[z];
    const float drh = d.dZ[z + (size_t)GC * P], r = d.r_at(t)[e];
    d.dhp[e] += drh * r;
    const float dR = drh * hp[e] * r * (1.0f - r);
    d.RU[z] = dR;
    d.dRUsum[z] = first ? dR : d.dRUsum[z] + dR;
  }
}
// from dZ = dRU *^T [Wr|Wu]: dh_{t-1} complete, d(stem pre-activation) into dXb
__global__ void gru_bwd3_kernel(GruDev d, int t) {
  const int P = d.H * d.W;
  GRU_FOR_EACH(e, d.S()) {
    const int pix = (int)(e % P), c = (int)((e / P) % GC), b = (int)(e / ((size_t)P * GC));
    const size_t z = ((size_t)b * 2 * GC + c) * P + pix;
    d.dh[e] = d.dhp[e] + d.dZ[z + (size_t)GC * P];
    const float pre = gru_stem_pre(d, b, c, t - 1, pix);
    const float dX = d.dXb[e] + d.dZ[z];
    d.dXb[e] = dX * (pre > 20.0f ? 1.0f : sigmoid_precise(pre));      // softplus' (threshold 20, models/kys.py:102)
  }
}
// stem parameter gradients of one step: GRU_STEM_SPLIT CTAs per channel (64 CTAs left most of the GPU idle for
// 150 us per step), each owning a fixed slice of (b, pixel) and its own accumulator => deterministic
__global__ void __launch_bounds__(512) gru_stem_grad_kernel(GruDev d, int t, int first) {
  __shared__ double sh[16];
  const int c = blockIdx.x / GRU_STEM_SPLIT, part = blockIdx.x % GRU_STEM_SPLIT, P = d.H * d.W;
  const size_t Nall = (size_t)d.B * P, per = (Nall + GRU_STEM_SPLIT - 1) / GRU_STEM_SPLIT;
  const size_t i0 = (size_t)part * per, N = i0 + per < Nall ? i0 + per : Nall;
  float* acc = d.stem_acc + (size_t)part * 4 * GC;
  double s[4] = {0, 0, 0, 0};
  for (size_t i = i0 + threadIdx.x; i < N; i += blockDim.x) {
    const int b = (int)(i / P), pix = (int)(i % P), y = pix / d.W, x = pix - y * d.W;
    const float g = d.dXb[((size_t)b * GC + c) * P + pix];
    s[0] += g;
    s[1] += (double)g * d.fv.at(b, 0, t - 1, y, x);
    s[2] += (double)g * d.fv.at(b, 1, t - 1, y, x);
    s[3] += (double)g * d.fv.at(b, 2, t - 1, y, x);
  }
  for (int k = 0; k < 4; ++k) {
    const double tot = block_sum_d(s[k], sh);
    if (threadIdx.x == 0) acc[c * 4 + k] = first ? (float)tot : acc[c * 4 + k] + (float)tot;
    __syncthreads();
  }
}
// optional d loss / d x of the stem (contiguous [B,3,T,H,W])
__global__ void gru_dx_kernel(GruDev d, int t) {
  const int P = d.H * d.W;
  const size_t n = (size_t)d.B * 3 * P;
  GRU_FOR_EACH(e, n) {
    const int pix = (int)(e % P), i = (int)((e / P) % 3), b = (int)(e / ((size_t)P * 3));
    float s = 0.f;
    for (int c = 0; c < GC; ++c) s = fmaf(__ldg(d.Wp + c * 3 + i), d.dXb[((size_t)b * GC + c) * P + pix], s);
    if (t == 1 && i == 2) s += d.dgoal[(size_t)b * P + pix];     // x[:,2,0] also feeds the readout (models/kys.py:127)
    d.dx[(((size_t)b * 3 + i) * d.T + (t - 1)) * P + pix] = s;
  }
}
__global__ void gru_stem_out_kernel(GruDev d) {      // stem_acc [64][4] -> dparams (preproc.weight [64][3], preproc.bias [64])
  const int c = threadIdx.x;
  if (c >= GC) return;
  double s[4] = {0, 0, 0, 0};
  for (int part = 0; part < GRU_STEM_SPLIT; ++part)
    for (int k = 0; k < 4; ++k) s[k] += (double)d.stem_acc[(size_t)part * 4 * GC + c * 4 + k];
  d.dparams[d.off_bp + c] = (float)s[0];
  for (int i = 0; i < 3; ++i) d.dparams[d.off_wp + c * 3 + i] = (float)s[1 + i];
}

// ---- readout head (models/kys.py:127-133): r = readout_conv(h_T); cat with the goal channel x[:,2,0]; 5x5 conv (pad 2);
//      global mean; Linear(1,1).  One CTA per sample (same scheme as the InT readout, 64 channels).
__device__ __forceinline__ void gru_head_maps(const GruDev& d, int b, float* sR, float* sG, float* sWt, float* sWr) {
  const int H = d.H, W = d.W, P = H * W, Wp2 = W + 4, Hp2 = H + 4, tid = threadIdx.x;
  const float* hT = d.h_at(d.T);
  if (tid < 50) sWt[tid] = __ldg(d.hWt + tid);
  if (tid < GC) sWr[tid] = __ldg(d.hWr + tid);
  for (int i = tid; i < 2 * Hp2 * Wp2; i += 256) sR[i] = 0.f;      // sR and sG are contiguous
  __syncthreads();
  const float br = __ldg(d.hbr);
  for (int pix = tid; pix < P; pix += 256) {
    float r = br;
#pragma unroll 8
    for (int c = 0; c < GC; ++c) r = fmaf(sWr[c], hT[((size_t)b * GC + c) * P + pix], r);
    const int y = pix / W, x = pix % W;
    sR[(y + 2) * Wp2 + x + 2] = r;
    sG[(y + 2) * Wp2 + x + 2] = d.fv.at(b, 2, 0, y, x);
  }
  __syncthreads();
}
__device__ __forceinline__ float gru_head_conv_sum(const GruDev& d, const float* sR, const float* sG, const float* sWt) {
  const int W = d.W, P = d.H * d.W, Wp2 = W + 4;
  float s = 0.f;
  for (int pix = threadIdx.x; pix < P; pix += 256) {
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
  return s;
}
__device__ __forceinline__ float gru_block_sum(float v, float* sRed) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += sRed[i];
  return tot;
}
__global__ void __launch_bounds__(256) gru_readout_kernel(GruDev d) {
  extern __shared__ float sm[];
  const int P = d.H * d.W;
  float* sR = sm;
  float* sG = sm + (d.H + 4) * (d.W + 4);
  __shared__ float sWt[50], sWr[GC], sRed[8];
  const int b = blockIdx.x;
  gru_head_maps(d, b, sR, sG, sWt, sWr);
  const float tot = gru_block_sum(gru_head_conv_sum(d, sR, sG, sWt), sRed);
  if (threadIdx.x == 0) d.logits[b] = fmaf(__ldg(d.hWd), tot / (float)P + __ldg(d.hbt), __ldg(d.hbd));
}
// reverse of the head: dlogits -> d h_T (into d.dh), per-sample head gradients, d x[:,2,0]
__global__ void __launch_bounds__(256) gru_readout_bwd_kernel(GruDev d) {
  extern __shared__ float sm[];
  const int H = d.H, W = d.W, P = H * W, Wp2 = W + 4, Hp2 = H + 4;
  float* sR = sm;
  float* sG = sm + Hp2 * Wp2;
  float* sDr = sG + Hp2 * Wp2;    // [P]
  __shared__ float sWt[50], sWr[GC], sRed[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* hT = d.h_at(d.T);
  gru_head_maps(d, b, sR, sG, sWt, sWr);
  const float pooled = gru_block_sum(gru_head_conv_sum(d, sR, sG, sWt), sRed) / (float)P + __ldg(d.hbt);
  const float dl = d.dlogits[b];
  const float dpool = dl * __ldg(d.hWd);
  const float dov = dpool / (float)P;                      // d(5x5 conv output), constant over the image
  float* hd = d.head_part + (size_t)b * GRU_HEAD_N;        // order: dWr[64], dbr, dWt[50], dbt, dWd, dbd
  if (tid == 0) {
    hd[GC + 1 + 50] = dpool;        // d target_conv.bias
    hd[GC + 1 + 51] = dl * pooled;  // d readout_dense.weight
    hd[GC + 1 + 52] = dl;           // d readout_dense.bias
  }
  for (int w = 0; w < 50; ++w) {    // d target_conv.weight: window sums of the padded maps
    const int ch = w / 25, ky = (w % 25) / 5, kx = w % 5;
    const float* srcm = ch ? sG : sR;
    float v = 0.f;
    for (int pix = tid; pix < P; pix += 256) v += srcm[(pix / W + ky) * Wp2 + pix % W + kx];
    v = gru_block_sum(v, sRed);
    if (tid == 0) hd[GC + 1 + w] = dov * v;
  }
  float sbr = 0.f;
  for (int pix = tid; pix < P; pix += 256) {
    const int y = pix / W, x = pix % W;
    float wsum = 0.f, gsum = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky)
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int yy = y - ky + 2, xx = x - kx + 2;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) { wsum += sWt[ky * 5 + kx]; gsum += sWt[25 + ky * 5 + kx]; }
      }
    const float dr = dov * wsum;
    d.dgoal[(size_t)b * P + pix] = dov * gsum;             // d x[b,2,0,y,x] via torch.cat(..., x[:,2,0]) (kys.py:127)
    sDr[pix] = dr;
    sbr += dr;
  }
  sbr = gru_block_sum(sbr, sRed);
  if (tid == 0) hd[GC] = sbr;                              // d readout_conv.bias
  __syncthreads();
  for (int c = 0; c < GC; ++c) {                           // d h_T = Wr dr; d readout_conv.weight[c] = sum dr h_T[c]
    float v = 0.f;
    for (int pix = tid; pix < P; pix += 256) {
      const size_t idx = ((size_t)b * GC + c) * P + pix;
      v = fmaf(sDr[pix], hT[idx], v);
      d.dh[idx] = sWr[c] * sDr[pix];
    }
    v = gru_block_sum(v, sRed);
    if (tid == 0) hd[c] = v;
  }
}
__global__ void gru_head_fold_kernel(GruDev d) {     // per-sample head gradients -> dparams (fp64, fixed order)
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= GRU_HEAD_N) return;
  double s = 0;
  for (int b = 0; b < d.B; ++b) s += (double)d.head_part[(size_t)b * GRU_HEAD_N + j];
  d.dparams[d.off_head + j] = (float)s;
}

// ---- tensor-core modes: the wide convolutions as 32 x 32-channel blocks on the tcgen05 kernels of the InT path ----
// [reset|update]: out blocks j = 0..3 (j < 2 reset rows, j >= 2 update rows), in blocks i = 0..3, pair = j*4 + i;
// candidate: out blocks j = 0..1, pair = 16 + j*4 + i.  Each block convolution addresses its 32 channels inside the
// wider [B,Cw,H,W] tensors through the per-sample strides of ConvArgs / WgradArgs; biases are added by the gate kernels.
constexpr size_t GRU_WTC_PAIR = (size_t)4 * C * C * TAPS;      // bytes: bf16 hi + lo
static const float* gru_wblock(const GruDev& d, int pair) {
  const int j = (pair & 15) >> 2, i = pair & 3;
  const float* w = pair >= 16 ? d.wn : (j < 2 ? d.wr : d.wu);
  const int jj = pair >= 16 ? j : (j & 1);
  return w + (size_t)(jj * 32) * (2 * GC * TAPS) + (size_t)i * 32 * TAPS;
}
static int gru_prep_tc(const GruDev& d, cudaStream_t st) {
  for (int pair = 0; pair < 24; ++pair) {
    int rc = launch_prep_weights_tc(gru_wblock(d, pair), d.wtcF + pair * GRU_WTC_PAIR, d.wtcD + pair * GRU_WTC_PAIR, st,
                                    2 * GC * TAPS);
    if (rc) return rc;
  }
  return PTK_OK;
}
// out[B, nOut*32, .] = conv(in[B, nIn*32, .]) (transpose = 0) or its data gradient (transpose = 1: in has nOut blocks)
static int gru_conv_tc(const GruDev& d, const float* in, float* out, int pair0, int nOut, int nIn, int transpose, cudaStream_t st) {
  const size_t P = (size_t)d.H * d.W;
  const int nSrc = transpose ? nOut : nIn, nDst = transpose ? nIn : nOut;
  for (int o = 0; o < nDst; ++o)
    for (int s = 0; s < nSrc; ++s) {
      const int pair = pair0 + (transpose ? s * 4 + o : o * 4 + s);
      ConvArgs a{in + (size_t)s * 32 * P, reinterpret_cast<const float*>((transpose ? d.wtcD : d.wtcF) + pair * GRU_WTC_PAIR),
                 out + (size_t)o * 32 * P, d.B, d.H, d.W, s > 0, nullptr, nullptr, nullptr, d.conv_mode, 1 /* PDL chain */,
                 (long long)(nSrc * 32 * P), (long long)(nDst * 32 * P)};
      int rc;
      { ProfileScope ps(transpose ? PTK_PROF_CONV_DGRAD : PTK_PROF_CONV_FWD, st); rc = launch_conv7(a, st); }
      if (rc) return rc;
    }
  return PTK_OK;
}
static int gru_wgrad_tc(const GruDev& d, const float* in, const float* dy, int pair0, int nOut, int first, cudaStream_t st) {
  const size_t P = (size_t)d.H * d.W;
  for (int j = 0; j < nOut; ++j)
    for (int i = 0; i < 4; ++i) {
      WgradArgs w{in + (size_t)i * 32 * P, dy + (size_t)j * 32 * P, d.wacc_tc + (size_t)(pair0 + j * 4 + i) * d.wacc_pair,
                  d.B, d.H, d.W, first, d.conv_mode, (long long)(4 * 32 * P), (long long)(nOut * 32 * P)};
      int rc;
      { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgrad7_tc(w, st); }
      if (rc) return rc;
    }
  return PTK_OK;
}
// fold the per-CTA slots of all 24 pairs (fp64, fixed order) into conv_reset / conv_update / conv_state_new .weight
__global__ void gru_fold_tc_kernel(GruDev d, int nCta) {
  const size_t n = (size_t)3 * GC * 2 * GC * TAPS;
  GRU_FOR_EACH(e, n) {
    const int tap = (int)(e % TAPS), ic = (int)((e / TAPS) % (2 * GC)), oc = (int)((e / ((size_t)TAPS * 2 * GC)) % GC);
    const int which = (int)(e / ((size_t)TAPS * 2 * GC * GC));              // 0 reset, 1 update, 2 candidate
    const int j = (which == 1 ? 2 : 0) + (oc >> 5), pair = (which == 2 ? 16 : 0) + j * 4 + (ic >> 5);
    const float* wa = d.wacc_tc + (size_t)pair * d.wacc_pair + (size_t)tap * C * C + (oc & 31) * C + (ic & 31);
    double s = 0;
    for (int g = 0; g < nCta; ++g) s += (double)wa[(size_t)g * TAPS * C * C];
    const long long off = which == 0 ? d.off_wr : (which == 1 ? d.off_wu : d.off_wn);
    d.dparams[off + ((size_t)oc * 2 * GC + ic) * TAPS + tap] = (float)s;
  }
}

// ---- bf16x3 mode: the wide convolutions as single launches of the ConvLSTM's tcgen05 kernels --------------------------
// forward [r|u]: N = 128 stacked output channels, the 128 input channels walked in four 32-channel blocks (accumulators
// stay in TMEM); candidate: N = 64; transposes: N = 128; weight gradients: one launch per 32-channel input block
// (dY = 128 resp. 64 channels on the M axis), slots folded at the end.  Biases are added by the gate kernels.
static int gru_prep_wide(const GruDev& d, bool transposed, cudaStream_t st) {
  const float* ru[2] = {d.wr, d.wu};
  const float* nn[1] = {d.wn};
  int rc;
  if ((rc = launch_prep_weights_k_generic(ru, 2, GC, 2 * GC, 7, 128, 4, 0, d.wk_ruF, st))) return rc;
  if ((rc = launch_prep_weights_k_generic(nn, 1, GC, 2 * GC, 7, 64, 4, 0, d.wk_nF, st))) return rc;
  if (transposed) {
    if ((rc = launch_prep_weights_k_generic(ru, 2, GC, 2 * GC, 7, 128, 4, 1, d.wk_ruD, st))) return rc;
    if ((rc = launch_prep_weights_k_generic(nn, 1, GC, 2 * GC, 7, 128, 2, 1, d.wk_nD, st))) return rc;
  }
  return PTK_OK;
}
template <int NT>
static int gru_conv_wide(const GruDev& d, const float* in, const void* w, float* out, int Cin, int Cout, bool transposed, cudaStream_t st) {
  ConvKArgs k{in, w, nullptr, nullptr, out, d.B, d.H, d.W, Cin, Cout, 0, nullptr};
  ProfileScope ps(transposed ? PTK_PROF_CONV_DGRAD : PTK_PROF_CONV_FWD, st);
  return launch_convk_tc_n<7, NT>(k, st);
}
static int gru_wgrad_wide(const GruDev& d, const float* in, const float* dy, int Cout, int slot0, int first, cudaStream_t st) {
  const size_t P = (size_t)d.H * d.W, SF = wgradk_tc_slot_floats(7);
  int rc0 = launch_wgradk_prep_dy(dy, d.B, d.H, d.W, Cout, d.dyc_k, nullptr, st);   // read by 4 input blocks x 7 kernel rows
  if (rc0) return rc0;
  for (int i = 0; i < 4; ++i) {
    WgradKArgs k{in + (size_t)i * 32 * P, dy, d.slots_k + (size_t)(slot0 + i) * SF, d.B, d.H, d.W, 32, Cout, !first, nullptr,
                 (long long)(2 * GC * P), d.dyc_k};
    int rc;
    { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgradk_tc<7>(k, st); }
    if (rc) return rc;
  }
  return PTK_OK;
}
static int gru_fold_wide(const GruDev& d, cudaStream_t st) {
  const size_t SF = wgradk_tc_slot_floats(7);
  int rc;
  for (int i = 0; i < 4; ++i) {
    if ((rc = launch_fold_wgradk(d.slots_k + (size_t)i * SF, 7, 2, GC, 32, d.dparams + d.off_wr, d.off_wu - d.off_wr, st, i * 32, 2 * GC))) return rc;
    if ((rc = launch_fold_wgradk(d.slots_k + (size_t)(4 + i) * SF, 7, 1, GC, 32, d.dparams + d.off_wn, 0, st, i * 32, 2 * GC))) return rc;
  }
  return PTK_OK;
}

static int gru_prep(const GruDev& d, bool transposed, cudaStream_t st) {
  const int taps = 49;
  int rc;
  const float* ru[2] = {d.wr, d.wu};
  const float* nn[1] = {d.wn};
  if ((rc = launch_prepg(ru, 2, GC, 2 * GC, taps, 0, d.wruT, (size_t)4 * 2 * GC * taps * 32, st))) return rc;
  if ((rc = launch_prepg(nn, 1, GC, 2 * GC, taps, 0, d.wnT, (size_t)2 * 2 * GC * taps * 32, st))) return rc;
  if (transposed) {
    if ((rc = launch_prepg(ru, 2, GC, 2 * GC, taps, 1, d.wruD, (size_t)4 * 2 * GC * taps * 32, st))) return rc;
    if ((rc = launch_prepg(nn, 1, GC, 2 * GC, taps, 1, d.wnD, (size_t)4 * GC * taps * 32, st))) return rc;
  }
  PTK_CUDA_OK(cudaMemcpyAsync(d.bru, d.br, GC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PTK_CUDA_OK(cudaMemcpyAsync(d.bru + GC, d.bu, GC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return PTK_OK;
}

int run_gru_forward(const GruDev& d, cudaStream_t st) {
  int rc;
  const bool tc = d.conv_mode != PTK_CONV_FP32, wide = d.conv_mode == PTK_CONV_BF16X3;
  if ((rc = wide ? gru_prep_wide(d, false, st) : (tc ? gru_prep_tc(d, st) : gru_prep(d, false, st)))) return rc;
  PTK_CUDA_OK(cudaMemsetAsync(d.h_at(0), 0, d.S() * sizeof(float), st));
  const int grid = 148 * 8;
  for (int t = 1; t <= d.T; ++t) {
    {
      ProfileScope ps(PTK_PROF_POINT_FWD, st);
      gru_stem_cat_kernel<<<grid, 256, 0, st>>>(d, t, 0);
    }
    if (wide) {
      if ((rc = gru_conv_wide<128>(d, d.Z, d.wk_ruF, d.RU, 2 * GC, 2 * GC, false, st))) return rc;
    } else if (tc) {
      if ((rc = gru_conv_tc(d, d.Z, d.RU, 0, 4, 4, 0, st))) return rc;
    } else {
      ConvGArgs cru{d.Z, d.wruT, d.bru, nullptr, d.RU, d.B, d.H, d.W, 2 * GC, 2 * GC, 0};
      if ((rc = launch_convg<7>(cru, st))) return rc;
    }
    {
      ProfileScope ps(PTK_PROF_POINT_FWD, st);
      gru_gate_kernel<<<grid, 256, 0, st>>>(d, t, tc);
    }
    if (wide) {
      if ((rc = gru_conv_wide<64>(d, d.Z2, d.wk_nF, d.Nb, 2 * GC, GC, false, st))) return rc;
    } else if (tc) {
      if ((rc = gru_conv_tc(d, d.Z2, d.Nb, 16, 2, 4, 0, st))) return rc;
    } else {
      ConvGArgs cn{d.Z2, d.wnT, d.bnw, nullptr, d.Nb, d.B, d.H, d.W, 2 * GC, GC, 0};
      if ((rc = launch_convg<7>(cn, st))) return rc;
    }
    {
      ProfileScope ps(PTK_PROF_POINT_FWD, st);
      gru_update_kernel<<<grid, 256, 0, st>>>(d, t, tc);
    }
    PTK_LAUNCH_OK("gru step");
  }
  const size_t smr = (size_t)2 * (d.H + 4) * (d.W + 4) * sizeof(float);
  PTK_CUDA_OK(cudaFuncSetAttribute(gru_readout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
  gru_readout_kernel<<<d.B, 256, smr, st>>>(d);
  PTK_LAUNCH_OK("gru_readout_kernel");
  return PTK_OK;
}

int run_gru_backward(const GruDev& d, cudaStream_t st) {
  int rc;
  const int taps = 49, grid = 148 * 8, P = d.H * d.W;
  const bool tc = d.conv_mode != PTK_CONV_FP32, wide = d.conv_mode == PTK_CONV_BF16X3;
  if ((rc = wide ? gru_prep_wide(d, true, st) : (tc ? gru_prep_tc(d, st) : gru_prep(d, true, st)))) return rc;
  {
    const size_t smr = (size_t)(2 * (d.H + 4) * (d.W + 4) + P) * sizeof(float);
    PTK_CUDA_OK(cudaFuncSetAttribute(gru_readout_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
    gru_readout_bwd_kernel<<<d.B, 256, smr, st>>>(d);       // dlogits -> d.dh (= dL/dh_T), head gradients, d goal channel
    PTK_LAUNCH_OK("gru_readout_bwd_kernel");
  }
  for (int t = d.T; t >= 1; --t) {
    const int first = (t == d.T);
    {
      ProfileScope ps(PTK_PROF_POINT_BWD, st, 2);
      gru_bwd1_kernel<<<grid, 256, 0, st>>>(d, t, first);
      gru_stem_cat_kernel<<<grid, 256, 0, st>>>(d, t, 1);            // Z = cat(X_t, h_{t-1}), Z2 = cat(X_t, r h)
    }
    if (wide) {
      if ((rc = gru_wgrad_wide(d, d.Z2, d.Nb, GC, 4, first, st))) return rc;
      if ((rc = gru_conv_wide<128>(d, d.Nb, d.wk_nD, d.dZ, GC, 2 * GC, true, st))) return rc;
    } else if (tc) {
      if ((rc = gru_wgrad_tc(d, d.Z2, d.Nb, 16, 2, first, st))) return rc;
      if ((rc = gru_conv_tc(d, d.Nb, d.dZ, 16, 2, 4, 1, st))) return rc;
    } else {
      WgradGArgs wn{d.Z2, d.Nb, d.slots_n, d.B, d.H, d.W, 2 * GC, GC, d.groups, !first};
      if ((rc = launch_wgradg<7>(wn, st))) return rc;
      ConvGArgs cn{d.Nb, d.wnD, nullptr, nullptr, d.dZ, d.B, d.H, d.W, GC, 2 * GC, 0};
      if ((rc = launch_convg<7>(cn, st))) return rc;
    }
    {
      ProfileScope ps(PTK_PROF_POINT_BWD, st);
      gru_bwd2_kernel<<<grid, 256, 0, st>>>(d, t, first);
    }
    if (wide) {
      if ((rc = gru_wgrad_wide(d, d.Z, d.RU, 2 * GC, 0, first, st))) return rc;
      if ((rc = gru_conv_wide<128>(d, d.RU, d.wk_ruD, d.dZ, 2 * GC, 2 * GC, true, st))) return rc;
    } else if (tc) {
      if ((rc = gru_wgrad_tc(d, d.Z, d.RU, 0, 4, first, st))) return rc;
      if ((rc = gru_conv_tc(d, d.RU, d.dZ, 0, 4, 4, 1, st))) return rc;
    } else {
      WgradGArgs wru{d.Z, d.RU, d.slots_ru, d.B, d.H, d.W, 2 * GC, 2 * GC, d.groups, !first};
      if ((rc = launch_wgradg<7>(wru, st))) return rc;
      ConvGArgs cru{d.RU, d.wruD, nullptr, nullptr, d.dZ, d.B, d.H, d.W, 2 * GC, 2 * GC, 0};
      if ((rc = launch_convg<7>(cru, st))) return rc;
    }
    {
      ProfileScope ps(PTK_PROF_POINT_BWD, st, 2);
      gru_bwd3_kernel<<<grid, 256, 0, st>>>(d, t);
      gru_stem_grad_kernel<<<GC * GRU_STEM_SPLIT, 512, 0, st>>>(d, t, first);
    }
    if (d.dx) gru_dx_kernel<<<grid, 256, 0, st>>>(d, t);
    PTK_LAUNCH_OK("gru backward step");
  }
  // biases: sums over (t, b, y, x) of the gate pre-activation gradients; weights: fold the per-group slots
  chan_sum_kernel<<<GC, 512, 0, st>>>(d.dNsum, d.bias_grad, d.B, GC, P);
  chan_sum_kernel<<<2 * GC, 512, 0, st>>>(d.dRUsum, d.bias_grad + GC, d.B, 2 * GC, P);
  if (wide) {
    if ((rc = gru_fold_wide(d, st))) return rc;
  } else if (tc) {
    gru_fold_tc_kernel<<<148 * 8, 256, 0, st>>>(d, conv_num_ctas(d.B, d.H, d.W));
  } else {
    fold_slots_kernel<<<(2 * GC * 2 * GC * taps + 255) / 256, 256, 0, st>>>(d.slots_ru, d.groups, 2 * GC, 2 * GC, taps,
                                                                           d.dparams + d.off_wr, d.off_wu - d.off_wr, GC);
    fold_slots_kernel<<<(GC * 2 * GC * taps + 255) / 256, 256, 0, st>>>(d.slots_n, d.groups, GC, 2 * GC, taps,
                                                                       d.dparams + d.off_wn, 0, GC);
  }
  PTK_CUDA_OK(cudaMemcpyAsync(d.dparams + d.off_bn, d.bias_grad, GC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PTK_CUDA_OK(cudaMemcpyAsync(d.dparams + d.off_br, d.bias_grad + GC, GC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PTK_CUDA_OK(cudaMemcpyAsync(d.dparams + d.off_bu, d.bias_grad + 2 * GC, GC * sizeof(float), cudaMemcpyDeviceToDevice, st));
  gru_stem_out_kernel<<<1, GC, 0, st>>>(d);
  gru_head_fold_kernel<<<1, 128, 0, st>>>(d);
  PTK_LAUNCH_OK("gru fold");
  return PTK_OK;
}

}  // namespace ptk
