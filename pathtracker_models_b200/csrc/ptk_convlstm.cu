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
#include "ptk_convlstm.h"
#include "ptk_profile.h"

namespace ptk {

constexpr int LC = 25;            // hidden channels of the ConvLSTM (models/convlstm.py:108)
constexpr int LG = 4 * LC;        // stacked gate channels
constexpr int GTH = 16, GTW = 32; // pixel tile

template <int KS> struct ConvG {
  static constexpr int PAD = KS / 2;
  static constexpr int PADL = (PAD + 3) & ~3;          // interior starts at a 16-byte aligned column
  static constexpr int OFF = PADL - PAD;               // padded column of input x + kx - PAD is x + kx + OFF
  static constexpr int NLD = ((8 + KS - 1 + OFF) + 3) & ~3;   // floats a thread loads per input row
  static constexpr int ROWS = GTH + KS - 1;
  static constexpr int PITCH = (PADL + GTW + PAD + 3 + 4) & ~3 | 4;   // >= PADL + 32 + PAD, multiple of 4, = 4 mod 8
  static constexpr int ICC = (KS > 7) ? 4 : 8;         // input channels per staged chunk
  static constexpr int W_IC = KS * KS * 32;            // floats of weights per input channel (one 32-oc block)
  static constexpr size_t SMEM = (size_t)ICC * (ROWS * PITCH + W_IC) * sizeof(float);
};

struct ConvGArgs {
  const float* in;      // [B, Cin, H, W]
  const float* wT;      // [ocBlocks][Cin][KS*KS][32]
  const float* bias;    // [Cout] or null
  const float* addend;  // [B, Cout, H, W] or null
  float* out;           // [B, Cout, H, W]
  int B, H, W, Cin, Cout, accumulate;
  const float* skip = nullptr;   // RBP: device state, skip[1] != 0 => the Neumann series has terminated, do nothing
};

template <int KS>
__global__ void __launch_bounds__(256, 1) convg_kernel(ConvGArgs a) {
  using G = ConvG<KS>;
  if (a.skip != nullptr && a.skip[1] != 0.f) return;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;                                   // [ICC][ROWS][PITCH]
  float* sW = smem + G::ICC * G::ROWS * G::PITCH;      // [ICC][KS*KS][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cg = lane & 3, rl = lane >> 2, ocg = warp & 3, rh = warp >> 2, ty = rh * 8 + rl;
  const int ocBlocks = (a.Cout + 31) / 32;
  const int b = blockIdx.z / ocBlocks, ob = blockIdx.z % ocBlocks;
  const int x0 = blockIdx.x * GTW, y0 = blockIdx.y * GTH, H = a.H, W = a.W;
  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[j][o] = 0.f;

  for (int c0 = 0; c0 < a.Cin; c0 += G::ICC) {
    const int nic = (a.Cin - c0) < G::ICC ? (a.Cin - c0) : G::ICC;
    __syncthreads();
    // input chunk with zero-filled halo (scalar cp.async: the halo breaks 16-byte alignment)
    for (int i = tid; i < nic * G::ROWS * (GTW + 2 * G::PAD); i += 256) {
      const int icl = i / (G::ROWS * (GTW + 2 * G::PAD)), rem = i - icl * (G::ROWS * (GTW + 2 * G::PAD));
      const int r = rem / (GTW + 2 * G::PAD), c = rem - r * (GTW + 2 * G::PAD);
      const int gy = y0 + r - G::PAD, gx = x0 + c - G::PAD;
      const bool v = gy >= 0 && gy < H && gx >= 0 && gx < W;
      cp_async4_zfill(sIn + (icl * G::ROWS + r) * G::PITCH + c + G::OFF,
                      a.in + ((size_t)(b * a.Cin + c0 + icl) * H + (v ? gy : 0)) * W + (v ? gx : 0), v);
    }
    const float* wsrc = a.wT + ((size_t)ob * a.Cin + c0) * G::W_IC;
    for (int i = tid; i < nic * G::W_IC / 4; i += 256) cp_async16(sW + 4 * i, wsrc + 4 * i);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll 1
    for (int ic = 0; ic < nic; ++ic) {
      const float* sI = sIn + (ic * G::ROWS + ty) * G::PITCH + cg * 8;
      const float* sWi = sW + ic * G::W_IC + ocg * 8;
#pragma unroll 1
      for (int ky = 0; ky < KS; ++ky) {
        float in[G::NLD];
#pragma unroll
        for (int q = 0; q < G::NLD / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(sI + ky * G::PITCH + 4 * q);
          in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(sWi + (ky * KS + kx) * 32);
          const float4 w0 = wp[0], w1 = wp[1];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = in[j + kx + G::OFF];
            acc[j][0] = fmaf(v, w0.x, acc[j][0]); acc[j][1] = fmaf(v, w0.y, acc[j][1]);
            acc[j][2] = fmaf(v, w0.z, acc[j][2]); acc[j][3] = fmaf(v, w0.w, acc[j][3]);
            acc[j][4] = fmaf(v, w1.x, acc[j][4]); acc[j][5] = fmaf(v, w1.y, acc[j][5]);
            acc[j][6] = fmaf(v, w1.z, acc[j][6]); acc[j][7] = fmaf(v, w1.w, acc[j][7]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const int oc = ob * 32 + ocg * 8 + o;
    if (oc >= a.Cout) continue;
    const size_t off = ((size_t)(b * a.Cout + oc) * H + y0 + ty) * W + x0 + cg * 8;
    const float bb = a.bias ? __ldg(a.bias + oc) : 0.f;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[j][o] + bb;
    if (a.addend) {
      const float4 p0 = *reinterpret_cast<const float4*>(a.addend + off), p1 = *reinterpret_cast<const float4*>(a.addend + off + 4);
      v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
    }
    if (a.accumulate) {
      const float4 p0 = *reinterpret_cast<const float4*>(a.out + off), p1 = *reinterpret_cast<const float4*>(a.out + off + 4);
      v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
    }
    *reinterpret_cast<float4*>(a.out + off) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(a.out + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---- weight gradient: dW[oc][ic][ky][kx] = sum_{b,y,x} dY[b,oc,y,x] In[b,ic,y+ky-p,x+kx-p] ------------------
// CTA = (32-oc block, one ic, tile group); thread = (oc, ky) with the KS kx-accumulators in registers
// across the CTA's tiles; slot layout [group][Cout][Cin][KS*KS] (written, never read-modified).
struct WgradGArgs {
  const float* in;    // [B, Cin, H, W]
  const float* dy;    // [B, Cout, H, W]
  float* slots;       // [groups][Cout][Cin][KS*KS]
  int B, H, W, Cin, Cout, groups, accumulate;
};

template <int KS>
__global__ void __launch_bounds__(32 * KS, 2) wgradg_kernel(WgradGArgs a) {
  using G = ConvG<KS>;
  extern __shared__ __align__(16) float smem[];
  float* sDY = smem;                          // [32][16*32]
  float* sIn = smem + 32 * GTH * GTW;         // [ROWS][PITCH]
  const int tid = threadIdx.x, NT = 32 * KS;
  const int oc_l = tid % 32, ky = tid / 32;
  const int ocBlocks = (a.Cout + 31) / 32;
  const int ob = blockIdx.x % ocBlocks, ic = blockIdx.x / ocBlocks, grp = blockIdx.y;
  const int H = a.H, W = a.W, tilesX = W / GTW, tilesY = H / GTH, nTiles = a.B * tilesX * tilesY;
  float acc[KS];
#pragma unroll
  for (int i = 0; i < KS; ++i) acc[i] = 0.f;
  for (int tile = grp; tile < nTiles; tile += a.groups) {
    const int b = tile / (tilesX * tilesY), rem = tile % (tilesX * tilesY);
    const int y0 = (rem / tilesX) * GTH, x0 = (rem % tilesX) * GTW;
    __syncthreads();
    for (int i = tid; i < 32 * GTH * (GTW / 4); i += NT) {
      const int o = i / (GTH * GTW / 4), r = (i / (GTW / 4)) % GTH, q = i % (GTW / 4);
      const int oc = ob * 32 + o;
      cp_async16_zfill(sDY + (o * GTH + r) * GTW + 4 * q,
                       a.dy + ((size_t)(b * a.Cout + (oc < a.Cout ? oc : 0)) * H + y0 + r) * W + x0 + 4 * q, oc < a.Cout);
    }
    for (int i = tid; i < G::ROWS * (GTW + 2 * G::PAD); i += NT) {
      const int r = i / (GTW + 2 * G::PAD), c = i - r * (GTW + 2 * G::PAD);
      const int gy = y0 + r - G::PAD, gx = x0 + c - G::PAD;
      const bool v = gy >= 0 && gy < H && gx >= 0 && gx < W;
      cp_async4_zfill(sIn + r * G::PITCH + c + G::OFF, a.in + ((size_t)(b * a.Cin + ic) * H + (v ? gy : 0)) * W + (v ? gx : 0), v);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll 1
    for (int y = 0; y < GTH; ++y) {
#pragma unroll 1
      for (int xc = 0; xc < GTW / 8; ++xc) {
        float dy[8], in[G::NLD];
        {
          const float4* p = reinterpret_cast<const float4*>(sDY + (oc_l * GTH + y) * GTW + xc * 8);
          const float4 v0 = p[0], v1 = p[1];
          dy[0] = v0.x; dy[1] = v0.y; dy[2] = v0.z; dy[3] = v0.w; dy[4] = v1.x; dy[5] = v1.y; dy[6] = v1.z; dy[7] = v1.w;
        }
#pragma unroll
        for (int q = 0; q < G::NLD / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(sIn + (y + ky) * G::PITCH + xc * 8 + 4 * q);
          in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[kx] = fmaf(dy[j], in[j + kx + G::OFF], acc[kx]);
      }
    }
  }
  const int oc = ob * 32 + oc_l;
  if (oc < a.Cout) {
    float* dst = a.slots + (((size_t)grp * a.Cout + oc) * a.Cin + ic) * KS * KS + ky * KS;
#pragma unroll
    for (int kx = 0; kx < KS; ++kx) dst[kx] = a.accumulate ? dst[kx] + acc[kx] : acc[kx];
  }
}

// ---- weight re-layout: torch [Cout][Cin][KS][KS] (possibly 4 stacked gate tensors) -> [ocBlock][Cin][tap][32]
//      fwd: as is.  transpose (conv_transpose2d): roles of in/out channels swapped, taps flipped.
struct PrepGArgs {
  const float* w[4];   // up to 4 stacked tensors, each [cout_each][cin][KS*KS]
  int nstack, cout_each, cin, taps, transpose;
  float* dst;
};
__global__ void prepg_kernel(PrepGArgs p) {
  const int Cout = p.nstack * p.cout_each;
  const int total = Cout * p.cin * p.taps;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int oc = i / (p.cin * p.taps), ic = (i / p.taps) % p.cin, tap = i % p.taps;
  const float v = p.w[oc / p.cout_each][((size_t)(oc % p.cout_each) * p.cin + ic) * p.taps + tap];
  if (!p.transpose) {
    p.dst[(((size_t)(oc / 32) * p.cin + ic) * p.taps + tap) * 32 + (oc & 31)] = v;
  } else {   // "input" channel = oc, "output" channel = ic, tap -> taps-1-tap
    p.dst[(((size_t)(ic / 32) * Cout + oc) * p.taps + (p.taps - 1 - tap)) * 32 + (ic & 31)] = v;
  }
}

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

// ---- head: BatchNorm (batch statistics) + 1x1 conv 25 -> 2, forward and backward; one CTA per channel ----------
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}
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
// per-channel sums of a [B, Cn, P] tensor (bias gradients)
__global__ void __launch_bounds__(512) chan_sum_kernel(const float* __restrict__ t, float* out, int B, int Cn, int P) {
  __shared__ double sh[16];
  const int c = blockIdx.x;
  const size_t N = (size_t)B * P;
  double s = 0;
  for (size_t i = threadIdx.x; i < N; i += blockDim.x) s += t[((i / P) * Cn + c) * P + i % P];
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) out[c] = (float)s;
}
// fold wgrad slots [groups][Cout][Cin][taps] -> torch layout of `nstack` stacked tensors inside dparams
__global__ void fold_slots_kernel(const float* __restrict__ slots, int groups, int Cout, int Cin, int taps, float* dst0,
                                  long long stride_between) {
  const int total = Cout * Cin * taps, cout_each = LC;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double s = 0;
  for (int g = 0; g < groups; ++g) s += slots[(size_t)g * total + i];
  const int oc = i / (Cin * taps), rest = i % (Cin * taps);
  dst0[(size_t)(oc / cout_each) * stride_between + (size_t)(oc % cout_each) * Cin * taps + rest] = (float)s;
}

// ---- host side ------------------------------------------------------------------------------------------------
template <int KS>
static int launch_convg(const ConvGArgs& a, cudaStream_t st) {
  using G = ConvG<KS>;
  PTK_CUDA_OK(cudaFuncSetAttribute(convg_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
  dim3 grid(a.W / GTW, a.H / GTH, a.B * ((a.Cout + 31) / 32));
  ProfileScope ps(PTK_PROF_CONV_FWD, st);
  convg_kernel<KS><<<grid, 256, G::SMEM, st>>>(a);
  PTK_LAUNCH_OK("convg_kernel");
  return PTK_OK;
}
template <int KS>
static int launch_wgradg(const WgradGArgs& a, cudaStream_t st) {
  using G = ConvG<KS>;
  const size_t sm = (size_t)(32 * GTH * GTW + G::ROWS * G::PITCH) * sizeof(float);
  PTK_CUDA_OK(cudaFuncSetAttribute(wgradg_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  dim3 grid(((a.Cout + 31) / 32) * a.Cin, a.groups);
  ProfileScope ps(PTK_PROF_WGRAD, st);
  wgradg_kernel<KS><<<grid, 32 * KS, sm, st>>>(a);
  PTK_LAUNCH_OK("wgradg_kernel");
  return PTK_OK;
}
static int launch_prepg(const float* const* w, int nstack, int cout_each, int cin, int taps, int transpose, float* dst,
                        size_t dst_floats, cudaStream_t st) {
  PTK_CUDA_OK(cudaMemsetAsync(dst, 0, dst_floats * sizeof(float), st));
  PrepGArgs p{};
  for (int i = 0; i < nstack; ++i) p.w[i] = w[i];
  p.nstack = nstack; p.cout_each = cout_each; p.cin = cin; p.taps = taps; p.transpose = transpose; p.dst = dst;
  const int total = nstack * cout_each * cin * taps;
  prepg_kernel<<<(total + 255) / 256, 256, 0, st>>>(p);
  PTK_LAUNCH_OK("prepg_kernel");
  return PTK_OK;
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
  ConvGArgs c0{d.x, d.w0T, d.b0, nullptr, d.u, d.B, d.H, d.W, 1, LC, 0};
  if ((rc = launch_convg<7>(c0, st))) return rc;
  square_kernel<<<148 * 4, 256, 0, st>>>(d.u, d.xt, S);
  ConvGArgs cx{d.xt, d.wxT, d.bx, nullptr, d.Xg, d.B, d.H, d.W, LC, LG, 0};
  if ((rc = launch_convg<KS>(cx, st))) return rc;
  PTK_CUDA_OK(cudaMemsetAsync(d.h_at(0), 0, S * sizeof(float), st));
  PTK_CUDA_OK(cudaMemsetAsync(d.c_at(0), 0, S * sizeof(float), st));
  for (int t = 1; t <= d.T; ++t) {
    ConvGArgs ch{d.h_at(t - 1), d.whT, nullptr, d.Xg, d.G, d.B, d.H, d.W, LC, LG, 0};
    if ((rc = launch_convg<KS>(ch, st))) return rc;
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
  const int T = d.T;
  // seed dh = 1, dc = 0  -> jh
  lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(nullptr, 1.0f, nullptr, 0.0f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                           nullptr, nullptr, 0, d.B, P);
  ConvGArgs c1{d.delta, d.whD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0};
  if ((rc = launch_convg<KS>(c1, st))) return rc;
  // seed dh = 0, dc = 1  -> v = d c_T / d h_{T-1}
  lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(nullptr, 0.0f, nullptr, 1.0f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                           nullptr, nullptr, 0, d.B, P);
  ConvGArgs c2{d.delta, d.whD, nullptr, nullptr, d.dc, d.B, d.H, d.W, LG, LC, 0};
  if ((rc = launch_convg<KS>(c2, st))) return rc;
  penalty_kernel<<<148 * 4, 256, 0, st>>>(d.dh, d.dc, d.gates_at(T), T >= 2 ? d.gates_at(T - 1) : nullptr,
                                          d.c_at(T - 1), d.penalty, d.B, P);
  PTK_LAUNCH_OK("penalty_kernel");
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
  if ((rc = launch_prepg(d.wx, 4, LC, LC, taps, 1, d.wxD, (size_t)LG * taps * 32, st))) return rc;
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
      PTK_LAUNCH_OK("lstm_bwd_kernel");
    }
    WgradGArgs w{d.h_at(t - 1), d.delta, d.slots_h, d.B, d.H, d.W, LC, LG, d.groups, !first};
    if ((rc = launch_wgradg<KS>(w, st))) return rc;
    if (t > 1) {
      ConvGArgs c{d.delta, d.whD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0};
      if ((rc = launch_convg<KS>(c, st))) return rc;
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
  if ((rc = launch_wgradg<KS>(wx, st))) return rc;
  chan_sum_kernel<<<LG, 512, 0, st>>>(d.delta_sum, d.bx_grad, d.B, LG, P);
  ConvGArgs cx{d.delta_sum, d.wxD, nullptr, nullptr, d.dh, d.B, d.H, d.W, LG, LC, 0};   // dh reused as d xt
  if ((rc = launch_convg<KS>(cx, st))) return rc;
  dsquare_kernel<<<148 * 4, 256, 0, st>>>(d.u, d.dh, S);                                 // d u = 2 u d xt
  WgradGArgs w0{d.x, d.dh, d.slots_0, d.B, d.H, d.W, 1, LC, d.groups, 0};
  if ((rc = launch_wgradg<7>(w0, st))) return rc;
  chan_sum_kernel<<<LC, 512, 0, st>>>(d.dh, d.b0_grad, d.B, LC, P);
  // fold everything into the flat gradient vector (layout: ptk_convlstm_param_offsets)
  const long long per_gate = (long long)LC * LC * taps + LC + (long long)LC * LC * taps;   // Wx.w, Wx.b, Wh.w
  fold_slots_kernel<<<(LC * 49 + 255) / 256, 256, 0, st>>>(d.slots_0, d.groups, LC, 1, 49, d.dparams + d.off_w0, 0);
  fold_slots_kernel<<<(LG * LC * taps + 255) / 256, 256, 0, st>>>(d.slots_x, d.groups, LG, LC, taps, d.dparams + d.off_wx, per_gate);
  fold_slots_kernel<<<(LG * LC * taps + 255) / 256, 256, 0, st>>>(d.slots_h, d.groups, LG, LC, taps, d.dparams + d.off_wh, per_gate);
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
  if ((rc = launch_prepg(d.wx, 4, LC, LC, taps, 1, d.wxD, (size_t)LG * taps * 32, st))) return rc;
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
    if ((rc = launch_convg<KS>(c, st))) return rc;
    rbp_sumsq_kernel<<<RBP_CTAS, 256, 0, st>>>(d.dh, d.rbp_g, S, d.rbp_part, d.rbp_state);
    rbp_decide_kernel<<<1, 32, 0, st>>>(d.rbp_part, RBP_CTAS, d.rbp_state, 0);
    rbp_commit_kernel<<<148 * 4, 256, 0, st>>>(d.rbp_g, d.dc, d.dh, S, d.rbp_state);
  }
  PTK_LAUNCH_OK("rbp neumann series");
  // one cell-step backward with dh_T := g (dc_T = 0): deltas of the last step only
  lstm_bwd_kernel<<<148 * 4, 256, 0, st>>>(d.rbp_g, 0.f, nullptr, 0.f, d.gates_at(T), d.c_at(T - 1), d.c_at(T), d.delta,
                                           nullptr, d.delta_sum, 1, d.B, P);
  WgradGArgs w{d.h_at(T - 1), d.delta, d.slots_h, d.B, d.H, d.W, LC, LG, d.groups, 0};
  if ((rc = launch_wgradg<KS>(w, st))) return rc;
  return convlstm_grad_tail<KS>(d, st);
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
