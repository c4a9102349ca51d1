// Generic strict-fp32 direct convolution / weight-gradient kernels shared by the ConvLSTM (models/convlstm.py) and
// ConvGRU (models/kys.py) paths: arbitrary channel counts, kernel size 7 or 15 (template), 16x32 pixel tiles.
#pragma once
#include "ptk_common.cuh"
#include "ptk_profile.h"

namespace ptk {

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
  // per-oc pitch 16*32 + 4 floats: the 32 lanes of a warp read 32 different oc rows at the same pixel, and an
  // unpadded 2 KB pitch put all of them in one bank group (32-way conflict on the two LDS.128 per inner iteration)
  constexpr int DYP = GTH * GTW + 4;
  float* sDY = smem;                          // [32][DYP]
  float* sIn = smem + 32 * DYP;               // [ROWS][PITCH]
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
      cp_async16_zfill(sDY + o * DYP + r * GTW + 4 * q,
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
          const float4* p = reinterpret_cast<const float4*>(sDY + oc_l * DYP + y * GTW + xc * 8);
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
static __global__ void prepg_kernel(PrepGArgs p) {
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

// per-channel sums of a [B, Cn, P] tensor (bias gradients)
static __global__ void __launch_bounds__(512) chan_sum_kernel(const float* __restrict__ t, float* out, int B, int Cn, int P) {
  __shared__ double sh[16];
  const int c = blockIdx.x;
  const size_t N = (size_t)B * P;
  double s = 0;
  for (size_t i = threadIdx.x; i < N; i += blockDim.x) s += t[((i / P) * Cn + c) * P + i % P];
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) out[c] = (float)s;
}

// fold wgrad slots [groups][Cout][Cin][taps] -> torch layout of `nstack` stacked tensors inside dparams
static __global__ void fold_slots_kernel(const float* __restrict__ slots, int groups, int Cout, int Cin, int taps, float* dst0,
                                  long long stride_between, int cout_each) {
  const int total = Cout * Cin * taps;
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
  const size_t sm = (size_t)(32 * (GTH * GTW + 4) + G::ROWS * G::PITCH) * sizeof(float);
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

}  // namespace ptk
