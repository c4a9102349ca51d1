// Reverse-time (BPTT) loop of the InT / hGRU circuit.
//
// Replaces the autograd graph that loss.backward() (mainclean.py:204) walks through
// models/InT.py:223-235.  Per step the kernels below implement eqs (b1)-(b12) of SURVEY.md
// section 8 (restated and verified against autograd in oracle/int_oracle.py:backward_manual):
//
//   bwd_a   (b1)-(b2)  + BatchNorm1 backward sums          point-wise, recomputes A, Ge, G_e
//   bn_apply (b3)      dV -> dY1 in place                   element-wise
//   conv7   (b4)       dI' += dY1 *7^T W_exc                 (ptk_conv.cu, flipped weights)
//   wgrad7  (b4)       dW_exc += I' (star) dY1
//   bwd_c   (b5)-(b8)  + BatchNorm0 backward sums          point-wise, recomputes G_i, U, P, Q
//   bn_apply (b9)      dU -> dY0
//   conv7   (b9)       dGe += dY0 *7^T W_inh
//   wgrad7  (b9)       dW_inh += Ge (star) dY0
//   bwd_d   (b10)-(b12) attention gate + stem
//
// One thread = one pixel (32 channels in registers).  Sums over pixels (per-channel scalars and the
// 32x32 gate-weight outer products) are staged through shared memory per 128-pixel tile and
// accumulated per tile in global memory (read-add-write by the owning tile only => deterministic);
// `finalize` folds the tiles in double precision.
#include "ptk_common.cuh"
#include "ptk_int.h"
#include "ptk_profile.h"

namespace ptk {

constexpr int SP = PT + 4;   // staging pitch (floats): LDS.128-aligned rows, conflict-free

__device__ __forceinline__ void smem_fill_b(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
}

template <int ACT>
__device__ __forceinline__ void stem_pixel_pre(const FrameView& fv, int b, int t, int y, int x,
                                               const float* sWp, const float* sbp, float (&pre)[C],
                                               float (&xin)[3]) {
  xin[0] = fv.at(b, 0, t, y, x); xin[1] = fv.at(b, 1, t, y, x); xin[2] = fv.at(b, 2, t, y, x);
#pragma unroll
  for (int c = 0; c < C; ++c)
    pre[c] = fmaf(sWp[c * 3 + 2], xin[2], fmaf(sWp[c * 3 + 1], xin[1], fmaf(sWp[c * 3 + 0], xin[0], sbp[c])));
}

// sum over the tile's 128 pixels of a staged [32][SP] buffer: thread k reduces a quarter of row k/4;
// after the two shuffles every thread of the 4-group holds the row sum.
__device__ __forceinline__ float tile_rowsum(const float* sQ) {
  const int k = threadIdx.x, row = k >> 2, q = k & 3;
  const float4* p = reinterpret_cast<const float4*>(sQ + row * SP + q * 32);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 v = p[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return s;
}
__device__ __forceinline__ void acc_store(float* dst, float v, int first) {
  if (first) *dst = v; else *dst += v;
}
// dst[oc*32+ic] (+)= sum_px sG[oc][px] * sV[ic][px]
__device__ __forceinline__ void tile_outer(const float* sG, const float* sV, float* dst, int first) {
  const int k = threadIdx.x, oc = k >> 2, icg = k & 3;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(sG + oc * SP);
#pragma unroll 4
  for (int p = 0; p < PT / 4; ++p) {
    const float4 g = g4[p];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(sV + (icg + 4 * j) * SP + 4 * p);
      acc[j] = fmaf(g.x, v.x, fmaf(g.y, v.y, fmaf(g.z, v.z, fmaf(g.w, v.w, acc[j]))));
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc_store(dst + oc * C + icg + 4 * j, acc[j], first);
}
__device__ __forceinline__ void stage(float* sQ, const float (&v)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) sQ[c * SP + threadIdx.x] = v[c];
}

// step-local BatchNorm-backward sums: per-tile partials -> last CTA folds them (double) into
// bn_tot = (sum dV)/N, (sum dV*xhat)/N and accumulates d(bn bias), d(bn weight).
__device__ __forceinline__ void bn_sums_fold(const IntDev& d, int bn, float s0, float s1, int first,
                                             int* sIsLast) {
  const int k = threadIdx.x, row = k >> 2, q = k & 3;
  const int nTiles = gridDim.x;
  if (q == 0) {
    d.bn_sums[((size_t)blockIdx.x * 2 + 0) * C + row] = s0;
    d.bn_sums[((size_t)blockIdx.x * 2 + 1) * C + row] = s1;
  }
  __threadfence();
  __syncthreads();
  if (k == 0) *sIsLast = (atomicAdd(d.counter, 1u) == (unsigned)(nTiles - 1));
  __syncthreads();
  if (*sIsLast) {
    __threadfence();
    if (k < 2 * C) {
      double s = 0;
#pragma unroll 8
      for (int i = 0; i < nTiles; ++i) s += (double)__ldcg(d.bn_sums + (size_t)i * 2 * C + k);
      const double N = (double)d.B * d.H * d.W;
      d.bn_tot[k] = (float)(s / N);
      // k < C: sum dV -> d(bn bias); k >= C: sum dV*xhat -> d(bn weight)
      double* a = d.bn_acc + (size_t)bn * 2 * C + k;
      *a = first ? s : (*a + s);
    }
    if (k == 0) *d.counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// (b1)-(b2): needs dE', dI' (in gE, gI), E_t, I_t, I_{t+1}, Y1_t, BN1 stats, x_t.
// writes: A, Ge (scratch), geb = dG_e*G_e*(1-G_e), gE <- dE'*(1-G_e), gI += dR*V*kappa, dAct <- dV.
template <int CELL, int ACT>
__global__ void __launch_bounds__(PT) int_bwd_a_kernel(IntDev d, int t, int first) {
  extern __shared__ __align__(16) float dsm[];
  float* sWaw = dsm;              float* sWau = sWaw + C * C;
  float* sWew = sWau + C * C;     float* sWeu = sWew + C * C;
  float* sQ0 = sWeu + C * C;      float* sQ1 = sQ0 + C * SP;
  float* sQ2 = sQ1 + C * SP;      float* sQ3 = sQ2 + C * SP;
  __shared__ float sWp[C * 3], sbp[C], sba[C], sbe[C], sKa[C], sGa[C], sW1[C], sB1[C], sM1[C], sR1[C];
  __shared__ int sIsLast;
  smem_fill_b(sWaw, d.gateT[0], C * C);
  smem_fill_b(sWau, d.gateT[1], C * C);
  smem_fill_b(sWew, d.gateT[4], C * C);
  smem_fill_b(sWeu, d.gateT[5], C * C);
  smem_fill_b(sWp, d.Wp, C * 3);
  smem_fill_b(sbp, d.bp, C);
  smem_fill_b(sKa, d.kappa, C);
  smem_fill_b(sGa, d.gamma, C);
  smem_fill_b(sW1, d.bn_w[1], C);
  smem_fill_b(sB1, d.bn_b[1], C);
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    sba[c] = __ldg(d.gate_b[0] + c) + __ldg(d.gate_b[1] + c);
    sbe[c] = __ldg(d.gate_b[4] + c) + __ldg(d.gate_b[5] + c);
    const float* st = d.stat(t, 1);
    sM1[c] = st[c];
    sR1[c] = st[C + c];
  }
  __syncthreads();
  const int P = d.H * d.W, tilesPer = P / PT;
  const int b = blockIdx.x / tilesPer, pix = (blockIdx.x % tilesPer) * PT + threadIdx.x;
  const size_t base = (size_t)b * C * P + pix;
  const float* Et = d.ckE + d.ck_stride * t;
  const float* It = d.ckI + d.ck_stride * t;
  const float* I2t = d.ckI + d.ck_stride * (t + 1);
  const float* Y1 = d.Y1 + d.y_stride * t;

  float E[C], S[C], Ge[C], Gx[C];
  {
    float X[C], xin[3];
    stem_pixel_pre<ACT>(d.fv, b, t, pix / d.W, pix % d.W, sWp, sbp, X, xin);
#pragma unroll
    for (int c = 0; c < C; ++c) X[c] = actf<ACT>(X[c]);
#pragma unroll
    for (int c = 0; c < C; ++c) E[c] = Et[base + (size_t)c * P];
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = sba[c];
    matvec32(acc, sWaw, X);
    matvec32(acc, sWau, E);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float A = sigmoidf_(acc[c]);
      Ge[c] = A * E[c];
      d.A[base + (size_t)c * P] = A;           // backward always uses the [B,C,P] scratch
      d.Ge[base + (size_t)c * P] = Ge[c];
      S[c] = (CELL == PTK_CELL_INT) ? It[base + (size_t)c * P] : A;
    }
  }
  {
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = sbe[c];
    matvec32(acc, sWew, S);
    matvec32(acc, sWeu, Ge);
#pragma unroll
    for (int c = 0; c < C; ++c) Gx[c] = sigmoidf_(acc[c]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const size_t idx = base + (size_t)c * P;
    const float I2 = I2t[idx];
    const float xh = (Y1[idx] - sM1[c]) * sR1[c];
    const float V = fmaf(sW1[c], xh, sB1[c]);
    const float kIg = fmaf(sKa[c], I2, sGa[c]);
    const float R = V * kIg;
    const float Ehat = actf<ACT>(R);
    const float gEn = d.gE[idx];
    const float dGx = gEn * (Ehat - E[c]);
    d.gE[idx] = gEn * (1.0f - Gx[c]);
    const float dR = gEn * Gx[c] * dactf<ACT>(R);
    const float dV = dR * kIg;
    const float dRV = dR * V;
    d.gI[idx] = fmaf(dRV, sKa[c], first ? 0.0f : d.gI[idx]);
    d.geb[idx] = dGx * Gx[c] * (1.0f - Gx[c]);
    d.dAct[idx] = dV;
    sQ0[c * SP + threadIdx.x] = dRV * I2;   // -> d kappa
    sQ1[c * SP + threadIdx.x] = dRV;        // -> d gamma
    sQ2[c * SP + threadIdx.x] = dV;         // -> sum dV
    sQ3[c * SP + threadIdx.x] = dV * xh;    // -> sum dV * xhat
  }
  __syncthreads();
  const float r0 = tile_rowsum(sQ0), r1 = tile_rowsum(sQ1), r2 = tile_rowsum(sQ2), r3 = tile_rowsum(sQ3);
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if ((threadIdx.x & 3) == 0) {
    acc_store(acc + ACC_KAPPA + (threadIdx.x >> 2), r0, first);
    acc_store(acc + ACC_GAMMA + (threadIdx.x >> 2), r1, first);
  }
  bn_sums_fold(d, 1, r2, r3, first, &sIsLast);
}

// (b3)/(b9): dAct <- w*rstd*(dAct - mean(dAct) - xhat*mean(dAct*xhat)), xhat from the saved conv output.
__global__ void __launch_bounds__(256) bn_apply_kernel(IntDev d, const float* __restrict__ Y, int t, int bn) {
  const int P = d.H * d.W;
  const size_t n4 = (size_t)d.B * C * P / 4;
  const float* st = d.stat(t, bn);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4 / P) % C);
    const float mean = st[c], rstd = st[C + c];
    const float k = __ldg(d.bn_w[bn] + c) * rstd, m = d.bn_tot[c], n = d.bn_tot[C + c];
    float4 g = reinterpret_cast<float4*>(d.dAct)[i];
    const float4 y = reinterpret_cast<const float4*>(Y)[i];
    g.x = k * (g.x - m - (y.x - mean) * rstd * n);
    g.y = k * (g.y - m - (y.y - mean) * rstd * n);
    g.z = k * (g.z - m - (y.z - mean) * rstd * n);
    g.w = k * (g.w - m - (y.w - mean) * rstd * n);
    reinterpret_cast<float4*>(d.dAct)[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------
// (b5)-(b8): needs complete dI' (gI), geb, A/Ge (scratch), I_t, Y0_t, BN0 stats, x_t.
// writes: dAct <- dU, dGe, dX, gI <- dI (+ dS for InT), dS (hGRU), per-tile sums.
template <int CELL, int ACT>
__global__ void __launch_bounds__(PT) int_bwd_c_kernel(IntDev d, int t, int first) {
  extern __shared__ __align__(16) float dsm[];
  float* sWiwT = dsm;             float* sWiuT = sWiwT + C * C;   // [in][out] for the G_i recompute
  float* sWiw = sWiuT + C * C;    float* sWiu = sWiw + C * C;     // [out][in] for the transposed mat-vecs
  float* sWew = sWiu + C * C;     float* sWeu = sWew + C * C;
  float* sQ0 = sWeu + C * C;      float* sQ1 = sQ0 + C * SP;
  float* sQ2 = sQ1 + C * SP;      float* sQ3 = sQ2 + C * SP;
  __shared__ float sWp[C * 3], sbp[C], sbi[C], sAl[C], sMu[C], sW0[C], sB0[C], sM0[C], sR0[C];
  __shared__ int sIsLast;
  smem_fill_b(sWiwT, d.gateT[2], C * C);
  smem_fill_b(sWiuT, d.gateT[3], C * C);
  smem_fill_b(sWiw, d.gate[2], C * C);
  smem_fill_b(sWiu, d.gate[3], C * C);
  smem_fill_b(sWew, d.gate[4], C * C);
  smem_fill_b(sWeu, d.gate[5], C * C);
  smem_fill_b(sWp, d.Wp, C * 3);
  smem_fill_b(sbp, d.bp, C);
  smem_fill_b(sAl, d.alpha, C);
  smem_fill_b(sMu, d.mu, C);
  smem_fill_b(sW0, d.bn_w[0], C);
  smem_fill_b(sB0, d.bn_b[0], C);
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    sbi[c] = __ldg(d.gate_b[2] + c) + __ldg(d.gate_b[3] + c);
    const float* st = d.stat(t, 0);
    sM0[c] = st[c];
    sR0[c] = st[C + c];
  }
  __syncthreads();
  const int P = d.H * d.W, tilesPer = P / PT;
  const int b = blockIdx.x / tilesPer, pix = (blockIdx.x % tilesPer) * PT + threadIdx.x;
  const size_t base = (size_t)b * C * P + pix;
  const float* It = d.ckI + d.ck_stride * t;
  const float* Y0 = d.Y0 + d.y_stride * t;
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  const int k = threadIdx.x;

  float X[C], S[C], gi[C], dS[C];
  {
    float xin[3];
    stem_pixel_pre<ACT>(d.fv, b, t, pix / d.W, pix % d.W, sWp, sbp, X, xin);
#pragma unroll
    for (int c = 0; c < C; ++c) X[c] = actf<ACT>(X[c]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
    S[c] = (CELL == PTK_CELL_INT) ? It[base + (size_t)c * P] : d.A[base + (size_t)c * P];
  float z[C];
  {
    float Gi[C];
#pragma unroll
    for (int c = 0; c < C; ++c) Gi[c] = sbi[c];
    matvec32(Gi, sWiwT, X);
    matvec32(Gi, sWiuT, S);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const size_t idx = base + (size_t)c * P;
      const float G = sigmoidf_(Gi[c]);
      const float I = (CELL == PTK_CELL_INT) ? S[c] : It[idx];
      const float xh = (Y0[idx] - sM0[c]) * sR0[c];
      const float U = fmaf(sW0[c], xh, sB0[c]);
      const float aSm = fmaf(sAl[c], S[c], sMu[c]);
      const float Pp = U * aSm;
      const float Q = actf<ACT>(Pp);
      const float d1 = X[c] - Q;
      const float Ihat = actf<ACT>(d1);
      const float gIn = d.gI[idx];
      const float dGi = gIn * (Ihat - I);
      d.gI[idx] = gIn * (1.0f - G);                 // dI (InT adds dS below)
      const float zz = gIn * G * dactf<ACT>(d1);
      z[c] = zz;
      gi[c] = dGi * G * (1.0f - G);
      const float dP = -zz * dactf<ACT>(Pp);
      const float dPU = dP * U;
      const float dU = dP * aSm;
      dS[c] = dPU * sAl[c];
      d.dAct[idx] = dU;
      sQ0[c * SP + k] = dPU * S[c];   // -> d alpha
      sQ1[c * SP + k] = dPU;          // -> d mu
      sQ2[c * SP + k] = dU;           // -> sum dU
      sQ3[c * SP + k] = dU * xh;      // -> sum dU * xhat
    }
  }
  __syncthreads();
  {
    const float r0 = tile_rowsum(sQ0), r1 = tile_rowsum(sQ1), r2 = tile_rowsum(sQ2), r3 = tile_rowsum(sQ3);
    if ((k & 3) == 0) {
      acc_store(acc + ACC_ALPHA + (k >> 2), r0, first);
      acc_store(acc + ACC_MU + (k >> 2), r1, first);
    }
    bn_sums_fold(d, 0, r2, r3, first, &sIsLast);   // contains __syncthreads: staging buffers free again
  }
  // dX = z + Wi_w^T gi
  matvec32(z, sWiw, gi);
#pragma unroll
  for (int c = 0; c < C; ++c) d.dX[base + (size_t)c * P] = z[c];
  // dS += Wi_u^T gi + We_w^T ge ;  dGe = We_u^T ge
  matvec32(dS, sWiu, gi);
  float ge[C];
#pragma unroll
  for (int c = 0; c < C; ++c) ge[c] = d.geb[base + (size_t)c * P];
  matvec32(dS, sWew, ge);
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const size_t idx = base + (size_t)c * P;
    if (CELL == PTK_CELL_INT) d.gI[idx] += dS[c];   // (b10) InT: dI += dS
    else d.dS[idx] = dS[c];                          //       hGRU: dA += dS (in bwd_d)
  }
  {
    float dGe[C];
#pragma unroll
    for (int c = 0; c < C; ++c) dGe[c] = 0.f;
    matvec32(dGe, sWeu, ge);
#pragma unroll
    for (int c = 0; c < C; ++c) d.dGe[base + (size_t)c * P] = dGe[c];
  }
  // gate-weight outer products and bias sums
  stage(sQ0, ge);
  stage(sQ1, S);
  stage(sQ2, gi);
  stage(sQ3, X);
  __syncthreads();
  {
    const float rge = tile_rowsum(sQ0), rgi = tile_rowsum(sQ2);
    if ((k & 3) == 0) {
      acc_store(acc + ACC_BE + (k >> 2), rge, first);
      acc_store(acc + ACC_BI + (k >> 2), rgi, first);
    }
  }
  tile_outer(sQ0, sQ1, acc + ACC_WEW, first);   // ge (x) S
  tile_outer(sQ2, sQ3, acc + ACC_WIW, first);   // gi (x) X
  tile_outer(sQ2, sQ1, acc + ACC_WIU, first);   // gi (x) S
  __syncthreads();
#pragma unroll
  for (int c = 0; c < C; ++c) sQ1[c * SP + k] = d.Ge[base + (size_t)c * P];
  __syncthreads();
  tile_outer(sQ0, sQ1, acc + ACC_WEU, first);   // ge (x) Ge
}

// ------------------------------------------------------------------------------------------------
// (b10)-(b12): needs complete dGe, dE partial (gE), A, E_t, dX, [dS], x_t.
// writes gE <- dE (for step t-1), stem / attention-gate parameter sums, optional dx.
template <int CELL, int ACT>
__global__ void __launch_bounds__(PT) int_bwd_d_kernel(IntDev d, int t, int first) {
  extern __shared__ __align__(16) float dsm[];
  float* sWaw = dsm;              float* sWau = sWaw + C * C;   // [out][in]
  float* sQ0 = sWau + C * C;      float* sQ1 = sQ0 + C * SP;
  float* sQ2 = sQ1 + C * SP;
  __shared__ float sWp[C * 3], sbp[C], sXin[3][PT];
  smem_fill_b(sWaw, d.gate[0], C * C);
  smem_fill_b(sWau, d.gate[1], C * C);
  smem_fill_b(sWp, d.Wp, C * 3);
  smem_fill_b(sbp, d.bp, C);
  __syncthreads();
  const int P = d.H * d.W, tilesPer = P / PT;
  const int b = blockIdx.x / tilesPer, pix = (blockIdx.x % tilesPer) * PT + threadIdx.x;
  const int py = pix / d.W, px = pix % d.W;
  const size_t base = (size_t)b * C * P + pix;
  const float* Et = d.ckE + d.ck_stride * t;
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  const int k = threadIdx.x;

  float a[C], E[C], dE[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const size_t idx = base + (size_t)c * P;
    const float A = d.A[idx], dGe = d.dGe[idx];
    E[c] = Et[idx];
    float dA = dGe * E[c];
    if (CELL == PTK_CELL_HGRU) dA += d.dS[idx];
    a[c] = dA * A * (1.0f - A);
    dE[c] = fmaf(dGe, A, d.gE[idx]);              // dE partial, finished below
  }
  matvec32(dE, sWau, a);                           // dE += Wa_u^T a
#pragma unroll
  for (int c = 0; c < C; ++c) d.gE[base + (size_t)c * P] = dE[c];
  float dXv[C];
#pragma unroll
  for (int c = 0; c < C; ++c) dXv[c] = d.dX[base + (size_t)c * P];
  matvec32(dXv, sWaw, a);                          // dX += Wa_w^T a
  float pre[C], xin[3];
  stem_pixel_pre<ACT>(d.fv, b, t, py, px, sWp, sbp, pre, xin);
  // stage: a, X, E  -> outer products; then dpre
  stage(sQ0, a);
#pragma unroll
  for (int c = 0; c < C; ++c) sQ1[c * SP + k] = actf<ACT>(pre[c]);
  stage(sQ2, E);
  __syncthreads();
  {
    const float ra = tile_rowsum(sQ0);
    if ((k & 3) == 0) acc_store(acc + ACC_BA + (k >> 2), ra, first);
  }
  tile_outer(sQ0, sQ1, acc + ACC_WAW, first);   // a (x) X
  tile_outer(sQ0, sQ2, acc + ACC_WAU, first);   // a (x) E
  __syncthreads();
  float dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float dp = dXv[c] * dactf<ACT>(pre[c]);
    sQ0[c * SP + k] = dp;
    dx0 = fmaf(sWp[c * 3 + 0], dp, dx0);
    dx1 = fmaf(sWp[c * 3 + 1], dp, dx1);
    dx2 = fmaf(sWp[c * 3 + 2], dp, dx2);
  }
  sXin[0][k] = xin[0]; sXin[1][k] = xin[1]; sXin[2][k] = xin[2];
  if (d.dx != nullptr) {   // contiguous [B,3,T,H,W]
    const size_t plane = (size_t)d.T * P;
    float* o = d.dx + (size_t)b * 3 * plane + (size_t)t * P + pix;
    if (t == 0) dx2 += d.dgoal[(size_t)b * P + pix];
    o[0] = dx0; o[plane] = dx1; o[2 * plane] = dx2;
  }
  __syncthreads();
  {
    const float rp = tile_rowsum(sQ0);
    if ((k & 3) == 0) acc_store(acc + ACC_BP + (k >> 2), rp, first);
  }
  if (k < C * 3) {   // dWp[c][i] += sum_px dpre[c][px] * x[i][px]
    const int c = k / 3, i = k % 3;
    float s = 0.f;
#pragma unroll 8
    for (int p = 0; p < PT; ++p) s = fmaf(sQ0[c * SP + p], sXin[i][p], s);
    acc_store(acc + ACC_WP + k, s, first);
  }
}

// ------------------------------------------------------------------------------------------------
// Readout backward (reverse of models/InT.py:236-241); one CTA per sample.
// head partial layout per sample: [0..31] dWr, [32] dbr, [33..82] dWt, [83] dbt, [84] dWd, [85] dbd
constexpr int HD_WR = 0, HD_BR = 32, HD_WT = 33, HD_BT = 83, HD_WD = 84, HD_BD = 85, HD_N = 86;

__global__ void __launch_bounds__(256) int_readout_bwd_kernel(IntDev d, const float* __restrict__ E) {
  extern __shared__ float sm[];
  const int H = d.H, W = d.W, P = H * W, Wp2 = W + 4, Hp2 = H + 4;
  float* sR = sm;                 // zero-padded readout map r
  float* sG = sm + Hp2 * Wp2;     // zero-padded goal map
  float* sDr = sG + Hp2 * Wp2;    // [P] d r
  __shared__ float sWt[50], sWr[C], sRed[8], sPool;
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
  auto block_sum = [&](float v) -> float {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    __syncthreads();
    if ((tid & 31) == 0) sRed[tid >> 5] = v;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += sRed[i];
    return tot;
  };
  // pooled value (for d dense weight)
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
  const float pooled = block_sum(s) / (float)P + __ldg(d.bt);
  const float dl = d.dlogits[b];
  const float dpool = dl * __ldg(d.Wd);
  const float dov = dpool / (float)P;     // d(conv output), constant over the image
  float* hd = d.head + (size_t)b * HEAD_N;
  if (tid == 0) {
    hd[HD_WD] = dl * pooled;
    hd[HD_BD] = dl;
    hd[HD_BT] = dpool;
  }
  // d target_conv weight: window sums of the padded maps
  for (int w = 0; w < 50; ++w) {
    const int ch = w / 25, ky = (w % 25) / 5, kx = w % 5;
    const float* src = ch ? sG : sR;
    float v = 0.f;
    for (int pix = tid; pix < P; pix += 256) v += src[(pix / W + ky) * Wp2 + pix % W + kx];
    v = block_sum(v);
    if (tid == 0) hd[HD_WT + w] = dov * v;
  }
  // d r[y][x] = dov * sum over taps whose source pixel is inside the image
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
    d.dgoal[(size_t)b * P + pix] = dov * gsum;   // d x[b,2,0,y,x] via torch.cat(..., x[:,2,0]) (InT.py:236)
    sDr[pix] = dr;
    sbr += dr;
  }
  sbr = block_sum(sbr);
  if (tid == 0) hd[HD_BR] = sbr;
  __syncthreads();
  // dE = Wr * dr (initial dE' of the reverse loop); dWr[c] = sum dr * E[c]
  for (int c = 0; c < C; ++c) {
    float v = 0.f;
    for (int pix = tid; pix < P; pix += 256) {
      const size_t idx = ((size_t)b * C + c) * P + pix;
      v = fmaf(sDr[pix], E[idx], v);
      d.gE[idx] = sWr[c] * sDr[pix];
    }
    v = block_sum(v);
    if (tid == 0) hd[HD_WR + c] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: fold per-tile / per-CTA / per-sample partials (double) into the flat gradient vector.
struct FinalizeArgs {
  int nTiles, B, wgroups;
  int64_t off[PTK_INT_NPARAM + 1];
};

__global__ void __launch_bounds__(256) int_finalize_kernel(IntDev d, FinalizeArgs f) {
  const int64_t total = f.off[PTK_INT_NPARAM];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int p = 0;
    while (e >= f.off[p + 1]) ++p;
    const int j = (int)(e - f.off[p]);
    double s = 0;
    int accoff = -1, headoff = -1;
    switch (p) {
      case PTK_P_PREPROC_W: accoff = ACC_WP + j; break;
      case PTK_P_PREPROC_B: accoff = ACC_BP + j; break;
      case PTK_P_ALPHA: accoff = ACC_ALPHA + j; break;
      case PTK_P_MU: accoff = ACC_MU + j; break;
      case PTK_P_GAMMA: accoff = ACC_GAMMA + j; break;
      case PTK_P_KAPPA: accoff = ACC_KAPPA + j; break;
      case PTK_P_AW_W: accoff = ACC_WAW + j; break;
      case PTK_P_AU_W: accoff = ACC_WAU + j; break;
      case PTK_P_IW_W: accoff = ACC_WIW + j; break;
      case PTK_P_IU_W: accoff = ACC_WIU + j; break;
      case PTK_P_EW_W: accoff = ACC_WEW + j; break;
      case PTK_P_EU_W: accoff = ACC_WEU + j; break;
      case PTK_P_AW_B: case PTK_P_AU_B: accoff = ACC_BA + j; break;
      case PTK_P_IW_B: case PTK_P_IU_B: accoff = ACC_BI + j; break;
      case PTK_P_EW_B: case PTK_P_EU_B: accoff = ACC_BE + j; break;
      case PTK_P_READOUT_W: headoff = HD_WR + j; break;
      case PTK_P_READOUT_B: headoff = HD_BR; break;
      case PTK_P_TARGET_W: headoff = HD_WT + j; break;
      case PTK_P_TARGET_B: headoff = HD_BT; break;
      case PTK_P_DENSE_W: headoff = HD_WD; break;
      case PTK_P_DENSE_B: headoff = HD_BD; break;
      default: break;
    }
    if (accoff >= 0) {
      for (int i = 0; i < f.nTiles; ++i) s += (double)d.acc[(size_t)i * ACC_N + accoff];
    } else if (headoff >= 0) {
      for (int i = 0; i < f.B; ++i) s += (double)d.head[(size_t)i * HEAD_N + headoff];
    } else if (p == PTK_P_W_EXC || p == PTK_P_W_INH) {
      // dW[oc][ic][tap] = sum_groups wacc[(grp*4 + ic/8)][tap][oc*8 + ic%8]
      const float* wa = (p == PTK_P_W_EXC) ? d.wacc_exc : d.wacc_inh;
      const int oc = j / (C * TAPS), ic = (j / TAPS) % C, tap = j % TAPS;
      for (int g = 0; g < f.wgroups; ++g)
        s += (double)wa[((size_t)(g * 4 + ic / 8) * TAPS + tap) * 256 + oc * 8 + (ic & 7)];
    } else if (p == PTK_P_BN0_W) s = d.bn_acc[0 * 2 * C + C + j];
    else if (p == PTK_P_BN0_B) s = d.bn_acc[0 * 2 * C + j];
    else if (p == PTK_P_BN1_W) s = d.bn_acc[1 * 2 * C + C + j];
    else if (p == PTK_P_BN1_B) s = d.bn_acc[1 * 2 * C + j];
    d.dparams[e] = (float)s;
  }
}

// ---- host ----------------------------------------------------------------------------------
static constexpr size_t SM_BWD_A = (size_t)(4 * C * C + 4 * C * SP) * sizeof(float);
static constexpr size_t SM_BWD_C = (size_t)(6 * C * C + 4 * C * SP) * sizeof(float);
static constexpr size_t SM_BWD_D = (size_t)(2 * C * C + 3 * C * SP) * sizeof(float);

template <int CELL, int ACT>
static int run_backward_t(const IntDev& d, cudaStream_t st) {
  const int P = d.H * d.W;
  const int tiles = d.B * P / PT;
  PTK_CUDA_OK(cudaFuncSetAttribute(int_bwd_a_kernel<CELL, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BWD_A));
  PTK_CUDA_OK(cudaFuncSetAttribute(int_bwd_c_kernel<CELL, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BWD_C));
  PTK_CUDA_OK(cudaFuncSetAttribute(int_bwd_d_kernel<CELL, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BWD_D));
  const size_t smr = (size_t)(2 * (d.H + 4) * (d.W + 4) + P) * sizeof(float);
  PTK_CUDA_OK(cudaFuncSetAttribute(int_readout_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
  { ProfileScope ps(PTK_PROF_MISC, st);
  int_readout_bwd_kernel<<<d.B, 256, smr, st>>>(d, d.ckE + d.ck_stride * d.T); }
  PTK_LAUNCH_OK("int_readout_bwd_kernel");
  const int ew_blocks = 148 * 4;
  for (int t = d.T - 1; t >= 0; --t) {
    const int first = (t == d.T - 1);
    int rc;
    { ProfileScope ps(PTK_PROF_POINT_BWD, st);
    int_bwd_a_kernel<CELL, ACT><<<tiles, PT, SM_BWD_A, st>>>(d, t, first); }
    PTK_LAUNCH_OK("int_bwd_a_kernel");
    { ProfileScope ps(PTK_PROF_BN_APPLY, st);
    bn_apply_kernel<<<ew_blocks, 256, 0, st>>>(d, d.Y1 + d.y_stride * t, t, 1); }
    PTK_LAUNCH_OK("bn_apply_kernel");
    ConvArgs c1{d.dAct, d.w_exc_d, d.gI, d.B, d.H, d.W, 1, nullptr, nullptr, nullptr};
    { ProfileScope ps(PTK_PROF_CONV_DGRAD, st); rc = launch_conv7(c1, st); }
    if (rc) return rc;
    WgradArgs w1{d.ckI + d.ck_stride * (t + 1), d.dAct, d.wacc_exc, d.B, d.H, d.W, first};
    { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgrad7(w1, WGRAD_GROUPS, st); }
    if (rc) return rc;
    { ProfileScope ps(PTK_PROF_POINT_BWD, st);
    int_bwd_c_kernel<CELL, ACT><<<tiles, PT, SM_BWD_C, st>>>(d, t, first); }
    PTK_LAUNCH_OK("int_bwd_c_kernel");
    { ProfileScope ps(PTK_PROF_BN_APPLY, st);
    bn_apply_kernel<<<ew_blocks, 256, 0, st>>>(d, d.Y0 + d.y_stride * t, t, 0); }
    PTK_LAUNCH_OK("bn_apply_kernel");
    ConvArgs c0{d.dAct, d.w_inh_d, d.dGe, d.B, d.H, d.W, 1, nullptr, nullptr, nullptr};
    { ProfileScope ps(PTK_PROF_CONV_DGRAD, st); rc = launch_conv7(c0, st); }
    if (rc) return rc;
    WgradArgs w0{d.Ge, d.dAct, d.wacc_inh, d.B, d.H, d.W, first};
    { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgrad7(w0, WGRAD_GROUPS, st); }
    if (rc) return rc;
    { ProfileScope ps(PTK_PROF_POINT_BWD, st);
    int_bwd_d_kernel<CELL, ACT><<<tiles, PT, SM_BWD_D, st>>>(d, t, first); }
    PTK_LAUNCH_OK("int_bwd_d_kernel");
  }
  FinalizeArgs f;
  f.nTiles = tiles; f.B = d.B; f.wgroups = WGRAD_GROUPS;
  const ParamOffsets po = param_offsets(C, KS);
  for (int i = 0; i <= PTK_INT_NPARAM; ++i) f.off[i] = po.o[i];
  { ProfileScope ps(PTK_PROF_MISC, st);
  int_finalize_kernel<<<148 * 2, 256, 0, st>>>(d, f); }
  PTK_LAUNCH_OK("int_finalize_kernel");
  return PTK_OK;
}

int run_int_backward(const IntDev& d, int cell, int act, cudaStream_t st) {
  if (cell == PTK_CELL_INT) {
    if (act == PTK_ACT_SOFTPLUS) return run_backward_t<PTK_CELL_INT, PTK_ACT_SOFTPLUS>(d, st);
    return run_backward_t<PTK_CELL_INT, PTK_ACT_TANH>(d, st);
  }
  if (act == PTK_ACT_SOFTPLUS) return run_backward_t<PTK_CELL_HGRU, PTK_ACT_SOFTPLUS>(d, st);
  return run_backward_t<PTK_CELL_HGRU, PTK_ACT_TANH>(d, st);
}

}  // namespace ptk
