// Reverse-time (BPTT) loop of the InT / hGRU circuit.
//
// Replaces the autograd graph that loss.backward() (mainclean.py:204) walks through
// models/InT.py:223-235.  Per step the kernels below implement eqs (b1)-(b12) of SURVEY.md
// section 8 (restated and verified against autograd in oracle/int_oracle.py:backward_manual):
//
//   bwd_a     (b1)-(b2)   + BatchNorm1 backward sums     element-wise (gates A, G_e come from the checkpoint)
//   bn_apply  (b3)        dV -> dY1 in place
//   conv7     (b4)        dY1 *7^T W_exc -> cgI (added to dI' by bwd_c; ptk_conv*.cu, flipped weights)
//   wgrad7    (b4)        dW_exc += I' (star) dY1
//   bwd_c     (b5)-(b8)   + BatchNorm0 backward sums     element-wise + 4 transposed 1x1 products
//   bn_apply  (b9)        dU -> dY0
//   conv7     (b9)        dY0 *7^T W_inh -> cGe (added to dGe by bwd_d)
//   wgrad7    (b9)        dW_inh += Ge (star) dY0
//   bwd_d     (b10)-(b12) attention gate + stem           element-wise + 2 transposed 1x1 products
//   gate_wgrad            the six 32x32 gate-weight gradients + gate-bias gradients of the step
//
// Point-wise kernels use the warp-tile scheme of ptk_pw.cuh.  Sums over pixels live in registers
// across a persistent CTA's tiles and are accumulated per CTA slot in global memory (read-add-write
// by the owning CTA only => deterministic); `finalize` folds the slots in double precision.
#include <stdlib.h>

#include "ptk_int.h"
#include "ptk_profile.h"
#include "ptk_pw.cuh"
#include "ptk_tc.cuh"

#ifndef PTK_BWD_MMA
#define PTK_BWD_MMA 0
#endif
#ifndef PTK_BWD_D_MMA
#define PTK_BWD_D_MMA 1
#endif

namespace ptk {

constexpr int TILE_F = C * TP;
constexpr int CELL_NOINH = 2;   // internal: InT(no_inh=True), models/InT.py:167-168

__device__ __forceinline__ void acc_store(float* dst, float v, int first) {
  if (first) *dst = v; else *dst += v;
}

// step-local BatchNorm-backward sums: per-CTA partials -> last CTA folds them (double) into
// bn_tot = (sum dV)/N, (sum dV*xhat)/N and accumulates d(bn bias), d(bn weight).
__device__ __forceinline__ void bn_sums_fold(const IntDev& d, int bn, const float* sPart /*[2][32]*/, int first,
                                             int* sIsLast, float* sScratch /* >= 2 KB, 8-byte aligned, free */) {
  const int k = threadIdx.x, nCta = gridDim.x;
  if (k < 2 * C) d.bn_sums[(size_t)blockIdx.x * 2 * C + k] = sPart[k];
  __threadfence();
  __syncthreads();
  if (k == 0) *sIsLast = (atomicAdd(d.counter, 1u) == (unsigned)(nCta - 1));
  __syncthreads();
  if (*sIsLast) {
    // The whole grid waits for this CTA, so the fold is latency-critical: 128 threads, 16 independent
    // 16-byte loads in flight per thread (a dependent chain of nCta L2 round trips used to cost ~25 us).
    // thread (q, g): float4 q of the 64 sums, rows g, g+8, g+16, ...; fixed order => deterministic.
    __threadfence();
    const int q = k & 15, g = k >> 4;
    double s[4] = {0, 0, 0, 0};
    const float4* src = reinterpret_cast<const float4*>(d.bn_sums) + q;
    int i = g;
    for (; i + 8 * 15 < nCta; i += 8 * 16) {
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __ldcg(src + (size_t)(i + 8 * j) * (2 * C / 4));
#pragma unroll
      for (int j = 0; j < 16; ++j) { s[0] += (double)v[j].x; s[1] += (double)v[j].y; s[2] += (double)v[j].z; s[3] += (double)v[j].w; }
    }
    for (; i < nCta; i += 8) {
      const float4 v = __ldcg(src + (size_t)i * (2 * C / 4));
      s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z; s[3] += (double)v.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);   // rows g and g^1
    double* sD = reinterpret_cast<double*>(sScratch);   // [4 warps][64]
    if ((k & 31) < 16) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sD[(k >> 5) * 2 * C + 4 * q + e] = s[e];
    }
    __syncthreads();
    if (k < 2 * C) {
      const double t = (sD[k] + sD[2 * C + k]) + (sD[4 * C + k] + sD[6 * C + k]);
      const double N = (double)d.B * d.H * d.W;
      d.bn_tot[k] = (float)(t / N);
      // k < C: sum dV -> d(bn bias); k >= C: sum dV*xhat -> d(bn weight)
      double* a = d.bn_acc + (size_t)bn * 2 * C + k;
      *a = first ? t : (*a + t);
    }
    if (k == 0) *d.counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// (b1)-(b2) for one channel of the thread's patch: dE', dI' + checkpointed E_t, A_t, I_{t+1}, Y1_t, G_e.
// writes Ge = A*E (scratch, for wgrad / gate_wgrad), geb = dG_e*G_e*(1-G_e), gE <- dE'*(1-G_e),
// gI += dR*V*kappa, dAct <- dV; returns the four per-channel partial sums of the patch.
struct BwdAConst {
  const float *Ka, *Ga, *W1, *B1, *M1, *R1;   // shared-memory per-channel constants of step t
};
struct BwdAIn {   // operands of one channel of the patch (prefetched one channel ahead)
  float4 e, a, i2, y1, gx, gi;
};
__device__ __forceinline__ void bwd_a_load(const IntDev& d, int t, int first, size_t idx, BwdAIn& r) {
  r.e = ld4(d.E_at(t) + idx);
  r.a = ld4(d.A_at(t) + idx);
  r.i2 = ld4(d.I_at(t + 1) + idx);
  r.y1 = ld4(d.Y1_at(t) + idx);
  r.gx = ld4(d.Gex_at(t) + idx);
  r.gi = first ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(d.gI + idx);
}
template <int ACT>
__device__ __forceinline__ void bwd_a_channel(const IntDev& d, int t, size_t idx, int c, const float (&gEn)[4],
                                              const BwdAIn& r, const BwdAConst& k, float (&part)[4]) {
  const float e[4] = {r.e.x, r.e.y, r.e.z, r.e.w}, a[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
  const float i2[4] = {r.i2.x, r.i2.y, r.i2.z, r.i2.w}, y1[4] = {r.y1.x, r.y1.y, r.y1.z, r.y1.w};
  const float gx[4] = {r.gx.x, r.gx.y, r.gx.z, r.gx.w}, gi_in[4] = {r.gi.x, r.gi.y, r.gi.z, r.gi.w};
  float oGe[4], oGE[4], oGI[4], oGeb[4], oDv[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float xh = (y1[p] - k.M1[c]) * k.R1[c];
    const float V = fmaf(k.W1[c], xh, k.B1[c]);
    const float kIg = fmaf(k.Ka[c], i2[p], k.Ga[c]);
    const float R = V * kIg;
    const float Ehat = actf<ACT>(R);
    const float dGx = gEn[p] * (Ehat - e[p]);
    oGE[p] = gEn[p] * (1.0f - gx[p]);
    const float dR = gEn[p] * gx[p] * dactf<ACT>(R);
    const float dV = dR * kIg;
    const float dRV = dR * V;
    oGI[p] = fmaf(dRV, k.Ka[c], gi_in[p]);
    oGeb[p] = dGx * gx[p] * (1.0f - gx[p]);
    oDv[p] = dV;
    oGe[p] = a[p] * e[p];
    part[0] = fmaf(dRV, i2[p], part[0]);   // -> d kappa
    part[1] += dRV;                         // -> d gamma
    part[2] += dV;                          // -> sum dV
    part[3] = fmaf(dV, xh, part[3]);        // -> sum dV * xhat
  }
  st4(d.gE + idx, oGE);
  st4(d.gI + idx, oGI);
  st4(d.geb_at(t) + idx, oGeb);
  st4(d.dV_at(t, 1) + idx, oDv);
  st4s(d.GeB_at(t) + idx, oGe, d.dy_tstride != 0);
}
__device__ __forceinline__ void bwd_a_fill_const(const IntDev& d, int t, float* sKa, float* sGa, float* sW1, float* sB1,
                                                 float* sM1, float* sR1) {
  cta_fill_async(sKa, d.kappa, C);     // the caller commits / waits (cp.async)
  cta_fill_async(sGa, d.gamma, C);
  cta_fill_async(sW1, d.bn_w[1], C);
  cta_fill_async(sB1, d.bn_b[1], C);
  if (threadIdx.x < C) {
    const float* st = d.stat(t, 1);
    sM1[threadIdx.x] = st[threadIdx.x];
    sR1[threadIdx.x] = st[C + threadIdx.x];
  }
}

// stand-alone (b1)-(b2) kernel.  Round 1 ran it only for the last time step and fused the other steps into the tail of
// int_bwd_d_kernel (dE' stays on chip); that fusion pushed int_bwd_d to 40-55 spilled words per thread at its 128-register
// budget (as many local as global load instructions, ncu r2i) and measured 4 % SLOWER per training step than this extra launch.
template <int ACT>
__global__ void __launch_bounds__(128, 4) int_bwd_a_kernel(IntDev d, int t, int first) {
  __shared__ __align__(16) float sKa[C], sGa[C], sW1[C], sB1[C], sM1[C], sR1[C];
  __shared__ __align__(16) float sPart[4 * 4 * C], sTot[4 * C];
  __shared__ int sIsLast;
  // prologue on parameters, forward statistics and forward checkpoints only: overlaps the previous kernel's tail (PDL)
  bwd_a_fill_const(d, t, sKa, sGa, sW1, sB1, sM1, sR1);
  cp_async_commit();
  for (int i = threadIdx.x; i < 4 * 4 * C; i += blockDim.x) sPart[i] = 0.f;
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  if (blockIdx.x * 4 + warp < nWT) {
    const int wt0 = blockIdx.x * 4 + warp, b0 = wt0 / tpi;
    const size_t tb = (size_t)b0 * C * P + (wt0 - b0 * tpi) * WPX;
    warp_tile_prefetch(d.E_at(t), tb, P);
    warp_tile_prefetch(d.A_at(t), tb, P);
    warp_tile_prefetch(d.I_at(t + 1), tb, P);
    warp_tile_prefetch(d.Y1_at(t), tb, P);
    warp_tile_prefetch(d.Gex_at(t), tb, P);
  }
  pdl_wait();
  pdl_launch_dependents();
  cp_async_wait<0>();
  __syncthreads();
  const BwdAConst k{sKa, sGa, sW1, sB1, sM1, sR1};
  float* sPw = sPart + warp * 4 * C;
  for (int wt = blockIdx.x * 4 + warp; wt < nWT; wt += gridDim.x * 4) {
    const int b = wt / tpi, pix = (wt - b * tpi) * WPX + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    BwdAIn cur, nxt;
    float4 gcur = ld4(d.gE + base + (size_t)L.oc0 * P), gnxt = gcur;
    bwd_a_load(d, t, first, base + (size_t)L.oc0 * P, cur);
#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
      const int c = L.oc0 + o;
      const size_t idx = base + (size_t)c * P;
      if (o < 7) { bwd_a_load(d, t, first, idx + P, nxt); gnxt = ld4(d.gE + idx + P); }   // prefetch next channel
      const float gEn[4] = {gcur.x, gcur.y, gcur.z, gcur.w};
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      bwd_a_channel<ACT>(d, t, idx, c, gEn, cur, k, part);
      part_add<4>(sPw, c, L, part);
      cur = nxt; gcur = gnxt;
    }
  }
  part_fold<4>(sPart, sTot);
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if (threadIdx.x < C) {
    acc_store(acc + ACC_KAPPA + threadIdx.x, sTot[threadIdx.x], first);
    acc_store(acc + ACC_GAMMA + threadIdx.x, sTot[C + threadIdx.x], first);
  }
  bn_sums_fold(d, 1, sTot + 2 * C, first, &sIsLast, sPart);
}

// (b1)-(b2) as a plain streaming kernel: one CTA per (sample, channel) row, one float4 per thread and iteration, seven
// independent 16-byte loads in flight per thread (the warp-tile version above walks 8 channels per thread with a one-channel
// prefetch: 29 us per launch for 92 MB; this one is bound by the memory system).  Per-row partial sums go to `bn_sums`
// ([B*C][4]); the last CTA folds them per channel in double, in a fixed order (deterministic): BatchNorm-backward means,
// d(bn weight / bias), and the step's d(kappa), d(gamma) (accumulated in accumulator slot 0).
// sample rows per CTA: chosen by the host so that the grid is ~512 CTAs of 256 threads, all resident at once (4 per SM):
// 4 rows at B = 64, 1 row at B = 16 (64x64 frames: a CTA then walks its 16 KB rows in four passes)
inline int bas_rows(int B) { const int r = B / 16; return r < 1 ? 1 : (r > 4 ? 4 : r); }
template <int ACT>
__global__ void __launch_bounds__(256, 4) int_bwd_a_stream_kernel(IntDev d, int t, int first, int BAS_R) {
  __shared__ float sRed[8][4];
  __shared__ int sIsLast;
  const int tid = threadIdx.x, P = d.H * d.W, n4 = P / 4;
  const int bg = blockIdx.x / C, c = blockIdx.x - bg * C;      // CTA = channel c of samples bg*BAS_R .. +BAS_R-1
  // prologue on parameters, forward statistics and forward checkpoints only (PDL)
  const float Ka = __ldg(d.kappa + c), Ga = __ldg(d.gamma + c), W1 = __ldg(d.bn_w[1] + c), B1 = __ldg(d.bn_b[1] + c);
  const float M1 = d.stat(t, 1)[c], R1 = d.stat(t, 1)[C + c];
  for (int r = 0; r < BAS_R; ++r) {
    const int b = bg * BAS_R + r;
    if (b < d.B && (tid & 7) == 0 && tid < n4) {
      const size_t row = ((size_t)b * C + c) * P + 4 * tid;
      prefetch_l2(d.E_at(t) + row); prefetch_l2(d.A_at(t) + (size_t)b * d.a_bstride + (size_t)c * P + 4 * tid);
      prefetch_l2(d.I_at(t + 1) + row); prefetch_l2(d.Y1_at(t) + row); prefetch_l2(d.Gex_at(t) + row);
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  const bool stream = d.dy_tstride != 0;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (int r = 0; r < BAS_R; ++r) {
  const int b = bg * BAS_R + r;
  if (b >= d.B) break;
  const size_t row = ((size_t)b * C + c) * P;
  const float4* E4 = reinterpret_cast<const float4*>(d.E_at(t) + row);
  const float4* A4 = reinterpret_cast<const float4*>(d.A_at(t) + (size_t)b * d.a_bstride + (size_t)c * P);
  const float4* I4 = reinterpret_cast<const float4*>(d.I_at(t + 1) + row);
  const float4* Y4 = reinterpret_cast<const float4*>(d.Y1_at(t) + row);
  const float4* G4 = reinterpret_cast<const float4*>(d.Gex_at(t) + row);
  float4* gE4 = reinterpret_cast<float4*>(d.gE + row);
  float4* gI4 = reinterpret_cast<float4*>(d.gI + row);
  float4* geb4 = reinterpret_cast<float4*>(d.geb_at(t) + row);
  float4* dV4 = reinterpret_cast<float4*>(d.dV_at(t, 1) + row);
  float4* Ge4 = reinterpret_cast<float4*>(d.GeB_at(t) + row);
  for (int i = tid; i < n4; i += 256) {
    const float4 e4 = E4[i], a4 = A4[i], i4 = I4[i], y4 = Y4[i], g4 = G4[i], ge4 = gE4[i];
    const float4 gi4 = first ? make_float4(0.f, 0.f, 0.f, 0.f) : gI4[i];
    const float e[4] = {e4.x, e4.y, e4.z, e4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w}, i2[4] = {i4.x, i4.y, i4.z, i4.w};
    const float y1[4] = {y4.x, y4.y, y4.z, y4.w}, gx[4] = {g4.x, g4.y, g4.z, g4.w}, gEn[4] = {ge4.x, ge4.y, ge4.z, ge4.w};
    const float gi_in[4] = {gi4.x, gi4.y, gi4.z, gi4.w};
    float oGe[4], oGE[4], oGI[4], oGeb[4], oDv[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float xh = (y1[p] - M1) * R1;
      const float V = fmaf(W1, xh, B1);
      const float kIg = fmaf(Ka, i2[p], Ga);
      const float R = V * kIg;
      const float Ehat = actf<ACT>(R);
      const float dGx = gEn[p] * (Ehat - e[p]);
      oGE[p] = gEn[p] * (1.0f - gx[p]);
      const float dR = gEn[p] * gx[p] * dactf<ACT>(R);
      const float dV = dR * kIg;
      const float dRV = dR * V;
      oGI[p] = fmaf(dRV, Ka, gi_in[p]);
      oGeb[p] = dGx * gx[p] * (1.0f - gx[p]);
      oDv[p] = dV;
      oGe[p] = a[p] * e[p];
      part[0] = fmaf(dRV, i2[p], part[0]);   // -> d kappa
      part[1] += dRV;                         // -> d gamma
      part[2] += dV;                          // -> sum dV
      part[3] = fmaf(dV, xh, part[3]);        // -> sum dV * xhat
    }
    gE4[i] = make_float4(oGE[0], oGE[1], oGE[2], oGE[3]);
    gI4[i] = make_float4(oGI[0], oGI[1], oGI[2], oGI[3]);
    geb4[i] = make_float4(oGeb[0], oGeb[1], oGeb[2], oGeb[3]);
    dV4[i] = make_float4(oDv[0], oDv[1], oDv[2], oDv[3]);
    if (stream) __stcs(Ge4 + i, make_float4(oGe[0], oGe[1], oGe[2], oGe[3]));
    else Ge4[i] = make_float4(oGe[0], oGe[1], oGe[2], oGe[3]);
  }
  }
  // sums of the CTA's rows in a fixed order: lanes (xor tree), then the 8 warps
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) part[q] += __shfl_xor_sync(0xffffffffu, part[q], m);
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) sRed[tid >> 5][q] = part[q];
  }
  __syncthreads();
  if (tid < 4) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sacc += sRed[w][tid];
    d.bn_sums[(size_t)blockIdx.x * 4 + tid] = sacc;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) sIsLast = (atomicAdd(d.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (sIsLast) {
    __threadfence();
    if (tid < 4 * C) {
      const int q = tid >> 5, cc = tid & 31;
      double sacc = 0;
      const int ng = (int)gridDim.x / C;
      for (int bb = 0; bb < ng; ++bb) sacc += (double)__ldcg(d.bn_sums + ((size_t)bb * C + cc) * 4 + q);
      const double N = (double)d.B * P;
      if (q == 0) acc_store(d.acc + ACC_KAPPA + cc, (float)sacc, first);
      else if (q == 1) acc_store(d.acc + ACC_GAMMA + cc, (float)sacc, first);
      else {
        const int k2 = (q - 2) * C + cc;          // [0, C): sum dV -> d(bn bias); [C, 2C): sum dV*xhat -> d(bn weight)
        d.bn_tot[k2] = (float)(sacc / N);
        double* ap = d.bn_acc + (size_t)1 * 2 * C + k2;
        *ap = first ? sacc : (*ap + sacc);
      }
    }
    if (tid == 0) *d.counter = 0u;
  }
}

// (b3)/(b9): dAct <- w*rstd*(dAct - mean(dAct) - xhat*mean(dAct*xhat)), xhat from the saved conv output.
__global__ void __launch_bounds__(256) bn_apply_kernel(IntDev d, const float* __restrict__ Y, int t, int bn) {
  {   // the saved conv output is a forward checkpoint: pull this thread's lines HBM -> L2 while the previous kernel drains
    const size_t n4p = (size_t)d.B * C * d.H * d.W / 4;
    if ((threadIdx.x & 7) == 0)
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4p; i += (size_t)gridDim.x * blockDim.x)
        prefetch_l2(Y + 4 * i);
  }
  pdl_wait();
  pdl_launch_dependents();
  const int P = d.H * d.W;
  const size_t n4 = (size_t)d.B * C * P / 4;
  const float* st = d.stat(t, bn);
  float* dA = d.dAct_at(t, bn);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4 / P) % C);
    const float mean = st[c], rstd = st[C + c];
    const float k = __ldg(d.bn_w[bn] + c) * rstd, m = d.bn_tot[c], n = d.bn_tot[C + c];
    float4 g = reinterpret_cast<float4*>(dA)[i];
    const float4 y = reinterpret_cast<const float4*>(Y)[i];
    g.x = k * (g.x - m - (y.x - mean) * rstd * n);
    g.y = k * (g.y - m - (y.y - mean) * rstd * n);
    g.z = k * (g.z - m - (y.z - mean) * rstd * n);
    g.w = k * (g.w - m - (y.w - mean) * rstd * n);
    reinterpret_cast<float4*>(dA)[i] = g;
  }
}

// ------------------------------------------------------------------------------------------------
// (b5)-(b8): complete dI' (gI), geb, checkpointed I_t, A_t, Y0_t, G_i, x_t.
// writes dAct <- dU, dGe, dX, gI <- dI (+ dS for InT), dS (hGRU), gi and X (scratch for gate_wgrad).
struct BwdCIn { float4 gi, cg, I, S, y0, G; };   // one channel of the patch, prefetched one channel ahead
template <int CELL>
__device__ __forceinline__ void bwd_c_load(const IntDev& d, const float* It, const float* At, const float* Y0,
                                           const float* Git, size_t idx, BwdCIn& r) {
  r.gi = ld4(d.gI + idx);
  r.cg = ld4(d.cgI + idx);      // dY1 *7^T W_exc (b4), written by the transposed convolution
  r.I = ld4(It + idx);
  if (CELL != PTK_CELL_INT) r.S = ld4(At + idx);
  r.y0 = ld4(Y0 + idx);
  r.G = ld4(Git + idx);
}
template <int CELL, int ACT>
__global__ void __launch_bounds__(128, 4) int_bwd_c_kernel(IntDev d, int t, int first) {
  extern __shared__ __align__(16) float dsm[];
  float* sWiw = dsm;              // [out][in] = [k][n] of the transposed products
  float* sWiu = sWiw + C * C;
  float* sWew = sWiu + C * C;
  float* sWeu = sWew + C * C;
  float* tiles = sWeu + C * C;
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sAl[C], sMu[C], sW0[C], sB0[C], sM0[C], sR0[C];
  __shared__ __align__(16) float sPart[4 * 4 * C], sTot[4 * C];
  __shared__ int sIsLast;
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  float* sGi = tiles + warp * 2 * TILE_F;
  float* sGe = sGi + TILE_F;
  float* sPw = sPart + warp * 4 * C;
  const float* It = d.I_at(t);
  const float* At = d.A_at(t);
  const float* Y0 = d.Y0_at(t);
  const float* Git = d.Gi_at(t);
  float* dU_out = d.dV_at(t, 0);
  float* gi_out = d.gi_at(t);
  float* X_out = d.X_at(t);
  int wt = blockIdx.x * 4 + warp;
  BwdCIn cur;
  float xin[3][4];
  // prologue on parameters, forward statistics and the (read-only) clip: overlaps the previous kernel's tail (PDL)
  // The mma.sync gate products (ptk_pw.cuh) pay in the forward kernels (-14 %) and in int_bwd_d (-3 %) but not here:
  // all-mma was slower (two live accumulator sets spill, the legacy HMMA tf32 pipe is as busy as the FMA pipe was)
  // and the hybrid below (one FFMA + three mma products) measured exactly equal to four FFMA products, so this
  // kernel keeps the register-tiled FFMA products.  -DPTK_BWD_MMA=1 selects the hybrid to re-measure.
  constexpr bool MMA = PTK_BWD_MMA && (ACT & ACT_FAST) != 0;
  const float* const* gsrc = MMA ? d.gateF : d.gate;
  cta_fill_async(sWiw, d.gate[2], C * C);      // always the plain [out][in] layout: its product stays on the FMA pipe
  const float* sWiw_ffma = sWiw;
  cta_fill_async(sWiu, gsrc[3], C * C);
  cta_fill_async(sWew, gsrc[4], C * C);
  cta_fill_async(sWeu, gsrc[5], C * C);
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  cta_fill_async(sAl, d.alpha, C);
  cta_fill_async(sMu, d.mu, C);
  cta_fill_async(sW0, d.bn_w[0], C);
  cta_fill_async(sB0, d.bn_b[0], C);
  cp_async_commit();
  for (int i = threadIdx.x; i < 4 * 4 * C; i += blockDim.x) sPart[i] = 0.f;
  if (threadIdx.x < C) {
    const float* st = d.stat(t, 0);
    sM0[threadIdx.x] = st[threadIdx.x];
    sR0[threadIdx.x] = st[C + threadIdx.x];
  }
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    frame_px(d.fv, b, t, pix_base + L.pxo, d.W, xin);
    const size_t tb = (size_t)b * C * P + pix_base;      // forward checkpoints of this tile: HBM -> L2
    warp_tile_prefetch(It, tb, P);
    warp_tile_prefetch(Y0, tb, P);
    warp_tile_prefetch(Git, tb, P);
    if (CELL != PTK_CELL_INT) warp_tile_prefetch(At, tb, P);
  }
  pdl_wait();
  pdl_launch_dependents();
  if (wt < nWT) {     // first tile's operands
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    warp_tile_load_async(sGe, d.geb_at(t), (size_t)C * P, b, pix_base, P);
    if (CELL != CELL_NOINH) bwd_c_load<CELL>(d, It, At, Y0, Git, (size_t)b * C * P + pix_base + L.pxo + (size_t)L.oc0 * P, cur);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * 4, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    if (!firstTile) {
      warp_tile_load_async(sGe, d.geb_at(t), (size_t)C * P, b, pix_base, P);
      cp_async_commit();
      if (CELL != CELL_NOINH) bwd_c_load<CELL>(d, It, At, Y0, Git, base + (size_t)L.oc0 * P, cur);
      frame_px(d.fv, b, t, pix, d.W, xin);
    }
    if (CELL == CELL_NOINH) {
      // InT(no_inh=True): only (b5) exists.  dS = We_w^T ge feeds dE (S = E), dGe = dI' + We_u^T ge.
#pragma unroll 1
      for (int o = 0; o < 8; ++o) {     // X is still needed by the gate weight gradients
        float X[4];
        stem_pre(sWp, sbp, L.oc0 + o, xin, X);
#pragma unroll
        for (int p = 0; p < 4; ++p) X[p] = actf<ACT>(X[p]);
        st4s(d.X_at(t) + base + (size_t)(L.oc0 + o) * P, X, d.dy_tstride != 0);
      }
      cp_async_wait<0>();
      __syncwarp();
      float acc[4][8];
      if (MMA) {
        Frag f;
        frag_zero(f);
        warp_mma_gemm(f, sGe, sWew);
        acc_fill(acc, nullptr, L);
        acc_to_global(d.dS, base, P, L, acc);          // zero, then the product in fragment layout
        __syncwarp();
        frag_add_to_global(d.dS, (size_t)b * C * P + pix_base, P, f);
        frag_zero(f);
        warp_mma_gemm(f, sGe, sWeu);
        __syncwarp();
        frag_to_tile(sGe, f);                           // ge is dead once both products have read it
        __syncwarp();
      } else {
        acc_fill(acc, nullptr, L);
        warp_gemm(acc, sGe, sWew, L);
        acc_to_global(d.dS, base, P, L, acc);
      }
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const float4 g1 = ld4(d.gI + base + (size_t)(L.oc0 + o) * P), g2 = ld4(d.cgI + base + (size_t)(L.oc0 + o) * P);
        acc[0][o] = g1.x + g2.x; acc[1][o] = g1.y + g2.y; acc[2][o] = g1.z + g2.z; acc[3][o] = g1.w + g2.w;
      }
      if (MMA) acc_add_tile(acc, sGe, L);
      else warp_gemm(acc, sGe, sWeu, L);
      acc_to_global(d.dGe, base, P, L, acc);
      acc_fill(acc, nullptr, L);
      acc_to_global(d.gI, base, P, L, acc);   // the step ignores its inhibition input: dI for step t-1 is zero
      __syncwarp();
      continue;
    }
#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
      const int c = L.oc0 + o;
      const size_t idx = base + (size_t)c * P;
      float gIn[4], I[4], S[4], y0[4], G[4], X[4], oDu[4], oGi[4], oDI[4], oZ[4], oDs[4];
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      BwdCIn nxt;
      if (o < 7) bwd_c_load<CELL>(d, It, At, Y0, Git, idx + P, nxt);     // prefetch the next channel
      gIn[0] = cur.gi.x + cur.cg.x; gIn[1] = cur.gi.y + cur.cg.y; gIn[2] = cur.gi.z + cur.cg.z; gIn[3] = cur.gi.w + cur.cg.w;
      I[0] = cur.I.x; I[1] = cur.I.y; I[2] = cur.I.z; I[3] = cur.I.w;
      if (CELL == PTK_CELL_INT) {
#pragma unroll
        for (int p = 0; p < 4; ++p) S[p] = I[p];
      } else {
        S[0] = cur.S.x; S[1] = cur.S.y; S[2] = cur.S.z; S[3] = cur.S.w;
      }
      y0[0] = cur.y0.x; y0[1] = cur.y0.y; y0[2] = cur.y0.z; y0[3] = cur.y0.w;
      G[0] = cur.G.x; G[1] = cur.G.y; G[2] = cur.G.z; G[3] = cur.G.w;
      stem_pre(sWp, sbp, c, xin, X);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        X[p] = actf<ACT>(X[p]);
        const float xh = (y0[p] - sM0[c]) * sR0[c];
        const float U = fmaf(sW0[c], xh, sB0[c]);
        const float aSm = fmaf(sAl[c], S[p], sMu[c]);
        const float Pp = U * aSm;
        const float Q = actf<ACT>(Pp);
        const float d1 = X[p] - Q;
        const float Ihat = actf<ACT>(d1);
        const float dGi = gIn[p] * (Ihat - I[p]);
        oDI[p] = gIn[p] * (1.0f - G[p]);
        const float zz = gIn[p] * G[p] * dactf<ACT>(d1);
        oGi[p] = dGi * G[p] * (1.0f - G[p]);
        const float dP = -zz * dactf<ACT>(Pp);
        const float dPU = dP * U;
        const float dU = dP * aSm;
        oDu[p] = dU;
        oZ[p] = zz;
        oDs[p] = dPU * sAl[c] + (CELL == PTK_CELL_INT ? oDI[p] : 0.0f);   // InT: dI + dS lands in gI
        part[0] = fmaf(dPU, S[p], part[0]);   // -> d alpha
        part[1] += dPU;                        // -> d mu
        part[2] += dU;                         // -> sum dU
        part[3] = fmaf(dU, xh, part[3]);       // -> sum dU * xhat
      }
      st4(dU_out + idx, oDu);
      st4s(gi_out + idx, oGi, d.dy_tstride != 0);
      st4s(X_out + idx, X, d.dy_tstride != 0);
      st4(d.dX + idx, oZ);        // parked: completed by the transposed products below
      st4(d.dS + idx, oDs);
      tile_st(sGi, L, o, oGi);
      if (CELL == PTK_CELL_HGRU) st4(d.gI + idx, oDI);
      part_add<4>(sPw, c, L, part);
      if (o < 7) cur = nxt;
    }
    cp_async_wait<0>();
    __syncwarp();
    float acc[4][8];
    if (MMA) {
      // hybrid: the product whose result must be added to the parked z stays on the FMA pipe (its accumulators
      // start from global memory in patch layout); the other three run on mma.sync, one accumulator set at a time
      acc_from_global(acc, d.dX, base, P, L);
      warp_gemm(acc, sGi, sWiw_ffma, L);  // dX = z + Wi_w^T gi
      acc_to_global(d.dX, base, P, L, acc);
      Frag f;
      frag_zero(f);
      warp_mma_gemm(f, sGi, sWiu);       // dS += Wi_u^T gi + We_w^T ge   (the gi tile is dead after this)
      warp_mma_gemm(f, sGe, sWew);
      __syncwarp();
      frag_to_tile(sGi, f);
      __syncwarp();
      acc_from_global(acc, d.dS, base, P, L);
      acc_add_tile(acc, sGi, L);
      if (CELL == PTK_CELL_INT) acc_to_global(d.gI, base, P, L, acc);   // (b10) InT: dI += dS
      else acc_to_global(d.dS, base, P, L, acc);                         //       hGRU: dA += dS (in bwd_d)
      frag_zero(f);
      warp_mma_gemm(f, sGe, sWeu);       // dGe = We_u^T ge
      __syncwarp();
      frag_to_tile(sGe, f);
      __syncwarp();
      acc_fill(acc, nullptr, L);
      acc_add_tile(acc, sGe, L);
      acc_to_global(d.dGe, base, P, L, acc);
    } else {
      acc_from_global(acc, d.dX, base, P, L);
      warp_gemm(acc, sGi, sWiw, L);      // dX = z + Wi_w^T gi
      acc_to_global(d.dX, base, P, L, acc);
      acc_from_global(acc, d.dS, base, P, L);
      warp_gemm(acc, sGi, sWiu, L);      // dS += Wi_u^T gi
      warp_gemm(acc, sGe, sWew, L);      //     + We_w^T ge
      if (CELL == PTK_CELL_INT) acc_to_global(d.gI, base, P, L, acc);   // (b10) InT: dI += dS
      else acc_to_global(d.dS, base, P, L, acc);                         //       hGRU: dA += dS (in bwd_d)
      acc_fill(acc, nullptr, L);
      warp_gemm(acc, sGe, sWeu, L);      // dGe = We_u^T ge
      acc_to_global(d.dGe, base, P, L, acc);
    }
    __syncwarp();
  }
  if (CELL == CELL_NOINH) return;      // no alpha / mu / BatchNorm0 in this variant
  part_fold<4>(sPart, sTot);
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if (threadIdx.x < C) {
    acc_store(acc + ACC_ALPHA + threadIdx.x, sTot[threadIdx.x], first);
    acc_store(acc + ACC_MU + threadIdx.x, sTot[C + threadIdx.x], first);
  }
  bn_sums_fold(d, 0, sTot + 2 * C, first, &sIsLast, sPart);
}

// ------------------------------------------------------------------------------------------------
// (b5)-(b8) with the tile's operands STAGED in shared memory (InT cell, tensor-core modes): see int_bwd_d_staged_kernel.
// All six operand tiles of a warp's tile are requested at once (cp.async); z and dS are parked in the tiles they replace
// instead of in global memory, the four transposed gate products run on mma.sync one accumulator set at a time and are added
// into those tiles in shared memory, and dX / dI / dGe leave once, in patch layout.  One 7-warp CTA per SM.
__device__ __forceinline__ void frag_add_to_tile(float* tile, const Frag& f) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        tile[(n * 8 + 2 * t + (e & 1)) * TP + m * 16 + g + 8 * (e >> 1)] += f.c[m][n][e];
}

// InT: 7 warps x 6 operand tiles.  hGRU (S = A_t, models/ffhgru_hierarchy.py:146-150) needs a seventh tile: 6 warps x 7 tiles.
template <int CELL> constexpr int bcs_nw() { return CELL == PTK_CELL_HGRU ? 6 : 7; }
template <int CELL> constexpr int bcs_nt() { return CELL == PTK_CELL_HGRU ? 7 : 6; }
template <int CELL> constexpr size_t sm_bwd_cs() { return (size_t)(4 * C * C + bcs_nw<CELL>() * bcs_nt<CELL>() * TILE_F) * sizeof(float); }

// bn_sums_fold for CTAs of more than 128 threads (only the first 128 take part in the load phase)
__device__ __forceinline__ void bn_sums_fold_wide(const IntDev& d, int bn, const float* sPart /*[2][32]*/, int first,
                                                  int* sIsLast, float* sScratch /* >= 2 KB, 8-byte aligned, free */) {
  const int k = threadIdx.x, nCta = gridDim.x;
  if (k < 2 * C) d.bn_sums[(size_t)blockIdx.x * 2 * C + k] = sPart[k];
  __threadfence();
  __syncthreads();
  if (k == 0) *sIsLast = (atomicAdd(d.counter, 1u) == (unsigned)(nCta - 1));
  __syncthreads();
  if (*sIsLast) {
    __threadfence();
    double* sD = reinterpret_cast<double*>(sScratch);   // [4 warps][64]
    if (k < 128) {
      const int q = k & 15, g = k >> 4;
      double s[4] = {0, 0, 0, 0};
      const float4* src = reinterpret_cast<const float4*>(d.bn_sums) + q;
      int i = g;
      for (; i + 8 * 15 < nCta; i += 8 * 16) {
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __ldcg(src + (size_t)(i + 8 * j) * (2 * C / 4));
#pragma unroll
        for (int j = 0; j < 16; ++j) { s[0] += (double)v[j].x; s[1] += (double)v[j].y; s[2] += (double)v[j].z; s[3] += (double)v[j].w; }
      }
      for (; i < nCta; i += 8) {
        const float4 v = __ldcg(src + (size_t)i * (2 * C / 4));
        s[0] += (double)v.x; s[1] += (double)v.y; s[2] += (double)v.z; s[3] += (double)v.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);   // rows g and g^1
      if ((k & 31) < 16) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sD[(k >> 5) * 2 * C + 4 * q + e] = s[e];
      }
    }
    __syncthreads();
    if (k < 2 * C) {
      const double tt = (sD[k] + sD[2 * C + k]) + (sD[4 * C + k] + sD[6 * C + k]);
      const double N = (double)d.B * d.H * d.W;
      d.bn_tot[k] = (float)(tt / N);
      double* a = d.bn_acc + (size_t)bn * 2 * C + k;
      *a = first ? tt : (*a + tt);
    }
    if (k == 0) *d.counter = 0u;
  }
}

template <int CELL, int ACT>
__global__ void __launch_bounds__(32 * bcs_nw<CELL>(), 1) int_bwd_c_staged_kernel(IntDev d, int t, int first) {
  constexpr int BCS_NW = bcs_nw<CELL>(), BCS_NT = bcs_nt<CELL>();
  extern __shared__ __align__(16) float dsm[];
  float* sWiw = dsm;              // [out][in] = [k][n] of the transposed products, mma B-fragment order
  float* sWiu = sWiw + C * C;
  float* sWew = sWiu + C * C;
  float* sWeu = sWew + C * C;
  float* tiles = sWeu + C * C;
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sAl[C], sMu[C], sW0[C], sB0[C], sM0[C], sR0[C];
  __shared__ __align__(16) float sPart[BCS_NW * 4 * C], sTot[4 * C];
  __shared__ int sIsLast;
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  const float* It = d.I_at(t);
  const float* Y0 = d.Y0_at(t);
  const float* Git = d.Gi_at(t);
  float* dU_out = d.dV_at(t, 0);
  float* gi_out = d.gi_at(t);
  float* X_out = d.X_at(t);
  float* tgI = tiles + (size_t)warp * BCS_NT * TILE_F;   // dI' partial -> dS parked -> dI (+ dS) of step t
  float* tcg = tgI + TILE_F;                             // dY1 *7^T W_exc (b4) -> z parked -> dX
  float* tI = tcg + TILE_F;                              // I_t -> dGe
  float* tY = tI + TILE_F;                               // Y0_t
  float* tG = tY + TILE_F;                               // G_i -> gi
  float* tge = tG + TILE_F;                              // ge
  float* tS = tge + TILE_F;                              // hGRU only: S = A_t
  int wt = blockIdx.x * BCS_NW + warp;
  float xin[3][4];
  // prologue on parameters, forward statistics, the (read-only) clip and forward checkpoints: overlaps the previous kernel (PDL)
  cta_fill_async(sWiw, d.gateF[2], C * C);
  cta_fill_async(sWiu, d.gateF[3], C * C);
  cta_fill_async(sWew, d.gateF[4], C * C);
  cta_fill_async(sWeu, d.gateF[5], C * C);
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  cta_fill_async(sAl, d.alpha, C);
  cta_fill_async(sMu, d.mu, C);
  cta_fill_async(sW0, d.bn_w[0], C);
  cta_fill_async(sB0, d.bn_b[0], C);
  for (int i = threadIdx.x; i < BCS_NW * 4 * C; i += blockDim.x) sPart[i] = 0.f;
  if (threadIdx.x < C) {
    const float* st = d.stat(t, 0);
    sM0[threadIdx.x] = st[threadIdx.x];
    sR0[threadIdx.x] = st[C + threadIdx.x];
  }
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    frame_px(d.fv, b, t, pix_base + L.pxo, d.W, xin);
    warp_tile_load_async(tI, It, (size_t)C * P, b, pix_base, P);
    warp_tile_load_async(tY, Y0, (size_t)C * P, b, pix_base, P);
    warp_tile_load_async(tG, Git, (size_t)C * P, b, pix_base, P);
    if (CELL == PTK_CELL_HGRU) warp_tile_load_async(tS, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P);
  }
  cp_async_commit();
  pdl_wait();
  pdl_launch_dependents();
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    warp_tile_load_async(tgI, d.gI, (size_t)C * P, b, pix_base, P);
    warp_tile_load_async(tcg, d.cgI, (size_t)C * P, b, pix_base, P);
    warp_tile_load_async(tge, d.geb_at(t), (size_t)C * P, b, pix_base, P);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  float* sPw = sPart + warp * 4 * C;
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * BCS_NW, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    if (!firstTile) {
      warp_tile_load_async(tI, It, (size_t)C * P, b, pix_base, P);
      warp_tile_load_async(tY, Y0, (size_t)C * P, b, pix_base, P);
      warp_tile_load_async(tG, Git, (size_t)C * P, b, pix_base, P);
      if (CELL == PTK_CELL_HGRU) warp_tile_load_async(tS, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P);
      warp_tile_load_async(tgI, d.gI, (size_t)C * P, b, pix_base, P);
      warp_tile_load_async(tcg, d.cgI, (size_t)C * P, b, pix_base, P);
      warp_tile_load_async(tge, d.geb_at(t), (size_t)C * P, b, pix_base, P);
      cp_async_commit();
      frame_px(d.fv, b, t, pix, d.W, xin);
      cp_async_wait<0>();
      __syncwarp();
    }
#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
      const int c = L.oc0 + o;
      const size_t idx = base + (size_t)c * P;
      float gi_[4], cg[4], I[4], S[4], y0[4], G[4], X[4], oDu[4], oGi[4], oZ[4], oDs[4], oDI[4];
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      tile_ld(tgI, L, o, gi_);
      tile_ld(tcg, L, o, cg);
      tile_ld(tI, L, o, I);
      tile_ld(tY, L, o, y0);
      tile_ld(tG, L, o, G);
      if (CELL == PTK_CELL_HGRU) tile_ld(tS, L, o, S);
      else {
#pragma unroll
        for (int p = 0; p < 4; ++p) S[p] = I[p];               // InT: S = I
      }
      stem_pre(sWp, sbp, c, xin, X);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float gIn = gi_[p] + cg[p];
        X[p] = actf<ACT>(X[p]);
        const float xh = (y0[p] - sM0[c]) * sR0[c];
        const float U = fmaf(sW0[c], xh, sB0[c]);
        const float aSm = fmaf(sAl[c], S[p], sMu[c]);
        const float Pp = U * aSm;
        const float Q = actf<ACT>(Pp);
        const float d1 = X[p] - Q;
        const float Ihat = actf<ACT>(d1);
        const float dGi = gIn * (Ihat - I[p]);
        oDI[p] = gIn * (1.0f - G[p]);
        const float zz = gIn * G[p] * dactf<ACT>(d1);
        oGi[p] = dGi * G[p] * (1.0f - G[p]);
        const float dP = -zz * dactf<ACT>(Pp);
        const float dPU = dP * U;
        const float dU = dP * aSm;
        oDu[p] = dU;
        oZ[p] = zz;
        oDs[p] = dPU * sAl[c] + (CELL == PTK_CELL_INT ? oDI[p] : 0.0f);   // InT: dI + dS lands in gI
        part[0] = fmaf(dPU, S[p], part[0]);   // -> d alpha
        part[1] += dPU;                        // -> d mu
        part[2] += dU;                         // -> sum dU
        part[3] = fmaf(dU, xh, part[3]);       // -> sum dU * xhat
      }
      st4(dU_out + idx, oDu);
      st4s(gi_out + idx, oGi, d.dy_tstride != 0);
      st4s(X_out + idx, X, d.dy_tstride != 0);
      tile_st(tG, L, o, oGi);
      tile_st(tcg, L, o, oZ);         // parked: completed by the transposed products below
      tile_st(tgI, L, o, oDs);
      if (CELL == PTK_CELL_HGRU) st4(d.gI + idx, oDI);        // hGRU: dI of step t-1 is the point-wise part alone
      part_add<4>(sPw, c, L, part);
    }
    __syncwarp();
    {
      Frag f;                          // one accumulator set at a time
      frag_zero(f);
      warp_mma_gemm(f, tG, sWiw);      // dX = z + Wi_w^T gi
      frag_add_to_tile(tcg, f);
      frag_zero(f);
      warp_mma_gemm(f, tG, sWiu);      // dI += dS + Wi_u^T gi + We_w^T ge
      warp_mma_gemm(f, tge, sWew);
      frag_add_to_tile(tgI, f);
      frag_zero(f);
      warp_mma_gemm(f, tge, sWeu);     // dGe = We_u^T ge
      frag_to_tile(tI, f);             // I_t is dead
    }
    __syncwarp();
#pragma unroll 2
    for (int o = 0; o < 8; ++o) {
      const size_t idx = base + (size_t)(L.oc0 + o) * P;
      float v[4];
      tile_ld(tcg, L, o, v);
      st4(d.dX + idx, v);
      tile_ld(tgI, L, o, v);
      st4((CELL == PTK_CELL_INT ? d.gI : d.dS) + idx, v);    // (b10) InT: dI += dS;  hGRU: dA += dS (in bwd_d)
      tile_ld(tI, L, o, v);
      st4(d.dGe + idx, v);
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < BCS_NW; ++w) sacc += sPart[w * 4 * C + i];
    sTot[i] = sacc;
  }
  __syncthreads();
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if (threadIdx.x < C) {
    acc_store(acc + ACC_ALPHA + threadIdx.x, sTot[threadIdx.x], first);
    acc_store(acc + ACC_MU + threadIdx.x, sTot[C + threadIdx.x], first);
  }
  bn_sums_fold_wide(d, 0, sTot + 2 * C, first, &sIsLast, sPart);
}

// ------------------------------------------------------------------------------------------------
// (b5)-(b8), staged operands, TWO warps per tile (InT cell, tensor-core modes).  int_bwd_c_staged_kernel keeps 7 warps per SM
// (six 4.6 KB operand tiles per warp fill the shared memory) and each warp walks a long dependent instruction stream: issue
// slots 29 % busy, 42 us per launch.  Here the two warps of a pair share one tile's operand tiles and split its CHANNELS:
// warp h takes channels oc0 + 4h .. + 3 of every lane's patch in the element-wise loops and output n-tiles 2h, 2h + 1 of the
// transposed gate products (all K = 32 input channels are read by both).  Same shared memory, twice the warps in flight.
__device__ __forceinline__ void pair_sync(int slot) { named_bar_sync(1 + slot, 64); }
// half of warp_tile_load_async: channels 16 h .. 16 h + 15
__device__ __forceinline__ void warp_tile_load_async_half(float* tile, const float* __restrict__ src, size_t bstride, int b,
                                                          int pix_base, int P, int h) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = lane + 32 * (4 * h + j), c = q >> 3, part = q & 7;
    cp_async16(tile + c * TP + part * 4, src + (size_t)b * bstride + (size_t)c * P + pix_base + part * 4);
  }
}
struct Frag2 { float c[2][2][4]; };   // [m-tile][local n-tile][4]: n-tiles n0, n0 + 1
__device__ __forceinline__ void frag2_zero(Frag2& f) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) f.c[m][n][e] = 0.f;
}
// warp_mma_gemm restricted to output n-tiles n0, n0 + 1 (same products, same order per accumulator)
__device__ __forceinline__ void warp_mma_gemm_n2(Frag2& f, const float* __restrict__ tile, const float* __restrict__ sWF, int n0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float* a = tile + (kt * 8 + t) * TP + m * 16 + g;
      split_tf32(a[0], ah[m][0], al[m][0]);
      split_tf32(a[8], ah[m][1], al[m][1]);
      split_tf32(a[4 * TP], ah[m][2], al[m][2]);
      split_tf32(a[4 * TP + 8], ah[m][3], al[m][3]);
    }
#pragma unroll
    for (int nl = 0; nl < 2; ++nl) {
      const float2 b = *reinterpret_cast<const float2*>(sWF + ((kt * 4 + n0 + nl) * 32 + lane) * 2);
      uint32_t bh[2], bl[2];
      split_tf32(b.x, bh[0], bl[0]);
      split_tf32(b.y, bh[1], bl[1]);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        mma_tf32(f.c[m][nl], al[m], bh);
        mma_tf32(f.c[m][nl], ah[m], bl);
        mma_tf32(f.c[m][nl], ah[m], bh);
      }
    }
  }
}
// two products with the SAME A operand (tile) against two matrices: the A fragments are loaded and split once.
// Per accumulator the sequence of mma instructions is the one of two warp_mma_gemm_n2 calls (bit-identical sums).
template <bool SECOND>
__device__ __forceinline__ void warp_mma_gemm_n2x2(Frag2& f1, Frag2& f2, const float* __restrict__ tile, const float* __restrict__ sWF1,
                                                   const float* __restrict__ sWF2, int n0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float* a = tile + (kt * 8 + t) * TP + m * 16 + g;
      split_tf32(a[0], ah[m][0], al[m][0]);
      split_tf32(a[8], ah[m][1], al[m][1]);
      split_tf32(a[4 * TP], ah[m][2], al[m][2]);
      split_tf32(a[4 * TP + 8], ah[m][3], al[m][3]);
    }
#pragma unroll
    for (int w = 0; w < (SECOND ? 2 : 1); ++w) {
      const float* sWF = w == 0 ? sWF1 : sWF2;
#pragma unroll
      for (int nl = 0; nl < 2; ++nl) {
        const float2 b = *reinterpret_cast<const float2*>(sWF + ((kt * 4 + n0 + nl) * 32 + lane) * 2);
        uint32_t bh[2], bl[2];
        split_tf32(b.x, bh[0], bl[0]);
        split_tf32(b.y, bh[1], bl[1]);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          float (&c)[4] = w == 0 ? f1.c[m][nl] : f2.c[m][nl];
          mma_tf32(c, al[m], bh);
          mma_tf32(c, ah[m], bl);
          mma_tf32(c, ah[m], bh);
        }
      }
    }
  }
}
template <bool ADD>
__device__ __forceinline__ void frag2_to_tile(float* tile, const Frag2& f, int n0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nl = 0; nl < 2; ++nl)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float* q = tile + ((n0 + nl) * 8 + 2 * t + (e & 1)) * TP + m * 16 + g + 8 * (e >> 1);
        if (ADD) *q += f.c[m][nl][e]; else *q = f.c[m][nl][e];
      }
}

// tile slots, operand tiles per slot (hGRU: a seventh tile for S = A_t on six slots), warps = 2 per slot
template <int CELL> constexpr int bcp_ns() { return CELL == PTK_CELL_HGRU ? 6 : 7; }
template <int CELL> constexpr int bcp_nt() { return CELL == PTK_CELL_HGRU ? 7 : 6; }
template <int CELL> constexpr size_t sm_bwd_cp() { return (size_t)(4 * C * C + bcp_ns<CELL>() * bcp_nt<CELL>() * TILE_F) * sizeof(float); }

template <int CELL, int ACT>
__global__ void __launch_bounds__(64 * bcp_ns<CELL>(), 1) int_bwd_c_pair_kernel(IntDev d, int t, int first) {
  constexpr int BCP_NS = bcp_ns<CELL>(), BCP_NT = bcp_nt<CELL>(), BCP_NW = 2 * BCP_NS;
  extern __shared__ __align__(16) float dsm[];
  float* sWiw = dsm;              // [out][in] = [k][n] of the transposed products, mma B-fragment order
  float* sWiu = sWiw + C * C;
  float* sWew = sWiu + C * C;
  float* sWeu = sWew + C * C;
  float* tiles = sWeu + C * C;
  __shared__ __align__(16) float sWp[C * 3], sbp[C], sAl[C], sMu[C], sW0[C], sB0[C], sM0[C], sR0[C];
  __shared__ __align__(16) float sPart[BCP_NW * 4 * C], sTot[4 * C];
  __shared__ int sIsLast;
  const Lane L;
  const int warp = threadIdx.x >> 5, slot = warp >> 1, h = warp & 1, O0 = 4 * h, N0 = 2 * h;
  const int P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  const float* It = d.I_at(t);
  const float* Y0 = d.Y0_at(t);
  const float* Git = d.Gi_at(t);
  float* dU_out = d.dV_at(t, 0);
  float* gi_out = d.gi_at(t);
  float* X_out = d.X_at(t);
  float* tgI = tiles + (size_t)slot * BCP_NT * TILE_F;   // dI' partial -> dS parked -> dI (+ dS) of step t
  float* tcg = tgI + TILE_F;                             // dY1 *7^T W_exc (b4) -> z parked -> dX
  float* tI = tcg + TILE_F;                              // I_t -> dGe
  float* tY = tI + TILE_F;                               // Y0_t
  float* tG = tY + TILE_F;                               // G_i -> gi
  float* tge = tG + TILE_F;                              // ge
  float* tS = tge + TILE_F;                              // hGRU only: S = A_t
  int wt = blockIdx.x * BCP_NS + slot;
  float xin[3][4];
  // prologue on parameters, forward statistics, the (read-only) clip and forward checkpoints: overlaps the previous kernel (PDL)
  cta_fill_async(sWiw, d.gateF[2], C * C);
  cta_fill_async(sWiu, d.gateF[3], C * C);
  cta_fill_async(sWew, d.gateF[4], C * C);
  cta_fill_async(sWeu, d.gateF[5], C * C);
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  cta_fill_async(sAl, d.alpha, C);
  cta_fill_async(sMu, d.mu, C);
  cta_fill_async(sW0, d.bn_w[0], C);
  cta_fill_async(sB0, d.bn_b[0], C);
  for (int i = threadIdx.x; i < BCP_NW * 4 * C; i += blockDim.x) sPart[i] = 0.f;
  if (threadIdx.x < C) {
    const float* st = d.stat(t, 0);
    sM0[threadIdx.x] = st[threadIdx.x];
    sR0[threadIdx.x] = st[C + threadIdx.x];
  }
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    frame_px(d.fv, b, t, pix_base + L.pxo, d.W, xin);
    warp_tile_load_async_half(tI, It, (size_t)C * P, b, pix_base, P, h);
    warp_tile_load_async_half(tY, Y0, (size_t)C * P, b, pix_base, P, h);
    warp_tile_load_async_half(tG, Git, (size_t)C * P, b, pix_base, P, h);
    if (CELL == PTK_CELL_HGRU) warp_tile_load_async_half(tS, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P, h);
  }
  cp_async_commit();
  pdl_wait();
  pdl_launch_dependents();
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    warp_tile_load_async_half(tgI, d.gI, (size_t)C * P, b, pix_base, P, h);
    warp_tile_load_async_half(tcg, d.cgI, (size_t)C * P, b, pix_base, P, h);
    warp_tile_load_async_half(tge, d.geb_at(t), (size_t)C * P, b, pix_base, P, h);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  float* sPw = sPart + warp * 4 * C;
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * BCP_NS, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    if (!firstTile) {
      warp_tile_load_async_half(tI, It, (size_t)C * P, b, pix_base, P, h);
      warp_tile_load_async_half(tY, Y0, (size_t)C * P, b, pix_base, P, h);
      warp_tile_load_async_half(tG, Git, (size_t)C * P, b, pix_base, P, h);
      if (CELL == PTK_CELL_HGRU) warp_tile_load_async_half(tS, d.A_at(t), (size_t)d.a_bstride, b, pix_base, P, h);
      warp_tile_load_async_half(tgI, d.gI, (size_t)C * P, b, pix_base, P, h);
      warp_tile_load_async_half(tcg, d.cgI, (size_t)C * P, b, pix_base, P, h);
      warp_tile_load_async_half(tge, d.geb_at(t), (size_t)C * P, b, pix_base, P, h);
      cp_async_commit();
      frame_px(d.fv, b, t, pix, d.W, xin);
      cp_async_wait<0>();
      pair_sync(slot);                 // both halves of every operand tile have landed
    }
#pragma unroll 1
    for (int oo = 0; oo < 4; ++oo) {
      const int o = O0 + oo, c = L.oc0 + o;
      const size_t idx = base + (size_t)c * P;
      float gi_[4], cg[4], I[4], S[4], y0[4], G[4], X[4], oDu[4], oGi[4], oZ[4], oDs[4], oDI[4];
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      tile_ld(tgI, L, o, gi_);
      tile_ld(tcg, L, o, cg);
      tile_ld(tI, L, o, I);
      tile_ld(tY, L, o, y0);
      tile_ld(tG, L, o, G);
      if (CELL == PTK_CELL_HGRU) tile_ld(tS, L, o, S);
      else {
#pragma unroll
        for (int p = 0; p < 4; ++p) S[p] = I[p];               // InT: S = I
      }
      stem_pre(sWp, sbp, c, xin, X);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float gIn = gi_[p] + cg[p];
        X[p] = actf<ACT>(X[p]);
        const float xh = (y0[p] - sM0[c]) * sR0[c];
        const float U = fmaf(sW0[c], xh, sB0[c]);
        const float aSm = fmaf(sAl[c], S[p], sMu[c]);
        const float Pp = U * aSm;
        const float Q = actf<ACT>(Pp);
        const float d1 = X[p] - Q;
        const float Ihat = actf<ACT>(d1);
        const float dGi = gIn * (Ihat - I[p]);
        oDI[p] = gIn * (1.0f - G[p]);
        const float zz = gIn * G[p] * dactf<ACT>(d1);
        oGi[p] = dGi * G[p] * (1.0f - G[p]);
        const float dP = -zz * dactf<ACT>(Pp);
        const float dPU = dP * U;
        const float dU = dP * aSm;
        oDu[p] = dU;
        oZ[p] = zz;
        oDs[p] = dPU * sAl[c] + (CELL == PTK_CELL_INT ? oDI[p] : 0.0f);   // InT: dI + dS lands in gI
        part[0] = fmaf(dPU, S[p], part[0]);   // -> d alpha
        part[1] += dPU;                        // -> d mu
        part[2] += dU;                         // -> sum dU
        part[3] = fmaf(dU, xh, part[3]);       // -> sum dU * xhat
      }
      st4(dU_out + idx, oDu);
      st4s(gi_out + idx, oGi, d.dy_tstride != 0);
      st4s(X_out + idx, X, d.dy_tstride != 0);
      tile_st(tG, L, o, oGi);
      tile_st(tcg, L, o, oZ);         // parked: completed by the transposed products below
      tile_st(tgI, L, o, oDs);
      if (CELL == PTK_CELL_HGRU) st4(d.gI + idx, oDI);        // hGRU: dI of step t-1 is the point-wise part alone
      part_add<4>(sPw, c, L, part);
    }
    pair_sync(slot);                   // gi, z, dS complete for all 32 channels; I_t dead
    {
      Frag2 f1, f2;                    // this warp's output channels 16 h .. 16 h + 15; each operand tile is split once
      frag2_zero(f1);
      frag2_zero(f2);
      warp_mma_gemm_n2x2<true>(f1, f2, tG, sWiw, sWiu, N0);     // dX = z + Wi_w^T gi;   dI += dS + Wi_u^T gi ...
      frag2_to_tile<true>(tcg, f1, N0);
      frag2_zero(f1);
      warp_mma_gemm_n2x2<true>(f2, f1, tge, sWew, sWeu, N0);    // ... + We_w^T ge;       dGe = We_u^T ge
      frag2_to_tile<true>(tgI, f2, N0);
      frag2_to_tile<false>(tI, f1, N0);
    }
    pair_sync(slot);
#pragma unroll 2
    for (int oo = 0; oo < 4; ++oo) {
      const size_t idx = base + (size_t)(L.oc0 + O0 + oo) * P;
      float v[4];
      tile_ld(tcg, L, O0 + oo, v);
      st4(d.dX + idx, v);
      tile_ld(tgI, L, O0 + oo, v);
      st4((CELL == PTK_CELL_INT ? d.gI : d.dS) + idx, v);    // (b10) InT: dI += dS;  hGRU: dA += dS (in bwd_d)
      tile_ld(tI, L, O0 + oo, v);
      st4(d.dGe + idx, v);
    }
    pair_sync(slot);                   // the partner is through with the tiles before the next tile's loads land
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < BCP_NW; ++w) sacc += sPart[w * 4 * C + i];
    sTot[i] = sacc;
  }
  __syncthreads();
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if (threadIdx.x < C) {
    acc_store(acc + ACC_ALPHA + threadIdx.x, sTot[threadIdx.x], first);
    acc_store(acc + ACC_MU + threadIdx.x, sTot[C + threadIdx.x], first);
  }
  bn_sums_fold_wide(d, 0, sTot + 2 * C, first, &sIsLast, sPart);
}

// ------------------------------------------------------------------------------------------------
struct BwdDIn { float4 a, dge, cg, e, ge, ds; };   // one channel of the patch, prefetched one channel ahead
template <int CELL>
__device__ __forceinline__ void bwd_d_load(const IntDev& d, const float* At, const float* Et, size_t idx, BwdDIn& r) {
  r.a = ld4(At + idx);
  r.dge = ld4(d.dGe + idx);
  r.cg = ld4(d.cGe + idx);
  r.e = ld4(Et + idx);
  r.ge = ld4(d.gE + idx);
  if (CELL != PTK_CELL_INT) r.ds = ld4(d.dS + idx);
}
// (b10)-(b12): complete dGe, dE partial (gE), A_t, E_t, dX, [dS], x_t.
// writes gE <- dE (for step t-1), a (scratch for gate_wgrad), stem parameter sums, optional dx.
// When t > 0 the kernel continues with (b1)-(b2) of step t-1 on the same patch (dE stays on chip).
template <int CELL, int ACT, bool DX, bool FUSE = true>   // DX: also form d(loss)/d(clip) (viz_model_att.py:189-196); training does not need it
__global__ void __launch_bounds__(128, 4) int_bwd_d_kernel(IntDev d, int t, int first) {
  extern __shared__ __align__(16) float dsm[];
  float* sWaw = dsm;              // [out][in]
  float* sWau = sWaw + C * C;
  float* tiles = sWau + C * C;
  __shared__ __align__(16) float sWp[C * 3], sbp[C];
  __shared__ __align__(16) float sPart[4 * 4 * C], sTot[4 * C];
  __shared__ __align__(16) float sKa[C], sGa[C], sW1[C], sB1[C], sM1[C], sR1[C], sPartA[4 * 4 * C], sTotA[4 * C];
  __shared__ int sIsLast;
  const bool fuse = FUSE && (t > 0);
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  const float* Et = d.E_at(t);
  const float* At = d.A_at(t);
  float* a_out = d.a_at(t);
  int wt = blockIdx.x * 4 + warp;
  BwdDIn cur;
  float xin[3][4];
  // prologue on parameters, forward statistics and the (read-only) clip: overlaps the previous kernel's tail (PDL)
  // two products per tile, one accumulator set at a time: here the mma.sync path pays (measured -3 % on this kernel),
  // unlike in int_bwd_c with its 4-6 products (see there).  PTK_BWD_D_MMA=0 restores the FFMA products.
  constexpr bool MMA = PTK_BWD_D_MMA && (ACT & ACT_FAST) != 0;
  const float* const* gsrc = MMA ? d.gateF : d.gate;
  cta_fill_async(sWaw, gsrc[0], C * C);
  cta_fill_async(sWau, gsrc[1], C * C);
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  if (fuse) bwd_a_fill_const(d, t - 1, sKa, sGa, sW1, sB1, sM1, sR1);
  cp_async_commit();
  for (int i = threadIdx.x; i < 4 * 4 * C; i += blockDim.x) { sPart[i] = 0.f; sPartA[i] = 0.f; }
  if (wt < nWT) {
    const int b = wt / tpi, pix = (wt - b * tpi) * WPX + L.pxo;
    frame_px(d.fv, b, t, pix, d.W, xin);
    const size_t tb = (size_t)b * C * P + (wt - b * tpi) * WPX;   // forward checkpoints of this tile: HBM -> L2
    warp_tile_prefetch(At, tb, P);
    warp_tile_prefetch(Et, tb, P);
    if (fuse) {
      warp_tile_prefetch(d.E_at(t - 1), tb, P);
      warp_tile_prefetch(d.A_at(t - 1), tb, P);
      warp_tile_prefetch(d.Y1_at(t - 1), tb, P);
      warp_tile_prefetch(d.Gex_at(t - 1), tb, P);
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  if (wt < nWT) {     // first tile's operands
    const int b = wt / tpi, pix = (wt - b * tpi) * WPX + L.pxo;
    bwd_d_load<CELL>(d, At, Et, (size_t)b * C * P + pix + (size_t)L.oc0 * P, cur);
  }
  cp_async_wait<0>();
  __syncthreads();
  const BwdAConst ka{sKa, sGa, sW1, sB1, sM1, sR1};
  float* sA = tiles + warp * 2 * TILE_F;
  float* sB = sA + TILE_F;             // parked dE for the fused (b1)-(b2) tail
  float* sPw = sPart + warp * 4 * C;   // [0] d bp, [1..3] d Wp[:, i]
  float* sPwA = sPartA + warp * 4 * C;
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * 4, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    if (!firstTile) {
      bwd_d_load<CELL>(d, At, Et, base + (size_t)L.oc0 * P, cur);
      frame_px(d.fv, b, t, pix, d.W, xin);
    }
#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
      const size_t idx = base + (size_t)(L.oc0 + o) * P;
      BwdDIn nxt;
      if (o < 7) bwd_d_load<CELL>(d, At, Et, idx + P, nxt);     // prefetch the next channel
      const float A[4] = {cur.a.x, cur.a.y, cur.a.z, cur.a.w}, E[4] = {cur.e.x, cur.e.y, cur.e.z, cur.e.w};
      const float gE[4] = {cur.ge.x, cur.ge.y, cur.ge.z, cur.ge.w};
      const float dGe[4] = {cur.dge.x + cur.cg.x, cur.dge.y + cur.cg.y, cur.dge.z + cur.cg.z, cur.dge.w + cur.cg.w};   // + dY0 *7^T W_inh (b9)
      float dS[4] = {0.f, 0.f, 0.f, 0.f}, oa[4], ode[4];
      if (CELL != PTK_CELL_INT) { dS[0] = cur.ds.x; dS[1] = cur.ds.y; dS[2] = cur.ds.z; dS[3] = cur.ds.w; }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        // hGRU: S = A, so dS joins dA;  no_inh: S = E, so dS joins dE
        const float dA = fmaf(dGe[p], E[p], CELL == PTK_CELL_HGRU ? dS[p] : 0.0f);
        oa[p] = dA * A[p] * (1.0f - A[p]);
        ode[p] = fmaf(dGe[p], A[p], gE[p]) + (CELL == CELL_NOINH ? dS[p] : 0.0f);
      }
      if (o < 7) cur = nxt;
      st4s(a_out + idx, oa, d.dy_tstride != 0);
      st4(d.gE + idx, ode);          // parked: completed by the transposed product below
      tile_st(sA, L, o, oa);
    }
    __syncwarp();
    if (MMA) {
      Frag f;                          // one accumulator set at a time
      if (t > 0) {
        frag_zero(f);
        warp_mma_gemm(f, sA, sWau);    // dE += Wa_u^T a   (= dE' of step t-1)
        if (fuse) frag_to_tile(sB, f);                                          // parked in the free tile for the fused tail
        else frag_add_to_global(d.gE, (size_t)b * C * P + pix_base, P, f);      // or completed in place for int_bwd_a_kernel(t-1)
      }
      frag_zero(f);
      warp_mma_gemm(f, sA, sWaw);      // dX += Wa_w^T a
      __syncwarp();
      frag_to_tile(sA, f);             // park dX
      __syncwarp();
      if (fuse) tile_add_global(sB, d.gE, base, P, L);
      tile_add_global(sA, d.dX, base, P, L);
    } else {
      float acc[4][8];
      acc_from_global(acc, d.gE, base, P, L);
      warp_gemm(acc, sA, sWau, L);     // dE += Wa_u^T a   (= dE' of step t-1)
      if (fuse) acc_to_tile(sB, L, acc);
      else if (t > 0) acc_to_global(d.gE, base, P, L, acc);
      acc_from_global(acc, d.dX, base, P, L);
      warp_gemm(acc, sA, sWaw, L);     // dX += Wa_w^T a
      __syncwarp();
      acc_to_tile(sA, L, acc);         // park dX; each thread re-reads only its own patch
    }
    BwdAIn acur, anxt;
    if (fuse) bwd_a_load(d, t - 1, 0, base + (size_t)L.oc0 * P, acur);   // in flight during the stem loop below
    float dxs[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int p = 0; p < 4; ++p) dxs[i][p] = 0.f;
#pragma unroll 1
    for (int o = 0; o < 8; ++o) {
      const int c = L.oc0 + o;
      float pre[4], dxv[4], part[4] = {0.f, 0.f, 0.f, 0.f};
      stem_pre(sWp, sbp, c, xin, pre);
      tile_ld(sA, L, o, dxv);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float dp = dxv[p] * dactf<ACT>(pre[p]);
        part[0] += dp;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          part[1 + i] = fmaf(dp, xin[i][p], part[1 + i]);
          if (DX) dxs[i][p] = fmaf(sWp[c * 3 + i], dp, dxs[i][p]);
        }
      }
      part_add<4>(sPw, c, L, part);
    }
    if (fuse) {              // (b1)-(b2) of step t-1, dE' taken from the parked tile
#pragma unroll 1
      for (int o = 0; o < 8; ++o) {
        const int c = L.oc0 + o;
        const size_t idx = base + (size_t)c * P;
        if (o < 7) bwd_a_load(d, t - 1, 0, idx + P, anxt);          // prefetch next channel
        float gEn[4], part[4] = {0.f, 0.f, 0.f, 0.f};
        tile_ld(sB, L, o, gEn);
        bwd_a_channel<ACT>(d, t - 1, idx, c, gEn, acur, ka, part);
        part_add<4>(sPwA, c, L, part);
        acur = anxt;
      }
    }
    if (DX) {   // contiguous [B,3,T,H,W]
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          dxs[i][p] += __shfl_xor_sync(0xffffffffu, dxs[i][p], 8);
          dxs[i][p] += __shfl_xor_sync(0xffffffffu, dxs[i][p], 16);
        }
      if (L.og == 0) {
        const size_t plane = (size_t)d.T * P;
        float* o = d.dx + (size_t)b * 3 * plane + (size_t)t * P + pix;
        if (t == 0) {
          float g[4];
          ld4v(d.dgoal + (size_t)b * P + pix, g);
#pragma unroll
          for (int p = 0; p < 4; ++p) dxs[2][p] += g[p];
        }
        st4(o, dxs[0]);
        st4(o + plane, dxs[1]);
        st4(o + 2 * plane, dxs[2]);
      }
    }
    __syncwarp();
  }
  part_fold<4>(sPart, sTot);
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    acc_store(acc + ACC_BP + c, sTot[c], first);
#pragma unroll
    for (int i = 0; i < 3; ++i) acc_store(acc + ACC_WP + c * 3 + i, sTot[(1 + i) * C + c], first);
  }
  if (fuse) {
    part_fold<4>(sPartA, sTotA);
    if (threadIdx.x < C) {
      acc_store(acc + ACC_KAPPA + threadIdx.x, sTotA[threadIdx.x], 0);
      acc_store(acc + ACC_GAMMA + threadIdx.x, sTotA[C + threadIdx.x], 0);
    }
    bn_sums_fold(d, 1, sTotA + 2 * C, 0, &sIsLast, sPartA);
  }
}

// ------------------------------------------------------------------------------------------------
// (b10)-(b12) with the tile's operands STAGED in shared memory (InT cell, tensor-core modes, un-fused (b1)-(b2)).
// int_bwd_d_kernel walks a tile channel by channel with a one-channel register prefetch: at B = 64 every warp owns exactly
// one tile, so that kernel's duration is the length of ONE tile's dependent chain of global loads (8 round trips + the
// accumulate passes).  Here a warp requests all five operand tiles of its tile at once (cp.async, 23 KB in flight per warp,
// no registers held), computes from shared memory, adds the Wa_u^T product into the parked dE tile in shared memory
// (instead of a parked global write + a fragment-layout global read-add-write) and fetches dX for the stem while the two
// gate products run.  One 7-warp CTA per SM: 2 048 tiles / (148 x 7) = 1.98, i.e. two full rounds.
// hGRU (S = A, models/ffhgru_hierarchy.py:146-150): dS, completed by int_bwd_c, joins dA -- a sixth operand tile.
constexpr int BDS_NW = 7;
template <int CELL> constexpr int bds_nt() { return CELL == PTK_CELL_HGRU ? 6 : 5; }
template <int CELL> constexpr size_t sm_bwd_ds() { return (size_t)(2 * C * C + BDS_NW * bds_nt<CELL>() * TILE_F) * sizeof(float); }

template <int CELL, int ACT, bool DX>
__global__ void __launch_bounds__(32 * BDS_NW, 1) int_bwd_d_staged_kernel(IntDev d, int t, int first) {
  constexpr int BDS_NT = bds_nt<CELL>();
  extern __shared__ __align__(16) float dsm[];
  float* sWaw = dsm;              // [out][in], mma B-fragment order
  float* sWau = sWaw + C * C;
  float* tiles = sWau + C * C;
  __shared__ __align__(16) float sWp[C * 3], sbp[C];
  __shared__ __align__(16) float sPart[BDS_NW * 4 * C], sTot[4 * C];
  const Lane L;
  const int warp = threadIdx.x >> 5, P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  const float* Et = d.E_at(t);
  const float* At = d.A_at(t);
  float* a_out = d.a_at(t);
  float* tA = tiles + (size_t)warp * BDS_NT * TILE_F;   // A_t -> a = dA A (1 - A) -> parked Wa_w^T a
  float* tE = tA + TILE_F;                              // E_t -> dX of this step
  float* tG = tE + TILE_F;                              // dGe (point-wise part)
  float* tC = tG + TILE_F;                              // dY0 *7^T W_inh (b9)
  float* tg = tC + TILE_F;                              // dE partial -> dE' of step t-1
  float* tS = tg + TILE_F;                              // hGRU only: dS
  int wt = blockIdx.x * BDS_NW + warp;
  float xin[3][4];
  // prologue on parameters, the (read-only) clip and forward checkpoints: overlaps the previous kernel's tail (PDL)
  cta_fill_async(sWaw, d.gateF[0], C * C);
  cta_fill_async(sWau, d.gateF[1], C * C);
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  for (int i = threadIdx.x; i < BDS_NW * 4 * C; i += blockDim.x) sPart[i] = 0.f;
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    frame_px(d.fv, b, t, pix_base + L.pxo, d.W, xin);
    warp_tile_load_async(tA, At, (size_t)d.a_bstride, b, pix_base, P);
    warp_tile_load_async(tE, Et, (size_t)C * P, b, pix_base, P);
  }
  cp_async_commit();
  pdl_wait();
  pdl_launch_dependents();
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    warp_tile_load_async(tG, d.dGe, (size_t)C * P, b, pix_base, P);
    warp_tile_load_async(tC, d.cGe, (size_t)C * P, b, pix_base, P);
    warp_tile_load_async(tg, d.gE, (size_t)C * P, b, pix_base, P);
    if (CELL == PTK_CELL_HGRU) warp_tile_load_async(tS, d.dS, (size_t)C * P, b, pix_base, P);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  float* sPw = sPart + warp * 4 * C;   // [0] d bp, [1..3] d Wp[:, i]
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * BDS_NW, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    if (!firstTile) {
      warp_tile_load_async(tA, At, (size_t)d.a_bstride, b, pix_base, P);
      warp_tile_load_async(tE, Et, (size_t)C * P, b, pix_base, P);
      warp_tile_load_async(tG, d.dGe, (size_t)C * P, b, pix_base, P);
      warp_tile_load_async(tC, d.cGe, (size_t)C * P, b, pix_base, P);
      warp_tile_load_async(tg, d.gE, (size_t)C * P, b, pix_base, P);
      if (CELL == PTK_CELL_HGRU) warp_tile_load_async(tS, d.dS, (size_t)C * P, b, pix_base, P);
      cp_async_commit();
      frame_px(d.fv, b, t, pix, d.W, xin);
      cp_async_wait<0>();
      __syncwarp();
    }
#pragma unroll 2
    for (int o = 0; o < 8; ++o) {
      const size_t idx = base + (size_t)(L.oc0 + o) * P;
      float A[4], E[4], gE[4], dg[4], cg[4], oa[4], ode[4], dS[4] = {0.f, 0.f, 0.f, 0.f};
      tile_ld(tA, L, o, A);
      tile_ld(tE, L, o, E);
      tile_ld(tG, L, o, dg);
      tile_ld(tC, L, o, cg);
      tile_ld(tg, L, o, gE);
      if (CELL == PTK_CELL_HGRU) tile_ld(tS, L, o, dS);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float dGe = dg[p] + cg[p];               // + dY0 *7^T W_inh (b9)
        const float dA = fmaf(dGe, E[p], dS[p]);       // hGRU: S = A, so dS joins dA
        oa[p] = dA * A[p] * (1.0f - A[p]);
        ode[p] = fmaf(dGe, A[p], gE[p]);
      }
      st4s(a_out + idx, oa, d.dy_tstride != 0);
      tile_st(tA, L, o, oa);
      tile_st(tg, L, o, ode);          // parked: completed by the transposed product below
    }
    __syncwarp();
    warp_tile_load_async(tE, d.dX, (size_t)C * P, b, pix_base, P);    // E_t is dead: dX arrives while the gate products run
    cp_async_commit();
    {
      Frag f;                          // one accumulator set at a time
      if (t > 0) {
        frag_zero(f);
        warp_mma_gemm(f, tA, sWau);    // dE += Wa_u^T a   (= dE' of step t-1)
        frag_add_to_tile(tg, f);
      }
      frag_zero(f);
      warp_mma_gemm(f, tA, sWaw);      // dX += Wa_w^T a
      __syncwarp();
      frag_to_tile(tA, f);             // park the product; the stem loop adds the dX tile
    }
    cp_async_wait<0>();
    __syncwarp();
    float dxs[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int p = 0; p < 4; ++p) dxs[i][p] = 0.f;
#pragma unroll 2
    for (int o = 0; o < 8; ++o) {
      const int c = L.oc0 + o;
      float pre[4], dxa[4], dxb[4], part[4] = {0.f, 0.f, 0.f, 0.f};
      if (t > 0) {                     // dE' of step t-1, read by int_bwd_a_kernel(t-1)
        float ge[4];
        tile_ld(tg, L, o, ge);
        st4(d.gE + base + (size_t)c * P, ge);
      }
      stem_pre(sWp, sbp, c, xin, pre);
      tile_ld(tA, L, o, dxa);
      tile_ld(tE, L, o, dxb);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float dp = (dxa[p] + dxb[p]) * dactf<ACT>(pre[p]);
        part[0] += dp;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          part[1 + i] = fmaf(dp, xin[i][p], part[1 + i]);
          if (DX) dxs[i][p] = fmaf(sWp[c * 3 + i], dp, dxs[i][p]);
        }
      }
      part_add<4>(sPw, c, L, part);
    }
    if (DX) {   // contiguous [B,3,T,H,W]
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          dxs[i][p] += __shfl_xor_sync(0xffffffffu, dxs[i][p], 8);
          dxs[i][p] += __shfl_xor_sync(0xffffffffu, dxs[i][p], 16);
        }
      if (L.og == 0) {
        const size_t plane = (size_t)d.T * P;
        float* o = d.dx + (size_t)b * 3 * plane + (size_t)t * P + pix;
        if (t == 0) {
          float g[4];
          ld4v(d.dgoal + (size_t)b * P + pix, g);
#pragma unroll
          for (int p = 0; p < 4; ++p) dxs[2][p] += g[p];
        }
        st4(o, dxs[0]);
        st4(o + plane, dxs[1]);
        st4(o + 2 * plane, dxs[2]);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BDS_NW; ++w) s += sPart[w * 4 * C + i];
    sTot[i] = s;
  }
  __syncthreads();
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    acc_store(acc + ACC_BP + c, sTot[c], first);
#pragma unroll
    for (int i = 0; i < 3; ++i) acc_store(acc + ACC_WP + c * 3 + i, sTot[(1 + i) * C + c], first);
  }
}

// ------------------------------------------------------------------------------------------------
// (b10)-(b12), staged operands, TWO warps per tile (InT cell): the channel split of int_bwd_c_pair_kernel applied to
// int_bwd_d_staged_kernel.  14 warps per SM on the same shared memory.
constexpr int BDP_NS = 7, BDP_NW = 2 * BDP_NS;
template <int CELL> constexpr int bdp_nt() { return CELL == PTK_CELL_HGRU ? 6 : 5; }     // hGRU: a sixth tile for dS
template <int CELL> constexpr size_t sm_bwd_dp() { return (size_t)(2 * C * C + BDP_NS * bdp_nt<CELL>() * TILE_F) * sizeof(float); }

template <int CELL, int ACT, bool DX>
__global__ void __launch_bounds__(32 * BDP_NW, 1) int_bwd_d_pair_kernel(IntDev d, int t, int first) {
  constexpr int BDP_NT = bdp_nt<CELL>();
  extern __shared__ __align__(16) float dsm[];
  float* sWaw = dsm;              // [out][in], mma B-fragment order
  float* sWau = sWaw + C * C;
  float* tiles = sWau + C * C;
  __shared__ __align__(16) float sWp[C * 3], sbp[C];
  __shared__ __align__(16) float sPart[BDP_NW * 4 * C], sTot[4 * C];
  const Lane L;
  const int warp = threadIdx.x >> 5, slot = warp >> 1, h = warp & 1, O0 = 4 * h, N0 = 2 * h;
  const int P = d.H * d.W, tpi = P / WPX, nWT = d.B * tpi;
  const float* Et = d.E_at(t);
  const float* At = d.A_at(t);
  float* a_out = d.a_at(t);
  float* tA = tiles + (size_t)slot * BDP_NT * TILE_F;   // A_t -> a = dA A (1 - A) -> parked Wa_w^T a
  float* tE = tA + TILE_F;                              // E_t -> dX of this step
  float* tG = tE + TILE_F;                              // dGe (point-wise part)
  float* tC = tG + TILE_F;                              // dY0 *7^T W_inh (b9)
  float* tg = tC + TILE_F;                              // dE partial -> dE' of step t-1
  float* tS = tg + TILE_F;                              // hGRU only: dS
  int wt = blockIdx.x * BDP_NS + slot;
  float xin[3][4];
  cta_fill_async(sWaw, d.gateF[0], C * C);
  cta_fill_async(sWau, d.gateF[1], C * C);
  cta_fill_async(sWp, d.Wp, C * 3);
  cta_fill_async(sbp, d.bp, C);
  for (int i = threadIdx.x; i < BDP_NW * 4 * C; i += blockDim.x) sPart[i] = 0.f;
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    frame_px(d.fv, b, t, pix_base + L.pxo, d.W, xin);
    warp_tile_load_async_half(tA, At, (size_t)d.a_bstride, b, pix_base, P, h);
    warp_tile_load_async_half(tE, Et, (size_t)C * P, b, pix_base, P, h);
  }
  cp_async_commit();
  pdl_wait();
  pdl_launch_dependents();
  if (wt < nWT) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX;
    warp_tile_load_async_half(tG, d.dGe, (size_t)C * P, b, pix_base, P, h);
    warp_tile_load_async_half(tC, d.cGe, (size_t)C * P, b, pix_base, P, h);
    warp_tile_load_async_half(tg, d.gE, (size_t)C * P, b, pix_base, P, h);
    if (CELL == PTK_CELL_HGRU) warp_tile_load_async_half(tS, d.dS, (size_t)C * P, b, pix_base, P, h);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  float* sPw = sPart + warp * 4 * C;   // [0] d bp, [1..3] d Wp[:, i]
  for (bool firstTile = true; wt < nWT; wt += gridDim.x * BDP_NS, firstTile = false) {
    const int b = wt / tpi, pix_base = (wt - b * tpi) * WPX, pix = pix_base + L.pxo;
    const size_t base = (size_t)b * C * P + pix;
    if (!firstTile) {
      warp_tile_load_async_half(tA, At, (size_t)d.a_bstride, b, pix_base, P, h);
      warp_tile_load_async_half(tE, Et, (size_t)C * P, b, pix_base, P, h);
      warp_tile_load_async_half(tG, d.dGe, (size_t)C * P, b, pix_base, P, h);
      warp_tile_load_async_half(tC, d.cGe, (size_t)C * P, b, pix_base, P, h);
      warp_tile_load_async_half(tg, d.gE, (size_t)C * P, b, pix_base, P, h);
      if (CELL == PTK_CELL_HGRU) warp_tile_load_async_half(tS, d.dS, (size_t)C * P, b, pix_base, P, h);
      cp_async_commit();
      frame_px(d.fv, b, t, pix, d.W, xin);
      cp_async_wait<0>();
      pair_sync(slot);
    }
#pragma unroll 2
    for (int oo = 0; oo < 4; ++oo) {
      const int o = O0 + oo;
      const size_t idx = base + (size_t)(L.oc0 + o) * P;
      float A[4], E[4], gE[4], dg[4], cg[4], oa[4], ode[4], dS[4] = {0.f, 0.f, 0.f, 0.f};
      tile_ld(tA, L, o, A);
      tile_ld(tE, L, o, E);
      tile_ld(tG, L, o, dg);
      tile_ld(tC, L, o, cg);
      tile_ld(tg, L, o, gE);
      if (CELL == PTK_CELL_HGRU) tile_ld(tS, L, o, dS);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float dGe = dg[p] + cg[p];               // + dY0 *7^T W_inh (b9)
        const float dA = fmaf(dGe, E[p], dS[p]);       // hGRU: S = A, so dS joins dA
        oa[p] = dA * A[p] * (1.0f - A[p]);
        ode[p] = fmaf(dGe, A[p], gE[p]);
      }
      st4s(a_out + idx, oa, d.dy_tstride != 0);
      tile_st(tA, L, o, oa);
      tile_st(tg, L, o, ode);          // parked: completed by the transposed product below
    }
    pair_sync(slot);                   // a complete for all 32 channels; E_t dead
    warp_tile_load_async_half(tE, d.dX, (size_t)C * P, b, pix_base, P, h);    // dX arrives while the gate products run
    cp_async_commit();
    Frag2 fx;
    {
      Frag2 f;
      frag2_zero(f);
      frag2_zero(fx);
      if (t > 0) {                           // a is split once for both products
        warp_mma_gemm_n2x2<true>(f, fx, tA, sWau, sWaw, N0);   // dE += Wa_u^T a (= dE' of step t-1);  dX += Wa_w^T a
        frag2_to_tile<true>(tg, f, N0);
      } else {
        warp_mma_gemm_n2x2<false>(fx, f, tA, sWaw, sWaw, N0);  // dX += Wa_w^T a: parked in tA below, once BOTH warps have read a
      }
    }
    cp_async_wait<0>();
    pair_sync(slot);
    frag2_to_tile<false>(tA, fx, N0);
    pair_sync(slot);
    float dxs[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int p = 0; p < 4; ++p) dxs[i][p] = 0.f;
#pragma unroll 2
    for (int oo = 0; oo < 4; ++oo) {
      const int o = O0 + oo, c = L.oc0 + o;
      float pre[4], dxa[4], dxb[4], part[4] = {0.f, 0.f, 0.f, 0.f};
      if (t > 0) {                     // dE' of step t-1, read by int_bwd_a_kernel(t-1)
        float ge[4];
        tile_ld(tg, L, o, ge);
        st4(d.gE + base + (size_t)c * P, ge);
      }
      stem_pre(sWp, sbp, c, xin, pre);
      tile_ld(tA, L, o, dxa);
      tile_ld(tE, L, o, dxb);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float dp = (dxa[p] + dxb[p]) * dactf<ACT>(pre[p]);
        part[0] += dp;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          part[1 + i] = fmaf(dp, xin[i][p], part[1 + i]);
          if (DX) dxs[i][p] = fmaf(sWp[c * 3 + i], dp, dxs[i][p]);
        }
      }
      part_add<4>(sPw, c, L, part);
    }
    if (DX) {   // contiguous [B,3,T,H,W]: the two warps' channel halves meet through the dead dGe tile
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          dxs[i][p] += __shfl_xor_sync(0xffffffffu, dxs[i][p], 8);
          dxs[i][p] += __shfl_xor_sync(0xffffffffu, dxs[i][p], 16);
        }
      if (h == 1 && L.og == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) st4(tG + i * TP + L.pxo, dxs[i]);
      }
      pair_sync(slot);
      if (h == 0 && L.og == 0) {
        const size_t plane = (size_t)d.T * P;
        float* o = d.dx + (size_t)b * 3 * plane + (size_t)t * P + pix;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float v[4];
          ld4v(tG + i * TP + L.pxo, v);
#pragma unroll
          for (int p = 0; p < 4; ++p) dxs[i][p] += v[p];
        }
        if (t == 0) {
          float g[4];
          ld4v(d.dgoal + (size_t)b * P + pix, g);
#pragma unroll
          for (int p = 0; p < 4; ++p) dxs[2][p] += g[p];
        }
        st4(o, dxs[0]);
        st4(o + plane, dxs[1]);
        st4(o + 2 * plane, dxs[2]);
      }
    }
    pair_sync(slot);                   // the partner is through with the tiles before the next tile's loads land
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BDP_NW; ++w) s += sPart[w * 4 * C + i];
    sTot[i] = s;
  }
  __syncthreads();
  float* acc = d.acc + (size_t)blockIdx.x * ACC_N;
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    acc_store(acc + ACC_BP + c, sTot[c], first);
#pragma unroll
    for (int i = 0; i < 3; ++i) acc_store(acc + ACC_WP + c * 3 + i, sTot[(1 + i) * C + c], first);
  }
}

// ------------------------------------------------------------------------------------------------
// gate weight / bias gradients of one step: six 32x32 outer-product sums over all pixels
//   dWe_w += ge (x) S, dWe_u += ge (x) Ge, dWi_w += gi (x) X, dWi_u += gi (x) S, dWa_w += a (x) X,
//   dWa_u += a (x) E;  d(bias) += sum ge / gi / a.
// 64-pixel tiles, cp.async double-buffered; 384 threads.
//   MMA = false (strict fp32 mode): 6 matrices x 64 threads, 4x4 outputs each, FFMA.
//   MMA = true (tensor-core modes): 12 warps = 6 matrices x 2 row halves; each warp owns a 16 x 32 block of
//     dW as mma.sync m16n8k8 tf32 fragments (K = pixels).  Operands are split x = hi + lo in registers
//     (mask + subtract) and three products hi*hi + hi*lo + lo*hi are issued, so the result carries
//     ~2^-21 relative error per product (same idea as the split-bf16 convolutions).  The [c][GSP] tile
//     layout (GSP = 68 = 4 mod 32) makes every fragment load a conflict-free LDS.32.
constexpr int GPX = 64, GSP = GPX + 4, GT_F = C * GSP;
constexpr int GNST = 3;   // cp.async stages: the all-steps launch streams its 7 operand tensors from HBM, two tiles in flight

// The kernel covers steps t0 .. t0 + nT - 1: one step per launch in the strict-fp32 mode, ALL steps in one launch after the
// reverse loop in the tensor-core modes (the operands of every step are kept, IntDev::geb_at ...), persistent CTAs.
template <int CELL, bool MMA>
__global__ void __launch_bounds__(384, 1) gate_wgrad_kernel(IntDev d, int t0, int nT, int first) {
  extern __shared__ __align__(16) float dsm[];
  __shared__ __align__(8) uint64_t sFull[GNST];      // one mbarrier per stage: 7 x 32 bulk copies of 256 B each
  const int tid = threadIdx.x, P = d.H * d.W, tpi = P / GPX, perT = d.B * tpi, nTiles = nT * perT;
  if (tid == 0) {
    for (int i = 0; i < GNST; ++i) mbar_init(&sFull[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  // a tile = 64 pixels x 32 channels of the 7 operand tensors; row (tensor k, channel c) is 256 contiguous bytes in global
  // memory: warp k < 7 issues tensor k's 32 rows as cp.async.bulk copies (lane = channel).  (The first version issued 3 808
  // 16-byte cp.async per tile with a dynamically indexed pointer table in local memory: the kernel was bound by that
  // address arithmetic, not by HBM.)
  auto issue = [&](int tile, int buf) {
    const int t = t0 + tile / perT, tile_t = tile % perT;
    const int b = tile_t / tpi, pix_base = (tile_t - b * tpi) * GPX;
    const int k = tid >> 5, c = tid & 31;
    if (tid == 0) mbar_expect_tx(&sFull[buf], 7 * C * GPX * 4);
    if (k < 7) {
      const float* sp = k == 0 ? d.geb_at(t) : k == 1 ? d.gi_at(t) : k == 2 ? d.a_at(t)
                      : k == 3 ? ((CELL == PTK_CELL_INT) ? d.I_at(t) : (CELL == PTK_CELL_HGRU ? d.A_at(t) : d.E_at(t)))
                      : k == 4 ? d.GeB_at(t) : k == 5 ? d.X_at(t) : d.E_at(t);
      bulk_g2s(dsm + (size_t)buf * 7 * GT_F + k * GT_F + c * GSP, sp + ((size_t)b * C + c) * P + pix_base, GPX * 4, &sFull[buf]);
    }
  };
  const int m = tid >> 6, r = tid & 63, ocq = r >> 3, icq = r & 7;   // MMA: m = warp / 2 as well
  const int half = (tid >> 5) & 1, fg = (tid & 31) >> 2, ft = tid & 3;   // MMA fragment coordinates
  const int gsel = m >> 1;                       // 0: ge, 1: gi, 2: a
  const int vsel = (m == 0 || m == 3) ? 3 : (m == 1 ? 4 : (m == 5 ? 6 : 5));   // S, Ge, X, S, X, E
  // two-level sums: `acc` / `bsum` hold the current time step, `tot` / `btot` all finished ones (one long fp32 chain
  // over 64 steps x 7 tiles lost ~1e-4 on the hGRU gate gradients at T = 64; per-step partial sums do not)
  // MMA mapping: warp = (matrix m, pixel half ph); it owns the whole 32 x 32 block of dW_m for 32 of the tile's 64 pixels
  // (acc[8][4]: m-tile mt = i >> 2, n-tile nt = i & 3), so a tile's operands are split once per (matrix, pixel) instead of
  // once per (matrix, row half, pixel): 64 instead of 96 splits per lane and tile.  The two pixel halves keep separate slots.
  // FFMA mapping (strict fp32): acc[4][4] as before (rows ocq + 8 i, columns icq + 8 j).
  constexpr int NACC = MMA ? 8 : 4;
  float acc[NACC][4], tot[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }
  float bsum = 0.f, btot = 0.f;
  int buf = 0, t_cur = (int)blockIdx.x / perT, it = 0;
#pragma unroll
  for (int s = 0; s < GNST - 1; ++s)
    if ((int)(blockIdx.x + s * gridDim.x) < nTiles) issue(blockIdx.x + s * gridDim.x, s);
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++it) {
    if (tile / perT != t_cur) {       // a new time step starts: retire the finished one
      t_cur = tile / perT;
#pragma unroll
      for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
      btot += bsum;
      bsum = 0.f;
    }
    const int nxt = tile + (GNST - 1) * gridDim.x;
    if (nxt < nTiles) issue(nxt, (buf + GNST - 1) % GNST);      // that stage was released by the barrier ending the previous iteration
    mbar_wait(&sFull[buf], (uint32_t)((it / GNST) & 1));
    const float* base = dsm + (size_t)buf * 7 * GT_F;
    if (MMA) {
      const float* gA = base + gsel * GT_F + fg * GSP + ft + half * (GPX / 2);      // `half` = pixel half of this warp
      const float* vB = base + vsel * GT_F + fg * GSP + ft + half * (GPX / 2);
      // acc[mt*4 + nt][0..3]: rows 16 mt + fg / + 8, columns 8 nt + 2 ft (+1)
#pragma unroll 2
      for (int ks = 0; ks < GPX / 16; ++ks) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          split_tf32(gA[mt * 16 * GSP + ks * 8], ah[mt][0], al[mt][0]);
          split_tf32(gA[(mt * 16 + 8) * GSP + ks * 8], ah[mt][1], al[mt][1]);
          split_tf32(gA[mt * 16 * GSP + ks * 8 + 4], ah[mt][2], al[mt][2]);
          split_tf32(gA[(mt * 16 + 8) * GSP + ks * 8 + 4], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t bh[2], bl[2];
          split_tf32(vB[nt * 8 * GSP + ks * 8], bh[0], bl[0]);
          split_tf32(vB[nt * 8 * GSP + ks * 8 + 4], bh[1], bl[1]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_tf32(acc[mt * 4 + nt], al[mt], bh);
            mma_tf32(acc[mt * 4 + nt], ah[mt], bl);
            mma_tf32(acc[mt * 4 + nt], ah[mt], bh);
          }
        }
      }
    } else {
    const float* gT = base + gsel * GT_F + ocq * GSP;
    const float* vT = base + vsel * GT_F + icq * GSP;
#pragma unroll 2
    for (int p = 0; p < GPX / 4; ++p) {
      float4 g[4], v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = *reinterpret_cast<const float4*>(gT + i * 8 * GSP + 4 * p);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(vT + j * 8 * GSP + 4 * p);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] = fmaf(g[i].x, v[j].x, fmaf(g[i].y, v[j].y, fmaf(g[i].z, v[j].z, fmaf(g[i].w, v[j].w, acc[i][j]))));
    }
    }
    if (tid < 3 * C) {   // bias sums: row sums of ge / gi / a
      const float4* row = reinterpret_cast<const float4*>(base + (tid >> 5) * GT_F + (tid & 31) * GSP);
#pragma unroll 4
      for (int p = 0; p < GPX / 4; ++p) {
        const float4 v = row[p];
        bsum += (v.x + v.y) + (v.z + v.w);
      }
    }
    __syncthreads();
    buf = (buf + 1) % GNST;
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += tot[i][j];
  bsum += btot;
  float* ga = d.gacc + (size_t)(MMA ? 2 * blockIdx.x + half : blockIdx.x) * GACC_N;
  if (MMA) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        acc_store(ga + m * 1024 + ((i >> 2) * 16 + fg + 8 * (e >> 1)) * C + (i & 3) * 8 + 2 * ft + (e & 1), acc[i][e], first);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_store(ga + m * 1024 + (ocq + 8 * i) * C + icq + 8 * j, acc[i][j], first);
  }
  if (tid < 3 * C) {
    // bias sums are whole-tile row sums: with the two-slot MMA mapping they live in the CTA's even slot, the odd slot holds zero
    float* gb = d.gacc + (size_t)(MMA ? 2 * blockIdx.x : blockIdx.x) * GACC_N + GACC_BE + tid;
    acc_store(gb, bsum, first);
    if (MMA) gb[GACC_N] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// Readout backward (reverse of models/InT.py:236-241); one CTA per sample.
// head partial layout per sample: [0..31] dWr, [32] dbr, [33..82] dWt, [83] dbt, [84] dWd, [85] dbd
constexpr int HD_WR = 0, HD_BR = 32, HD_WT = 33, HD_BT = 83, HD_WD = 84, HD_BD = 85;

__global__ void __launch_bounds__(256) int_readout_bwd_kernel(IntDev d, const float* __restrict__ E) {
  extern __shared__ float sm[];
  const int H = d.H, W = d.W, P = H * W, Wp2 = W + 4, Hp2 = H + 4;
  float* sR = sm;                 // zero-padded readout map r
  float* sG = sm + Hp2 * Wp2;     // zero-padded goal map
  float* sDr = sG + Hp2 * Wp2;    // [P] d r
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
    const float* srcm = ch ? sG : sR;
    float v = 0.f;
    for (int pix = tid; pix < P; pix += 256) v += srcm[(pix / W + ky) * Wp2 + pix % W + kx];
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
// finalize: fold per-CTA / per-sample partials (double) into the flat gradient vector.
struct FinalizeArgs {
  int pw_ctas, gate_ctas, B, wgroups;
  int wgrad_tc;   // > 0: number of CTA slots of the tensor-core weight gradient ([cta][tap][oc][ic])
  int64_t off[PTK_INT_NPARAM + 1];
};

__global__ void __launch_bounds__(256) int_finalize_kernel(IntDev d, FinalizeArgs f) {
  const int64_t total = f.off[PTK_INT_NPARAM];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int p = 0;
    while (e >= f.off[p + 1]) ++p;
    const int j = (int)(e - f.off[p]);
    double s = 0;
    int accoff = -1, gaccoff = -1, headoff = -1;
    switch (p) {
      case PTK_P_PREPROC_W: accoff = ACC_WP + j; break;
      case PTK_P_PREPROC_B: accoff = ACC_BP + j; break;
      case PTK_P_ALPHA: accoff = ACC_ALPHA + j; break;
      case PTK_P_MU: accoff = ACC_MU + j; break;
      case PTK_P_GAMMA: accoff = ACC_GAMMA + j; break;
      case PTK_P_KAPPA: accoff = ACC_KAPPA + j; break;
      case PTK_P_AW_W: gaccoff = GACC_WAW + j; break;
      case PTK_P_AU_W: gaccoff = GACC_WAU + j; break;
      case PTK_P_IW_W: gaccoff = GACC_WIW + j; break;
      case PTK_P_IU_W: gaccoff = GACC_WIU + j; break;
      case PTK_P_EW_W: gaccoff = GACC_WEW + j; break;
      case PTK_P_EU_W: gaccoff = GACC_WEU + j; break;
      case PTK_P_AW_B: case PTK_P_AU_B: gaccoff = GACC_BA + j; break;
      case PTK_P_IW_B: case PTK_P_IU_B: gaccoff = GACC_BI + j; break;
      case PTK_P_EW_B: case PTK_P_EU_B: gaccoff = GACC_BE + j; break;
      case PTK_P_READOUT_W: headoff = HD_WR + j; break;
      case PTK_P_READOUT_B: headoff = HD_BR; break;
      case PTK_P_TARGET_W: headoff = HD_WT + j; break;
      case PTK_P_TARGET_B: headoff = HD_BT; break;
      case PTK_P_DENSE_W: headoff = HD_WD; break;
      case PTK_P_DENSE_B: headoff = HD_BD; break;
      default: break;
    }
    if (accoff >= 0) {
      for (int i = 0; i < f.pw_ctas; ++i) s += (double)d.acc[(size_t)i * ACC_N + accoff];
    } else if (gaccoff >= 0) {
      for (int i = 0; i < f.gate_ctas; ++i) s += (double)d.gacc[(size_t)i * GACC_N + gaccoff];
    } else if (headoff >= 0) {
      for (int i = 0; i < f.B; ++i) s += (double)d.head[(size_t)i * HEAD_N + headoff];
    } else if (p == PTK_P_W_EXC || p == PTK_P_W_INH) {
      // dW[oc][ic][tap] = sum_groups wacc[(grp*4 + ic/8)][tap][oc*8 + ic%8]
      const float* wa = (p == PTK_P_W_EXC) ? d.wacc_exc : d.wacc_inh;
      const int oc = j / (C * TAPS), ic = (j / TAPS) % C, tap = j % TAPS;
      if (f.wgrad_tc > 0) {
        // slots are [cta][tap][oc][ic]: walk the bank in SLOT order (consecutive threads = consecutive oc*32 + ic of one tap, coalesced
        // 128-byte reads per slot) and scatter the one result to dW[oc][ic][tap]; same summation order per element as before
        const int tap2 = j / (C * C), q = j - tap2 * (C * C), oc2 = q / C, ic2 = q - oc2 * C;
        const float* src = wa + (size_t)tap2 * C * C + q;
#pragma unroll 8
        for (int g = 0; g < f.wgrad_tc; ++g) s += (double)__ldcg(src + (size_t)g * TAPS * C * C);
        d.dparams[f.off[p] + ((int64_t)oc2 * C + ic2) * TAPS + tap2] = (float)s;
        continue;
      } else {
        for (int g = 0; g < f.wgroups; ++g)
          s += (double)wa[((size_t)(g * 4 + ic / 8) * TAPS + tap) * 256 + oc * 8 + (ic & 7)];
      }
    } else if (p == PTK_P_BN0_W) s = d.bn_acc[0 * 2 * C + C + j];
    else if (p == PTK_P_BN0_B) s = d.bn_acc[0 * 2 * C + j];
    else if (p == PTK_P_BN1_W) s = d.bn_acc[1 * 2 * C + C + j];
    else if (p == PTK_P_BN1_B) s = d.bn_acc[1 * 2 * C + j];
    d.dparams[e] = (float)s;
  }
}

// ---- host ----------------------------------------------------------------------------------
static constexpr size_t SM_BWD_C = (size_t)(4 * C * C + 4 * 2 * TILE_F) * sizeof(float);
static constexpr size_t SM_BWD_D = (size_t)(2 * C * C + 4 * 2 * TILE_F) * sizeof(float);
static constexpr size_t SM_GATE = (size_t)(GNST * 7 * GT_F) * sizeof(float);

template <int CELL, int ACT>
static int run_backward_t(const IntDev& d, cudaStream_t st) {
  const int P = d.H * d.W;
  const int grid = d.pw_ctas;
  const int grid_c = grid;
  const bool tc = (d.conv_mode != PTK_CONV_FP32);
  const int grid_a = grid;
  PTK_CUDA_OK(cudaMemsetAsync(d.acc, 0, (size_t)PW_EW_CTAS * ACC_N * sizeof(float), st));
  if (CELL == CELL_NOINH) {   // buffers the skipped inhibition branch would have written
    const size_t Sb = (size_t)d.B * C * P * sizeof(float);
    PTK_CUDA_OK(cudaMemsetAsync(d.cGe, 0, Sb, st));
    PTK_CUDA_OK(cudaMemsetAsync(d.dX, 0, Sb, st));
    PTK_CUDA_OK(cudaMemsetAsync(d.gi_at(0), 0, d.dy_tstride ? Sb * d.T : Sb, st));
    PTK_CUDA_OK(cudaMemsetAsync(d.bn_acc, 0, 4 * C * sizeof(double), st));
    PTK_CUDA_OK(cudaMemsetAsync(d.wacc_inh, 0, (tc ? wgrad_ts_acc_floats(d.T, d.B, d.H, d.W) : (size_t)4 * WGRAD_GROUPS * TAPS * 256) * sizeof(float), st));
  }
  const long long gate_tiles = (long long)(d.B * P / GPX) * (tc ? d.T : 1);
  const int gate_grid = gate_tiles < GATE_CTAS ? (int)gate_tiles : GATE_CTAS;
  PTK_CUDA_OK(ptk_func_smem(int_bwd_c_kernel<CELL, ACT>, (int)SM_BWD_C));
  {
    constexpr int SC = CELL == PTK_CELL_HGRU ? PTK_CELL_HGRU : PTK_CELL_INT;
    PTK_CUDA_OK(ptk_func_smem(int_bwd_c_staged_kernel<SC, ACT>, (int)sm_bwd_cs<SC>()));
    PTK_CUDA_OK(ptk_func_smem(int_bwd_c_pair_kernel<SC, ACT>, (int)sm_bwd_cp<SC>()));
  }
  PTK_CUDA_OK(ptk_func_smem(int_bwd_d_kernel<CELL, ACT, true>, (int)SM_BWD_D));
  PTK_CUDA_OK(ptk_func_smem(int_bwd_d_kernel<CELL, ACT, false>, (int)SM_BWD_D));
  PTK_CUDA_OK(ptk_func_smem(int_bwd_d_kernel<CELL, ACT, false, false>, (int)SM_BWD_D));
  PTK_CUDA_OK(ptk_func_smem(int_bwd_d_kernel<CELL, ACT, true, false>, (int)SM_BWD_D));
  {
    constexpr int SC = CELL == PTK_CELL_HGRU ? PTK_CELL_HGRU : PTK_CELL_INT;
    PTK_CUDA_OK(ptk_func_smem(int_bwd_d_pair_kernel<SC, ACT, true>, (int)sm_bwd_dp<SC>()));
    PTK_CUDA_OK(ptk_func_smem(int_bwd_d_pair_kernel<SC, ACT, false>, (int)sm_bwd_dp<SC>()));
    PTK_CUDA_OK(ptk_func_smem(int_bwd_d_staged_kernel<SC, ACT, true>, (int)sm_bwd_ds<SC>()));
    PTK_CUDA_OK(ptk_func_smem(int_bwd_d_staged_kernel<SC, ACT, false>, (int)sm_bwd_ds<SC>()));
  }
  PTK_CUDA_OK(ptk_func_smem(gate_wgrad_kernel<CELL, true>, (int)SM_GATE));
  PTK_CUDA_OK(ptk_func_smem(gate_wgrad_kernel<CELL, false>, (int)SM_GATE));
  const size_t smr = (size_t)(2 * (d.H + 4) * (d.W + 4) + P) * sizeof(float);
  PTK_CUDA_OK(ptk_func_smem(int_readout_bwd_kernel, (int)smr));
  {
    ProfileScope ps(PTK_PROF_MISC, st);
    int_readout_bwd_kernel<<<d.B, 256, smr, st>>>(d, d.E_at(d.T));
  }
  PTK_LAUNCH_OK("int_readout_bwd_kernel");
  const int ew_blocks = 148 * 4;
  PTK_CUDA_OK(ptk_func_smem(bn_apply_kernel, 0));
  for (int t = d.T - 1; t >= 0; --t) {
    const int first = (t == d.T - 1);
    int rc;
    // (b1)-(b2) as a launch of their own (default), or fused into the tail of the previous iteration's int_bwd_d_kernel
    // (PTK_BWD_FUSE_A=1: the round-1 structure, kept for A/B runs)
    static const bool unfuse = [] { const char* e = getenv("PTK_BWD_FUSE_A"); return !(e && e[0] == '1'); }();
    if (first || unfuse) {
      ProfileScope ps(PTK_PROF_POINT_BWD, st);
      static const bool a_stream = [] { const char* e = getenv("PTK_BWD_A_STREAM"); return !(e && e[0] == '0'); }();
      // the streaming kernel keeps the step's d(kappa), d(gamma) in accumulator slot 0, the warp-tile kernel one partial per
      // CTA slot: one or the other for ALL steps of a call (a_stream && unfuse), never mixed
      if (a_stream && unfuse) PTK_CUDA_OK(launch_pdl(!first, int_bwd_a_stream_kernel<ACT>, dim3(((d.B + bas_rows(d.B) - 1) / bas_rows(d.B)) * C), dim3(256), 0, st, d, t, first, bas_rows(d.B)));
      else PTK_CUDA_OK(launch_pdl(!first, int_bwd_a_kernel<ACT>, dim3(grid_a), dim3(128), 0, st, d, t, first));
      PTK_LAUNCH_OK("int_bwd_a_kernel");
    }
    if (!d.bn_fused) {
      ProfileScope ps(PTK_PROF_BN_APPLY, st);
      PTK_CUDA_OK(launch_pdl(true, bn_apply_kernel, dim3(ew_blocks), dim3(256), 0, st, d, (const float*)d.Y1_at(t), t, 1));
      PTK_LAUNCH_OK("bn_apply_kernel");
    }
    ConvArgs c1{d.dV_at(t, 1), d.w_exc_d, d.cgI, d.B, d.H, d.W, 0, nullptr, nullptr, nullptr, d.conv_mode, 1};
    if (d.bn_fused) {   // (b3) inside the convolution's staging; dY1 of this step lands in its per-step store
      c1.bn_y = d.Y1_at(t); c1.bn_stat = d.stat(t, 1); c1.bn_tot = d.bn_tot; c1.bn_w = d.bn_w[1]; c1.bn_out = d.dAct_at(t, 1);
    }
    c1.v2 = (d.conv_mode != PTK_CONV_FP32) && conv_v2_enabled();
    { ProfileScope ps(PTK_PROF_CONV_DGRAD, st); rc = launch_conv7(c1, st); }
    if (rc) return rc;
    if (!tc) {   // strict fp32: per-step CUDA-core weight gradient (the tensor-core modes defer it, see below)
      WgradArgs w1{d.I_at(t + 1), d.dAct_at(t, 1), d.wacc_exc, d.B, d.H, d.W, first, d.conv_mode};
      { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgrad7(w1, WGRAD_GROUPS, st); }
      if (rc) return rc;
    }
    {
      ProfileScope ps(PTK_PROF_POINT_BWD, st);
      static const bool c_staged_on = [] { const char* e = getenv("PTK_BWD_C_STAGED"); return !(e && e[0] == '0'); }();
      static const bool c_pair_on = [] { const char* e = getenv("PTK_BWD_C_PAIR"); return !(e && e[0] == '0'); }();
      // two warps per tile: measured faster at 32x32 (16.56 -> 16.34 ms at B = 64) and at 64x64 (36.53 -> 36.08 ms at B = 16, T = 128)
      if ((CELL == PTK_CELL_INT || CELL == PTK_CELL_HGRU) && (ACT & ACT_FAST) != 0 && c_staged_on && c_pair_on) {
        constexpr int SC = CELL == PTK_CELL_HGRU ? PTK_CELL_HGRU : PTK_CELL_INT;
        constexpr int NS = bcp_ns<SC>();
        const int nWT = d.B * P / WPX, sg = (nWT + NS - 1) / NS;
        PTK_CUDA_OK(launch_pdl(true, int_bwd_c_pair_kernel<SC, ACT>, dim3(sg < 148 ? sg : 148), dim3(64 * NS), sm_bwd_cp<SC>(), st, d, t, first));
      } else
      if ((CELL == PTK_CELL_INT || CELL == PTK_CELL_HGRU) && (ACT & ACT_FAST) != 0 && c_staged_on) {   // operands staged in shared memory, one CTA per SM
        constexpr int SC = CELL == PTK_CELL_HGRU ? PTK_CELL_HGRU : PTK_CELL_INT;
        constexpr int NW = bcs_nw<SC>();
        const int nWT = d.B * P / WPX, sg = (nWT + NW - 1) / NW;
        PTK_CUDA_OK(launch_pdl(true, int_bwd_c_staged_kernel<SC, ACT>, dim3(sg < 148 ? sg : 148), dim3(32 * NW), sm_bwd_cs<SC>(), st, d, t, first));
      } else {
        PTK_CUDA_OK(launch_pdl(true, int_bwd_c_kernel<CELL, ACT>, dim3(grid_c), dim3(128), SM_BWD_C, st, d, t, first));
      }
    }
    PTK_LAUNCH_OK("int_bwd_c_kernel");
    if (CELL != CELL_NOINH) {
      if (!d.bn_fused) {
        ProfileScope ps(PTK_PROF_BN_APPLY, st);
        PTK_CUDA_OK(launch_pdl(true, bn_apply_kernel, dim3(ew_blocks), dim3(256), 0, st, d, (const float*)d.Y0_at(t), t, 0));
        PTK_LAUNCH_OK("bn_apply_kernel");
      }
      ConvArgs c0{d.dV_at(t, 0), d.w_inh_d, d.cGe, d.B, d.H, d.W, 0, nullptr, nullptr, nullptr, d.conv_mode, 1};
      if (d.bn_fused) {   // (b9)
        c0.bn_y = d.Y0_at(t); c0.bn_stat = d.stat(t, 0); c0.bn_tot = d.bn_tot; c0.bn_w = d.bn_w[0]; c0.bn_out = d.dAct_at(t, 0);
      }
    c0.v2 = (d.conv_mode != PTK_CONV_FP32) && conv_v2_enabled();
      { ProfileScope ps(PTK_PROF_CONV_DGRAD, st); rc = launch_conv7(c0, st); }
      if (rc) return rc;
      if (!tc) {
        WgradArgs w0{d.GeB_at(t), d.dAct_at(t, 0), d.wacc_inh, d.B, d.H, d.W, first, d.conv_mode};
        { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgrad7(w0, WGRAD_GROUPS, st); }
        if (rc) return rc;
      }
    }
    {
      ProfileScope ps(PTK_PROF_POINT_BWD, st);
      static const bool staged_on = [] { const char* e = getenv("PTK_BWD_D_STAGED"); return !(e && e[0] == '0'); }();
      constexpr bool can_stage = (CELL == PTK_CELL_INT || CELL == PTK_CELL_HGRU) && (ACT & ACT_FAST) != 0 && PTK_BWD_D_MMA;
      static const bool d_pair_on = [] { const char* e = getenv("PTK_BWD_D_PAIR"); return !(e && e[0] == '0'); }();
      if (can_stage && unfuse && staged_on && d_pair_on) {   // staged operands, two warps per tile
        constexpr int SC = CELL == PTK_CELL_HGRU ? PTK_CELL_HGRU : PTK_CELL_INT;
        const int nWT = d.B * P / WPX, sg = (nWT + BDP_NS - 1) / BDP_NS;
        const dim3 g2(sg < 148 ? sg : 148);
        if (d.dx != nullptr) PTK_CUDA_OK(launch_pdl(true, int_bwd_d_pair_kernel<SC, ACT, true>, g2, dim3(32 * BDP_NW), sm_bwd_dp<SC>(), st, d, t, first));
        else PTK_CUDA_OK(launch_pdl(true, int_bwd_d_pair_kernel<SC, ACT, false>, g2, dim3(32 * BDP_NW), sm_bwd_dp<SC>(), st, d, t, first));
      } else
      if (can_stage && unfuse && staged_on) {   // operands staged in shared memory, one 7-warp CTA per SM
        constexpr int SC = CELL == PTK_CELL_HGRU ? PTK_CELL_HGRU : PTK_CELL_INT;
        const int nWT = d.B * P / WPX, sg = (nWT + BDS_NW - 1) / BDS_NW;
        const dim3 g2(sg < 148 ? sg : 148);
        if (d.dx != nullptr) PTK_CUDA_OK(launch_pdl(true, int_bwd_d_staged_kernel<SC, ACT, true>, g2, dim3(32 * BDS_NW), sm_bwd_ds<SC>(), st, d, t, first));
        else PTK_CUDA_OK(launch_pdl(true, int_bwd_d_staged_kernel<SC, ACT, false>, g2, dim3(32 * BDS_NW), sm_bwd_ds<SC>(), st, d, t, first));
      } else
      if (d.dx != nullptr && unfuse) PTK_CUDA_OK(launch_pdl(true, int_bwd_d_kernel<CELL, ACT, true, false>, dim3(grid), dim3(128), SM_BWD_D, st, d, t, first));
      else if (d.dx != nullptr) PTK_CUDA_OK(launch_pdl(true, int_bwd_d_kernel<CELL, ACT, true, true>, dim3(grid), dim3(128), SM_BWD_D, st, d, t, first));
      else if (unfuse) PTK_CUDA_OK(launch_pdl(true, int_bwd_d_kernel<CELL, ACT, false, false>, dim3(grid), dim3(128), SM_BWD_D, st, d, t, first));
      else PTK_CUDA_OK(launch_pdl(true, int_bwd_d_kernel<CELL, ACT, false, true>, dim3(grid), dim3(128), SM_BWD_D, st, d, t, first));
    }
    PTK_LAUNCH_OK("int_bwd_d_kernel");
    if (!tc) {
      ProfileScope ps(PTK_PROF_GATE_WGRAD, st);
      PTK_CUDA_OK(launch_pdl(true, gate_wgrad_kernel<CELL, false>, dim3(gate_grid), dim3(384), SM_GATE, st, d, t, 1, first));
      PTK_LAUNCH_OK("gate_wgrad_kernel");
    }
  }
  int gate_slots = gate_grid;
  if (tc) {   // the six gate weight gradients of ALL steps in one launch
    ProfileScope ps(PTK_PROF_GATE_WGRAD, st);
    // tcgen05 version (ptk_gate_wgrad_tc.cu: one GEMM per 16 pixels, HBM-bound) or the mma.sync tf32 version (PTK_GATE_WGRAD_TC=0)
    static const bool gw_tc = [] { const char* e = getenv("PTK_GATE_WGRAD_TC"); return !(e && e[0] == '0'); }();
    if (gw_tc) {
      gate_slots = launch_gate_wgrad_tc(d, CELL, d.T, st);
      if (gate_slots < 0) return PTK_ECUDA;
    } else {
      PTK_CUDA_OK(launch_pdl(true, gate_wgrad_kernel<CELL, true>, dim3(gate_grid), dim3(384), SM_GATE, st, d, 0, d.T, 1));
      gate_slots = 2 * gate_grid;
    }
    PTK_LAUNCH_OK("gate_wgrad_kernel");
  }
  if (tc) {
    // the two 7x7 weight gradients, each ONE launch over all T steps: dW_exc = sum_t I_{t+1} (star) dY1_t,
    // dW_inh = sum_t Ge_t (star) dY0_t (accumulators stay in TMEM for the whole kernel, ptk_wgrad_ts.cu)
    int rc;
    WgradTsArgs w1{d.I_at(1), (long long)d.ck_stride, d.dYs[1], (long long)d.dy_tstride, d.wacc_exc, d.T, d.B, d.H, d.W, d.conv_mode, 1};
    { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgrad7_ts(w1, st); }
    if (rc) return rc;
    if (CELL != CELL_NOINH) {
      WgradTsArgs w0{d.GeT, (long long)d.dy_tstride, d.dYs[0], (long long)d.dy_tstride, d.wacc_inh, d.T, d.B, d.H, d.W, d.conv_mode, 1};
      { ProfileScope ps(PTK_PROF_WGRAD, st); rc = launch_wgrad7_ts(w0, st); }
      if (rc) return rc;
    }
  }
  FinalizeArgs f;
  f.pw_ctas = PW_EW_CTAS; f.gate_ctas = gate_slots; f.B = d.B; f.wgroups = WGRAD_GROUPS;
  f.wgrad_tc = tc ? wgrad_ts_ctas(d.T, d.B, d.H, d.W) : 0;
  const ParamOffsets po = param_offsets(C, KS);
  for (int i = 0; i <= PTK_INT_NPARAM; ++i) f.off[i] = po.o[i];
  {
    ProfileScope ps(PTK_PROF_MISC, st);
    int_finalize_kernel<<<148 * 2, 256, 0, st>>>(d, f);
  }
  PTK_LAUNCH_OK("int_finalize_kernel");
  return PTK_OK;
}

template <int CELL>
static int run_backward_c(const IntDev& d, int act, cudaStream_t st) {
  const bool fast = (d.conv_mode != PTK_CONV_FP32);
  if (act == PTK_ACT_SOFTPLUS) return fast ? run_backward_t<CELL, PTK_ACT_SOFTPLUS | ACT_FAST>(d, st)
                                           : run_backward_t<CELL, PTK_ACT_SOFTPLUS>(d, st);
  return fast ? run_backward_t<CELL, PTK_ACT_TANH | ACT_FAST>(d, st) : run_backward_t<CELL, PTK_ACT_TANH>(d, st);
}

int run_int_backward(const IntDev& d, int cell, int act, cudaStream_t st) {
  if (cell == CELL_NOINH) return run_backward_c<CELL_NOINH>(d, act, st);
  return cell == PTK_CELL_INT ? run_backward_c<PTK_CELL_INT>(d, act, st) : run_backward_c<PTK_CELL_HGRU>(d, act, st);
}

}  // namespace ptk
