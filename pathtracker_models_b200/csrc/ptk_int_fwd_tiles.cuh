// Per-tile bodies of the forward point-wise stages (eqs (1)-(13) of SURVEY.md section 8; reference: models/InT.py:145-179,
// models/ffhgru_hierarchy.py:135-173).  Shared by the per-phase kernels of the PDL chain (ptk_int_fwd.cu) and the persistent
// time-loop kernel (ptk_int_loop.cu), so that both execute the same instruction sequence per tile (bit-identical results).
#pragma once
#include "ptk_int.h"
#include "ptk_pw.cuh"

namespace ptk {

constexpr int TILE_F = C * TP;   // floats per warp tile

// stem of the thread's patch into the X tile (rolled: softplus is ~100 instructions per element)
template <int ACT>
__device__ __forceinline__ void stem_to_tile(float* sX, const float* sWp, const float* sbp, const float (&xin)[3][4],
                                             const Lane& L) {
#pragma unroll 1
  for (int o = 0; o < 8; ++o) {
    float pre[4];
    stem_pre(sWp, sbp, L.oc0 + o, xin, pre);
#pragma unroll
    for (int p = 0; p < 4; ++p) pre[p] = actf<ACT>(pre[p]);
    tile_st(sX, L, o, pre);
  }
}

// eqs (1)-(2) for frame t: A = sigmoid(Wa_w X + Wa_u E + b), Ge = A*E.  sE must hold E_t.
// NOINH (InT(no_inh=True), models/InT.py:167-168): inhibition := Ge, gated_inhibition := E, so the step's
// excitation gate G_e = sigmoid(We_w E + We_u Ge + b) (eq. 9) is formed here as well and Ge becomes I_{t+1}.
template <int ACT, bool NOINH>
__device__ __forceinline__ void attention_tile(const IntDev& d, int t, int b, int pix_base, float* sX,
                                               const float* sE, const float* sWaw, const float* sWau,
                                               const float* sba, const float* sWp, const float* sbp,
                                               const float* sWew, const float* sWeu, const float* sbe,
                                               const Lane& L, const float (&xin)[3][4]) {
  const int P = d.H * d.W, pix = pix_base + L.pxo;
  stem_to_tile<ACT>(sX, sWp, sbp, xin, L);
  __syncwarp();
  constexpr bool MMA = (ACT & ACT_FAST) != 0;   // tensor-core modes: weights are in B-fragment order, bias added at read
  float acc[4][8];
  if (MMA) {
    Frag f;
    frag_zero(f);
    warp_mma_gemm(f, sX, sWaw);
    warp_mma_gemm(f, sE, sWau);
    __syncwarp();
    frag_to_tile(sX, f);
    __syncwarp();
  } else {
    acc_fill(acc, sba, L);
    warp_gemm(acc, sX, sWaw, L);
    warp_gemm(acc, sE, sWau, L);
    __syncwarp();
    acc_to_tile(sX, L, acc);          // park the pre-activations; only this thread re-reads its patch
  }
  float* Aout = d.A_at(t) + (size_t)b * d.a_bstride + pix;
  float* Geout = d.Ge + (size_t)b * C * P + pix;
  float* Iout = d.I_at(t + 1) + (size_t)b * C * P + pix;
#pragma unroll 1
  for (int o = 0; o < 8; ++o) {
    float e[4], a[4], ge[4];
    tile_ld(sE, L, o, e);
    tile_ld(sX, L, o, a);
    const float bias = MMA ? sba[L.oc0 + o] : 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) { a[p] = sigm<ACT>(a[p] + bias); ge[p] = a[p] * e[p]; }
    st4(Aout + (size_t)(L.oc0 + o) * P, a);
    st4(Geout + (size_t)(L.oc0 + o) * P, ge);
    if (NOINH) {
      st4(Iout + (size_t)(L.oc0 + o) * P, ge);
      tile_st(sX, L, o, ge);
    }
  }
  if (NOINH) {
    __syncwarp();
    if (MMA) {
      Frag f;
      frag_zero(f);
      warp_mma_gemm(f, sE, sWew);
      warp_mma_gemm(f, sX, sWeu);
      __syncwarp();
      frag_to_tile(sX, f);          // Ge (in sX) is dead now
      __syncwarp();
      acc_fill(acc, sbe, L);
      acc_add_tile(acc, sX, L);
    } else {
      acc_fill(acc, sbe, L);
      warp_gemm(acc, sE, sWew, L);
      warp_gemm(acc, sX, sWeu, L);
    }
    float* Gx = d.Gex_at(t) + (size_t)b * C * P + pix;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float g[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) g[p] = sigm<ACT>(acc[p][o]);
      st4(Gx + (size_t)(L.oc0 + o) * P, g);
    }
  }
}

// ---- mid stage: eqs (3b)-(9): BatchNorm0 apply, inhibition drive, inhibition gate + blend, excitation gate -----------------
struct MidIn { float4 s, i, y0; };
template <int CELL>
__device__ __forceinline__ void mid_load(const float* Sg, const float* It, const float* Y0, size_t base, int c, int P,
                                         MidIn& r) {
  r.s = ld4(Sg + (size_t)c * P);
  if (CELL != PTK_CELL_INT) r.i = ld4(It + base + (size_t)c * P);
  r.y0 = ld4(Y0 + base + (size_t)c * P);
}
// shared-memory operands of the mid stage (weights in [in][out] or B-fragment order, per-channel vectors)
struct MidW {
  const float *Wiw, *Wiu, *Wew, *Weu, *Wp, *bp, *bi, *be, *Al, *Mu, *Scale, *Shift;
};
// Everything of a tile that does NOT depend on BatchNorm0's batch statistics: the excitation gate (eq. 9, written to its
// checkpoint), the stem X (left in sGX) and the inhibition-gate pre-activation (eq. 7, parked in sS); `cur` = channel 0 of the
// blend loop, in flight.  sS must hold S (I_t for InT, A_t for hGRU), sGX must hold Ge_t.
template <int CELL, int ACT>
__device__ __forceinline__ void mid_tile_head(const IntDev& d, int t, size_t base, const float* Sg, float* sS, float* sGX,
                                              const MidW& w, const Lane& L, const float (&xin)[3][4], MidIn& cur,
                                              const float* It, const float* Y0) {
  const int P = d.H * d.W;
  constexpr bool MMA = (ACT & ACT_FAST) != 0;
  float acc[4][8];
  // excitation gate (eq. 9) first: G_e = sigmoid(We_w S + We_u Ge + b), written straight to the checkpoint
  if (MMA) {
    Frag f;
    frag_zero(f);
    warp_mma_gemm(f, sS, w.Wew);
    warp_mma_gemm(f, sGX, w.Weu);
    __syncwarp();
    frag_to_tile(sGX, f);          // the Ge tile is dead once both products have read it
    __syncwarp();
    acc_fill(acc, w.be, L);
    acc_add_tile(acc, sGX, L);
  } else {
    acc_fill(acc, w.be, L);
    warp_gemm(acc, sS, w.Wew, L);
    warp_gemm(acc, sGX, w.Weu, L);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float gx[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) gx[p] = sigm<ACT>(acc[p][o]);
    st4(d.Gex_at(t) + base + (size_t)(L.oc0 + o) * P, gx);
  }
  __syncwarp();
  mid_load<CELL>(Sg, It, Y0, base, L.oc0, P, cur);   // channel 0 of the blend loop: in flight during the products below
  stem_to_tile<ACT>(sGX, w.Wp, w.bp, xin, L);     // the Ge tile is free now: X takes its place
  __syncwarp();
  // inhibition gate pre-activation (eq. 7), parked in the S tile (S is re-read from global below)
  if (MMA) {
    Frag f;
    frag_zero(f);
    warp_mma_gemm(f, sGX, w.Wiw);
    warp_mma_gemm(f, sS, w.Wiu);
    __syncwarp();
    frag_to_tile(sS, f);           // bias is added where the tile is read
    __syncwarp();
  } else {
    acc_fill(acc, w.bi, L);
    warp_gemm(acc, sGX, w.Wiw, L);
    warp_gemm(acc, sS, w.Wiu, L);
    __syncwarp();
    acc_to_tile(sS, L, acc);
  }
}
// The BatchNorm0-dependent rest: U = BN0(Y0), eqs (4)-(8); writes G_i and I_{t+1}.  w.Scale / w.Shift must be valid.
template <int CELL, int ACT>
__device__ __forceinline__ void mid_tile_tail(const IntDev& d, int t, size_t base, const float* Sg, const float* sS,
                                              const float* sGX, const MidW& w, const Lane& L, MidIn& cur,
                                              const float* It, const float* Y0, float* Iout) {
  const int P = d.H * d.W;
  constexpr bool MMA = (ACT & ACT_FAST) != 0;
#pragma unroll 1
  for (int o = 0; o < 8; ++o) {
    const int c = L.oc0 + o;
    const size_t idx = base + (size_t)c * P;
    MidIn nxt;
    if (o < 7) mid_load<CELL>(Sg, It, Y0, base, c + 1, P, nxt);
    float X[4], gi[4], out[4];
    tile_ld(sGX, L, o, X);
    tile_ld(sS, L, o, gi);
    const float S[4] = {cur.s.x, cur.s.y, cur.s.z, cur.s.w}, y0[4] = {cur.y0.x, cur.y0.y, cur.y0.z, cur.y0.w};
    float I[4];
    if (CELL == PTK_CELL_INT) {
#pragma unroll
      for (int p = 0; p < 4; ++p) I[p] = S[p];
    } else {
      I[0] = cur.i.x; I[1] = cur.i.y; I[2] = cur.i.z; I[3] = cur.i.w;
    }
    const float gbias = MMA ? w.bi[c] : 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      gi[p] = sigm<ACT>(gi[p] + gbias);
      const float U = fmaf(y0[p], w.Scale[c], w.Shift[c]);
      const float Pp = U * fmaf(w.Al[c], S[p], w.Mu[c]);
      const float Q = actf<ACT>(Pp);
      const float Ihat = actf<ACT>(X[p] - Q);
      out[p] = (1.0f - gi[p]) * I[p] + gi[p] * Ihat;
    }
    st4(d.Gi_at(t) + idx, gi);
    st4(Iout + idx, out);
    if (o < 7) cur = nxt;
  }
}

// ---- post stage: eqs (10b)-(13) for step t, then (1)-(2) for step t+1 (if any) ----------------------------------------------
struct PostIn { float4 y1, i2, e, gx; };
__device__ __forceinline__ void post_load(const float* Y1, const float* I2, const float* Et, const float* Gx,
                                          size_t idx, PostIn& r) {
  r.y1 = ld4(Y1 + idx); r.i2 = ld4(I2 + idx); r.e = ld4(Et + idx); r.gx = ld4(Gx + idx);
}
struct PostW {
  const float *Waw, *Wau, *Wew, *Weu, *Wp, *bp, *ba, *be, *Ka, *Ga, *Scale, *Shift, *Wr;
};
// `cur` = the operands of the tile's first channel (post_load of channel L.oc0), already in flight; xin = frame t+1.
template <int ACT, bool NOINH>
__device__ __forceinline__ void post_tile(const IntDev& d, int t, int has_next, int b, int pix_base, float* sX, float* sE,
                                          const PostW& w, float br, const Lane& L, const float (&xin)[3][4], PostIn& cur) {
  const int P = d.H * d.W, pix = pix_base + L.pxo;
  const size_t base = (size_t)b * C * P + pix;
  const float* Et = d.E_at(t);
  const float* I2 = d.I_at(t + 1);
  float* Eout = d.E_at(t + 1);
  const float* Y1 = d.Y1_at(t);
  const float* Gx = d.Gex_at(t);
  float sread[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int o = 0; o < 8; ++o) {
    const int c = L.oc0 + o;
    const size_t idx = base + (size_t)c * P;
    PostIn nxt;
    if (o < 7) post_load(Y1, I2, Et, Gx, idx + P, nxt);
    const float y1[4] = {cur.y1.x, cur.y1.y, cur.y1.z, cur.y1.w}, i2[4] = {cur.i2.x, cur.i2.y, cur.i2.z, cur.i2.w};
    const float e[4] = {cur.e.x, cur.e.y, cur.e.z, cur.e.w}, gx[4] = {cur.gx.x, cur.gx.y, cur.gx.z, cur.gx.w};
    float e2[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float V = fmaf(y1[p], w.Scale[c], w.Shift[c]);
      const float R = V * fmaf(w.Ka[c], i2[p], w.Ga[c]);
      const float Ehat = actf<ACT>(R);
      e2[p] = (1.0f - gx[p]) * e[p] + gx[p] * Ehat;
      sread[p] = fmaf(w.Wr[c], e2[p], sread[p]);
    }
    st4(Eout + idx, e2);
    tile_st(sE, L, o, e2);
    if (o < 7) cur = nxt;
  }
  if (d.states != nullptr) {   // testmode: states.append(readout_conv(excitation)), InT.py:233
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      sread[p] += __shfl_xor_sync(0xffffffffu, sread[p], 8);
      sread[p] += __shfl_xor_sync(0xffffffffu, sread[p], 16);
      sread[p] += br;
    }
    if (L.og == 0) st4(d.states + ((size_t)b * d.T + t) * P + pix, sread);
  }
  if (has_next) attention_tile<ACT, NOINH>(d, t + 1, b, pix_base, sX, sE, w.Waw, w.Wau, w.ba, w.Wp, w.bp, w.Wew, w.Weu, w.be, L, xin);
}

}  // namespace ptk
