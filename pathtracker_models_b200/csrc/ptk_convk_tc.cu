// Tensor-core convolution for the wide recurrent kernels of the ConvLSTM baseline (models/convlstm.py:84-90:
// four 25 -> 25 filt_size x filt_size convolutions of h, executed as ONE stacked 25 -> 100 convolution; and its
// transpose 100 -> 25 in the reverse-time loop): implicit GEMM on tcgen05 / TMEM, fp32 parity through the same
// 3-pass bf16 split as ptk_conv_tc.cu (In = Ih + Il, W = Wh + Wl, Y ~= Ih*Wh + Ih*Wl + Il*Wh, fp32 accumulation).
//
// Per CTA: 8 rows x 32 columns of one sample, all output channels.
//   * operand image: the zero-padded input tile (8 + KS - 1 rows, pitch 32 + KS - 1 pixels) of ONE 32-channel block
//     of the input, bf16 hi / lo, channel-chunk planar [chunk = ic/8][pixel][8 ic] = the K-major no-swizzle UMMA
//     canonical layout (SBO = 128 B, LBO = plane), so a tap is the same image addressed at start pixel ky*pitch + kx.
//     Inputs wider than 32 channels (the transposed convolution reads 100) are walked block by block: the MMA warp
//     hands the image back through an mbarrier (tcgen05.commit) and the staging warps refill it; accumulators stay in
//     TMEM across blocks.
//   * GEMM per MMA: M = 128 consecutive padded pixels (3 M-tiles cover 8 x pitch), K = 16 channels, and
//       forward  (NSTACK = 1): N = 112 >= 100 stacked gate channels -- one tap per instruction; A 4 KB + B 3.5 KB of
//                shared-memory operands per 56 tensor-pipe cycles: balanced, no stacking needed;
//       transpose(NSTACK = 3): N = 96 = three horizontally adjacent taps x 32 (>= 25) output channels: column block
//                dl holds the contribution to output pixel m - dl; the epilogue realigns through a shared-memory
//                tile (the operand image is dead by then).
//   * weights stream through a cp.async.bulk + mbarrier ring in "units" (one B operand: NT rows x 32 k, hi then lo);
//     warp 0 = producer, warp 1 = MMA issuer (one elected lane), warps 2-7 = staging; all 8 warps run the epilogue.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "ptk_common.cuh"
#include "ptk_kernels.h"
#include "ptk_tc.cuh"

namespace ptk {

template <int KS_, int NT_, int NSTACK_, int NBUF_ = 1>
struct ConvK {
  static constexpr int KSZ = KS_, NT = NT_, NSTACK = NSTACK_;
  static constexpr int NBUF = NBUF_;   // operand images: 2 = the next 32-channel block is staged while this one is multiplied
  static constexpr int PAD = KSZ / 2;
  static constexpr int TH = 8, TW = 32;
  static constexpr int WP = TW + KSZ - 1;                   // padded row pitch (pixels)
  static constexpr int ROWS = TH + KSZ - 1;
  static constexpr int MT = (TH * WP + 127) / 128;          // M-tiles
  static constexpr int NPIX = (((KSZ - 1) * WP + (KSZ - 1) + MT * 128) + 7) & ~7;   // last pixel an MMA can touch, rounded
  static constexpr int PLANE = NPIX * 16;                   // bytes per channel-chunk plane
  static constexpr int A_BYTES = 2 * 4 * PLANE;             // hi + lo, 4 chunks (32 channels)
  static constexpr int UPR = KSZ / NSTACK;                  // units per kernel row
  static constexpr int UNITS = KSZ * UPR;
  static constexpr int UNIT_BYTES = 2 * 4 * NT * 16;        // [Wh: 4 chunks x NT rows x 16 B][Wl: same]
  static constexpr int G = (UPR * UNIT_BYTES <= 32768) ? UPR : 1;   // units per ring stage (a whole kernel row if it is small)
  static constexpr int STAGE = G * UNIT_BYTES;
  static constexpr int NST = (STAGE * 6 + NBUF * A_BYTES + 512 <= 232448) ? 6 : ((STAGE * 4 + NBUF * A_BYTES + 512 <= 232448) ? 4 : 3);
  static constexpr int SMEM = NBUF * A_BYTES + NST * STAGE + 256;
  static constexpr int DSTRIDE = (NT + 31) & ~31;           // TMEM columns between M-tiles: accumulators start on 32-column boundaries
  static constexpr int DCOLS = MT * DSTRIDE;
  static constexpr int TCOLS = DCOLS <= 128 ? 128 : (DCOLS <= 256 ? 256 : 512);
  static constexpr int SP = MT * 128 + 4;                   // pitch (floats) of the realignment tile
  static_assert(KSZ % NSTACK == 0, "taps per row must split into whole stacks");
  static_assert(UNITS % G == 0, "ring stages must hold whole units");
  static_assert(NT % 16 == 0 && NT <= 256, "UMMA N constraint for M = 128");
  static_assert(DCOLS <= 512, "TMEM columns");
  static_assert(SMEM <= 232448, "shared memory");
  static_assert(32 * SP * 4 <= A_BYTES, "epilogue tile must fit in the operand image");
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <class K>
__global__ void __launch_bounds__(256, 1) convk_tc_kernel(ConvKArgs a) {
  if (a.skip != nullptr && a.skip[1] != 0.f) return;   // RBP: the Neumann series has terminated (uniform for the grid)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sA = smem_raw;
  uint8_t* sWr = smem_raw + K::NBUF * K::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + K::NBUF * K::A_BYTES + K::NST * K::STAGE);
  uint64_t* full = bars;                 // [NST]
  uint64_t* empty = bars + K::NST;       // [NST]
  uint64_t* a_full = bars + 2 * K::NST;  // [2] operand image of block kb staged (192 arrivals)
  uint64_t* a_free = a_full + 2;         // [2] MMAs of block kb have retired (tcgen05.commit)
  uint64_t* accum_bar = a_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 5);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int x0 = blockIdx.x * K::TW, y0 = blockIdx.y * K::TH, b = blockIdx.z;
  const int H = a.H, W = a.W;
  const int nblk = (a.Cin + 31) / 32;
  const size_t HW = (size_t)H * W;

  if (tid == 0) {
    for (int i = 0; i < K::NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&a_full[0], 192); mbar_init(&a_full[1], 192);
    mbar_init(&a_free[0], 1); mbar_init(&a_free[1], 1);
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(K::TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < K::NBUF * K::A_BYTES / 16; i += 256) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int nstg = nblk * (K::UNITS / K::G);   // ring stages in the whole weight stream

  if (warp == 0) {
    // ---- weight producer -----------------------------------------------------------------------------------
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wT);
      for (int s = 0; s < nstg; ++s) {
        const int st = s % K::NST, round = s / K::NST;
        if (round > 0) mbar_wait(&empty[st], (uint32_t)((round - 1) & 1));
        mbar_expect_tx(&full[st], K::STAGE);
        bulk_g2s(sWr + st * K::STAGE, wsrc + (size_t)s * K::STAGE, K::STAGE, &full[st]);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------
    const uint32_t aBase = smem_u32(sA), wBase = smem_u32(sWr);
    const uint64_t dA0 = umma_desc(0, K::PLANE, 128);
    const uint64_t dB0 = umma_desc(0, K::NT * 16, 128);
    const uint32_t aHiW = (uint32_t)(dA0 >> 32), aLoW = (uint32_t)dA0;
    const uint32_t bHiW = (uint32_t)(dB0 >> 32), bLoW = (uint32_t)dB0;
    constexpr uint32_t IDESC = umma_idesc(K::NT);
    int s_glob = 0;
    for (int kb = 0; kb < nblk; ++kb) {
      const int ab = kb % K::NBUF;
      mbar_wait(&a_full[ab], (uint32_t)((kb / K::NBUF) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aHi = aBase + (uint32_t)ab * K::A_BYTES, aLo = aHi + 4 * K::PLANE;
      for (int sb = 0; sb < K::UNITS / K::G; ++sb, ++s_glob) {
        const int st = s_glob % K::NST, round = s_glob / K::NST;
        mbar_wait(&full[st], (uint32_t)(round & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stBase = wBase + (uint32_t)st * K::STAGE;
#pragma unroll 1
        for (int g = 0; g < K::G; ++g) {
          const int unit = sb * K::G + g, ky = unit / K::UPR, u = unit - ky * K::UPR;
          const uint32_t poff = (uint32_t)(ky * K::WP + u * K::NSTACK);        // start pixel == 16-byte units
          const uint32_t bH = bLoW + ((stBase + (uint32_t)g * K::UNIT_BYTES) >> 4);
          const uint32_t bL = bH + (uint32_t)((64 * K::NT) >> 4);
          const bool first_unit = (kb == 0 && unit == 0);
#pragma unroll
          for (int j = 0; j < K::MT; ++j) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              if (kb * 32 + s * 16 >= a.Cin) continue;   // all-zero K-step (channels 112-127 of the 100-channel transpose)
              const uint32_t po = poff + (uint32_t)(j * 128);
              const uint32_t aH = aLoW + ((aHi + (uint32_t)(2 * s) * K::PLANE) >> 4) + po;
              const uint32_t aL = aLoW + ((aLo + (uint32_t)(2 * s) * K::PLANE) >> 4) + po;
              const uint32_t kadv = (uint32_t)(s * ((2 * K::NT * 16) >> 4));
              const uint64_t dBh = ((uint64_t)bHiW << 32) | (bH + kadv), dBl = ((uint64_t)bHiW << 32) | (bL + kadv);
              const uint32_t dcol = tmem_base + (uint32_t)(j * K::DSTRIDE);
              if (a.dbg != 2 && elect_one()) {
                umma_bf16(dcol, ((uint64_t)aHiW << 32) | aH, dBh, IDESC, (first_unit && s == 0) ? 0u : 1u);
                if (!a.single) {   // the two cross terms of the split; conv_mode = bf16 keeps the leading product only
                  umma_bf16(dcol, ((uint64_t)aHiW << 32) | aH, dBl, IDESC, 1u);
                  umma_bf16(dcol, ((uint64_t)aHiW << 32) | aL, dBh, IDESC, 1u);
                }
              }
            }
          }
        }
        __syncwarp();
        if (elect_one()) umma_commit(&empty[st]);
      }
      __syncwarp();
      if (kb + 1 == nblk) { if (elect_one()) umma_commit(accum_bar); }
      else if (kb + K::NBUF < nblk) { if (elect_one()) umma_commit(&a_free[ab]); }
    }
  } else {
    // ---- staging warps: fp32 NCHW -> bf16 hi/lo channel-chunk planar, one 32-channel block at a time ---------
    const int stid = tid - 64;
    for (int kb = 0; kb < nblk; ++kb) {
      const int ab = kb % K::NBUF;
      if (kb >= K::NBUF) mbar_wait(&a_free[ab], (uint32_t)((kb / K::NBUF - 1) & 1));
      uint8_t* sA = smem_raw + ab * K::A_BYTES;      // shadows the CTA-wide pointer: this block's image
      const float* inb = a.in + (size_t)b * a.Cin * HW;
      // interior columns: item = (chunk, row, 4-pixel group); 8 channel loads of 16 B -> 4 pixels
      for (int it = stid; it < 4 * K::ROWS * (K::TW / 4); it += 192) {
        const int q = it / (K::ROWS * (K::TW / 4)), rem = it - q * (K::ROWS * (K::TW / 4)), row = rem / (K::TW / 4), g = rem % (K::TW / 4);
        const int gy = y0 + row - K::PAD, ch0 = kb * 32 + q * 8;
        const bool rowok = gy >= 0 && gy < H;
        float4 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rowok && ch0 + e < a.Cin) v[e] = *reinterpret_cast<const float4*>(inb + (size_t)(ch0 + e) * HW + (size_t)gy * W + x0 + 4 * g);
        }
        const int p = row * K::WP + K::PAD + 4 * g;
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = reinterpret_cast<const float*>(&v[2 * e])[px], v1 = reinterpret_cast<const float*>(&v[2 * e + 1])[px];
            split_bf16x2(v0, v1, hi[e], lo[e]);
          }
          *reinterpret_cast<uint4*>(sA + q * K::PLANE + (p + px) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(sA + 4 * K::PLANE + q * K::PLANE + (p + px) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      // halo columns: zero outside the image; inside it only when W > 32
      for (int it = stid; it < 4 * K::ROWS * 2 * K::PAD; it += 192) {
        const int q = it / (K::ROWS * 2 * K::PAD), rem = it - q * (K::ROWS * 2 * K::PAD), row = rem / (2 * K::PAD), hc = rem % (2 * K::PAD);
        const int c = hc < K::PAD ? hc : (K::TW + hc);
        const int gy = y0 + row - K::PAD, gx = x0 + c - K::PAD, ch0 = kb * 32 + q * 8;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v0 = 0.f, v1 = 0.f;
          if (ok && ch0 + 2 * e < a.Cin) v0 = inb[(size_t)(ch0 + 2 * e) * HW + (size_t)gy * W + gx];
          if (ok && ch0 + 2 * e + 1 < a.Cin) v1 = inb[(size_t)(ch0 + 2 * e + 1) * HW + (size_t)gy * W + gx];
          split_bf16x2(v0, v1, hi[e], lo[e]);
        }
        const int p = row * K::WP + c;
        *reinterpret_cast<uint4*>(sA + q * K::PLANE + p * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sA + 4 * K::PLANE + q * K::PLANE + p * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
      mbar_arrive(&a_full[ab]);
    }
  }
  __syncwarp();

  // ---- epilogue (all 8 warps) --------------------------------------------------------------------------------
  mbar_wait(accum_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int quad = warp & 3, jpar = warp >> 2;
  float* outb = a.out + (size_t)b * a.Cout * HW;
  const float* addb = a.addend ? a.addend + (size_t)b * a.Cout * HW : nullptr;
  // TMEM -> [channel][padded pixel] tile in the (dead) operand image -> coalesced 16-byte global accesses, CH channels per
  // pass.  Stacked taps (NSTACK = 3) are realigned on the way: out[m] = D0[m] + D1[m+1] + D2[m+2], one phase per block.
  // (A first version stored straight from the TMEM registers, one 4-byte element per lane and channel: 108 us per CTA.)
  {
    float* sS = reinterpret_cast<float*>(sA);
    constexpr int NCH = K::NT / K::NSTACK;
    constexpr int CH = (NCH > 32 && K::A_BYTES >= 64 * K::SP * 4) ? 64 : 32;
#pragma unroll 1
    for (int c0 = 0; c0 < a.Cout && a.dbg != 1; c0 += CH) {
#pragma unroll 1
      for (int dl = 0; dl < K::NSTACK; ++dl) {
#pragma unroll 1
        for (int j = jpar; j < K::MT; j += 2) {
          const int m = j * 128 + quad * 32 + lane - dl;
#pragma unroll 1
          for (int cb = 0; cb < CH / 16; ++cb) {
            if (c0 + cb * 16 >= NCH) break;   // warp-uniform
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * K::DSTRIDE + dl * NCH + c0 + cb * 16), v);
            if (m >= 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float* p = sS + (cb * 16 + i) * K::SP + m;
                *p = dl ? (*p + v[i]) : v[i];
              }
            }
          }
        }
        __syncthreads();
      }
      const int nc = min(CH, a.Cout - c0);
      for (int it = tid; it < nc * K::TH * (K::TW / 4); it += 256) {
        const int n = it / (K::TH * (K::TW / 4)), rem = it - n * (K::TH * (K::TW / 4)), row = rem / (K::TW / 4), g = rem % (K::TW / 4);
        const float* src = sS + n * K::SP + row * K::WP + 4 * g;
        float4 r = make_float4(src[0], src[1], src[2], src[3]);
        const size_t off = (size_t)(c0 + n) * HW + (size_t)(y0 + row) * W + x0 + 4 * g;
        if (addb) { const float4 t = __ldg(reinterpret_cast<const float4*>(addb + off)); r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
        if (a.accumulate) { const float4 t = *reinterpret_cast<const float4*>(outb + off); r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
        if (a.bias) { const float t = __ldg(a.bias + c0 + n); r.x += t; r.y += t; r.z += t; r.w += t; }
        *reinterpret_cast<float4*>(outb + off) = r;
      }
      __syncthreads();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(K::TCOLS));
}

// Weight stream for convk_tc_kernel, written destination-first (pads are zeros): for block kb (32 K-channels), unit
// (ky, u): [Wh: chunk = k/8][row = dl*(NT/NSTACK) + n][k%8] then [Wl: same].
//   transposed = 0 (forward):   n = bank*ocb + oc (stacked gates), k = ic,             tap (ky, kx = u*NSTACK + dl)
//   transposed = 1 (data grad): n = ic,                            k = bank*ocb + oc,  tap flipped
struct PrepKArgs {
  const float* w[4];   // banks [ocb][icb][KS*KS]
  int nbank, ocb, icb, ks, nt, nstack, nblk, transposed;
  __nv_bfloat16* dst;
};
__global__ void prep_weights_k_kernel(PrepKArgs p) {
  const int upr = p.ks / p.nstack, units = p.ks * upr, nch = p.nt / p.nstack, taps = p.ks * p.ks;
  const size_t per_unit = (size_t)2 * 4 * p.nt * 8, total = (size_t)p.nblk * units * per_unit;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e & 7);
    size_t r = e >> 3;
    const int row = (int)(r % p.nt); r /= p.nt;
    const int chunk = (int)(r & 3); r >>= 2;
    const int lo = (int)(r & 1); r >>= 1;
    const int unit = (int)(r % units), kb = (int)(r / units);
    const int ky = unit / upr, u = unit - ky * upr, dl = row / nch, n = row - dl * nch, kx = u * p.nstack + dl;
    const int kg = kb * 32 + chunk * 8 + kk;
    float v = 0.f;
    if (!p.transposed) {
      if (n < p.nbank * p.ocb && kg < p.icb) v = p.w[n / p.ocb][((size_t)(n % p.ocb) * p.icb + kg) * taps + ky * p.ks + kx];
    } else {
      if (n < p.icb && kg < p.nbank * p.ocb)
        v = p.w[kg / p.ocb][((size_t)(kg % p.ocb) * p.icb + n) * taps + (p.ks - 1 - ky) * p.ks + (p.ks - 1 - kx)];
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    p.dst[e] = lo ? l : h;
  }
}

// generic entry points (ConvGRU, ptk_gru.cu): dense [B,Cin,H,W] -> [B,Cout,H,W], one tap per instruction, N = NT >= Cout
size_t convk_tc_stream_bytes(int ks, int nt, int nblk) { return (size_t)nblk * ks * ks * 2 * 4 * nt * 16; }
int launch_prep_weights_k_generic(const float* const* w, int nbank, int ocb, int icb, int ks, int nt, int nblk, int transposed,
                                  void* dst, cudaStream_t st) {
  PrepKArgs p;
  for (int i = 0; i < 4; ++i) p.w[i] = i < nbank ? w[i] : nullptr;
  p.nbank = nbank; p.ocb = ocb; p.icb = icb; p.ks = ks; p.nt = nt; p.nstack = 1; p.nblk = nblk; p.transposed = transposed;
  p.dst = (__nv_bfloat16*)dst;
  prep_weights_k_kernel<<<148 * 4, 256, 0, st>>>(p);
  PTK_LAUNCH_OK("prep_weights_k_kernel");
  return PTK_OK;
}
template <int KSZ, int NT>
int launch_convk_tc_n(const ConvKArgs& a, cudaStream_t st) {
  using X = ConvK<KSZ, NT, 1, 2>;    // 7x7: two operand images fit, re-staging overlaps the MMAs
  if (a.H % 8 || a.W % 32 || a.Cout > NT) { set_error("convk_tc_n: unsupported problem H=%d W=%d Cout=%d", a.H, a.W, a.Cout); return PTK_EINVAL; }
  PTK_CUDA_OK(ptk_func_smem(convk_tc_kernel<X>, X::SMEM));
  convk_tc_kernel<X><<<dim3(a.W / 32, a.H / 8, a.B), 256, X::SMEM, st>>>(a);
  PTK_LAUNCH_OK("convk_tc_kernel");
  return PTK_OK;
}
template int launch_convk_tc_n<7, 128>(const ConvKArgs&, cudaStream_t);
template int launch_convk_tc_n<7, 64>(const ConvKArgs&, cudaStream_t);

template <int KSZ>
size_t convk_tc_weight_bytes(int transposed) {
  using F = ConvK<KSZ, 112, 1>;
  using D = ConvK<KSZ, (KSZ % 3 == 0) ? 96 : 32, (KSZ % 3 == 0) ? 3 : 1>;
  return transposed ? (size_t)4 * D::UNITS * D::UNIT_BYTES : (size_t)F::UNITS * F::UNIT_BYTES;
}
template size_t convk_tc_weight_bytes<7>(int);
template size_t convk_tc_weight_bytes<15>(int);

template <int KSZ>
int launch_prep_weights_k(const float* const* w, int transposed, void* dst, cudaStream_t st) {
  using F = ConvK<KSZ, 112, 1>;
  using D = ConvK<KSZ, (KSZ % 3 == 0) ? 96 : 32, (KSZ % 3 == 0) ? 3 : 1>;
  PrepKArgs p;
  for (int i = 0; i < 4; ++i) p.w[i] = w[i];
  p.nbank = 4; p.ocb = 25; p.icb = 25; p.ks = KSZ; p.transposed = transposed;
  p.nt = transposed ? D::NT : F::NT;
  p.nstack = transposed ? D::NSTACK : F::NSTACK;
  p.nblk = transposed ? 4 : 1;
  p.dst = (__nv_bfloat16*)dst;
  prep_weights_k_kernel<<<148 * 4, 256, 0, st>>>(p);
  PTK_LAUNCH_OK("prep_weights_k_kernel");
  return PTK_OK;
}
template int launch_prep_weights_k<7>(const float* const*, int, void*, cudaStream_t);
template int launch_prep_weights_k<15>(const float* const*, int, void*, cudaStream_t);

// forward: Cin <= 32 (25), Cout <= 112 (100);  transposed: Cin <= 128 (100), Cout <= 32 (25)
template <int KSZ>
int launch_convk_tc(const ConvKArgs& a_in, int transposed, cudaStream_t st) {
  ConvKArgs a = a_in;
  static const int exp_d = getenv("PTK_CONVK_DBG") ? atoi(getenv("PTK_CONVK_DBG")) : 0;   // phase timing: 1 = no epilogue, 2 = no MMAs
  a.dbg = exp_d;
  using F = ConvK<KSZ, 112, 1>;
  using D = ConvK<KSZ, (KSZ % 3 == 0) ? 96 : 32, (KSZ % 3 == 0) ? 3 : 1>;
  if (a.H % 8 || a.W % 32) { set_error("convk_tc: H %% 8 / W %% 32 violated (%d x %d)", a.H, a.W); return PTK_EINVAL; }
  dim3 grid(a.W / 32, a.H / 8, a.B);
  if (!transposed) {
    if (a.Cin > 32 || a.Cout > F::NT) { set_error("convk_tc forward: %d -> %d channels unsupported", a.Cin, a.Cout); return PTK_EINVAL; }
    PTK_CUDA_OK(ptk_func_smem(convk_tc_kernel<F>, F::SMEM));
    convk_tc_kernel<F><<<grid, 256, F::SMEM, st>>>(a);
  } else {
    if (a.Cin > 128 || a.Cout > D::NT / D::NSTACK) { set_error("convk_tc transposed: %d -> %d channels unsupported", a.Cin, a.Cout); return PTK_EINVAL; }
    PTK_CUDA_OK(ptk_func_smem(convk_tc_kernel<D>, D::SMEM));
    convk_tc_kernel<D><<<grid, 256, D::SMEM, st>>>(a);
  }
  PTK_LAUNCH_OK("convk_tc_kernel");
  return PTK_OK;
}
template int launch_convk_tc<7>(const ConvKArgs&, int, cudaStream_t);
template int launch_convk_tc<15>(const ConvKArgs&, int, cudaStream_t);

}  // namespace ptk
