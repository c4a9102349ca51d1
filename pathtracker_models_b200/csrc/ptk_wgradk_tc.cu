// Tensor-core weight gradient of the ConvLSTM's recurrent convolution (tcgen05 / TMEM):
//
//   dWh[n = gate*25 + oc][ic][ky][kx] += sum_{b,y,x} delta[b,n,y,x] * h[b,ic,y+ky-P,x+kx-P]      (autograd of models/convlstm.py:84-87)
//
// GEMM per tap: D[m = n (128 >= 100)][ic (32 >= 25)] = sum_{k = pixel} delta[k][m] * h[k + shift(tap)][ic]; pixels are the
// K axis, both operands MN-major views of channel-chunk planar tiles ([8-channel plane][pixel][8 ch], ptk_wgrad_tc.cu).
// The gradient of one tap is a 128 x 32 fp32 block = 32 TMEM columns, so one kernel ROW (KS taps) fits in TMEM and the
// whole gradient (KS rows) does not.  Hence the work split: a CTA owns ONE kernel row ky and a strided share of the
// pixel tiles (4 rows x 32 columns of one sample), keeps its KS accumulators in TMEM across all of its tiles and writes
// them once at the end (read-add-write of its private slot: deterministic, no atomics).  For a fixed ky the h tile needs
// no row halo (4 rows shifted by ky - P), so a tile's operands are small (88 KB) and double-buffered: warps 2-7 stage
// tile i+1 (fp32 -> bf16 hi/lo split) while warp 1 issues the MMAs of tile i.
//   per (tap PAIR, 16-pixel K-step): three M = 128, N = 64 instructions  dh*hh + dh*hl + dl*hh  (3-pass split, as
//   everywhere).  N = 64 = [tap kx | tap kx+1]: the h tile is stored twice, the second copy shifted by one pixel, so that
//   the second 32 columns of the B operand are the neighbouring tap (an M = 128, N = 32 instruction is bound by the 4 KB
//   A read; the pair costs 6 KB instead of 2 x 5 KB).
#include <cuda_bf16.h>

#include "ptk_common.cuh"
#include "ptk_kernels.h"
#include "ptk_tc.cuh"

namespace ptk {

constexpr int WGK_NSTAGE_THREADS = 224;   // warps 0, 2-7 stage; warp 1 issues the MMAs
constexpr int WGK_NITEM = 3;              // 16-byte-vector items per staging pass: 24 LDG.128 in flight per thread

template <int KS_>
struct WgK {
  static constexpr int KSZ = KS_, PAD = KS_ / 2;
  static constexpr int TH = 4, TW = 32, KPIX = TH * TW, KSTEPS = KPIX / 16;
  static constexpr int WP = TW + KSZ - 1;                     // padded pitch of the h tile
  static constexpr int DY_PLANE = KPIX * 16 + 16;             // +16 B: the planes of one K row start in different bank groups
  static constexpr int IN_NPIX = TH * WP + 2;                 // +2: the odd partner tap of the last pair reads one pixel further
  static constexpr int IN_PLANE = IN_NPIX * 16 + 16;
  static constexpr int DY_BYTES = 32 * DY_PLANE;              // 16 hi planes (128 channels) + 16 lo planes
  static constexpr int IN_BYTES = 16 * IN_PLANE;              // [hh | hh shifted by one pixel | hl | hl shifted], 4 planes each
  static constexpr int BUF = DY_BYTES + IN_BYTES;
  static constexpr int SMEM = 2 * BUF + 128;
  static constexpr int NPAIR = (KSZ + 1) / 2;                 // taps are processed in horizontally adjacent pairs (N = 64)
  static constexpr int TCOLS = NPAIR * 64 <= 256 ? 256 : 512;
  static_assert(NPAIR * 64 <= 512, "one kernel row of accumulators must fit in TMEM");
  static_assert(BUF % 16 == 0 && SMEM <= 232448, "shared memory");
};

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both MN-major, M = 128, N = 64
__host__ __device__ constexpr uint32_t wgk_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <class K>
__global__ void __launch_bounds__(256, 1) wgradk_tc_kernel(WgradKArgs a) {
  if (a.skip != nullptr && a.skip[1] != 0.f) return;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * K::BUF);
  uint64_t* full = bars;           // [2] tile staged (one arrival per staging thread)
  uint64_t* empty = bars + 2;      // [2] MMAs of the tile retired (tcgen05.commit)
  uint64_t* accum_bar = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int ky = blockIdx.y, R = gridDim.x, r0 = blockIdx.x;
  const int H = a.H, W = a.W, tilesX = W / K::TW, tilesY = H / K::TH, nTiles = a.B * tilesX * tilesY;
  const size_t HW = (size_t)H * W;
  const int myTiles = r0 < nTiles ? (nTiles - r0 + R - 1) / R : 0;

  if (tid == 0) {
    const int nfull = WGK_NSTAGE_THREADS + (a.dyc ? 1 : 0);   // + the expect_tx arrival of the bulk copy
    mbar_init(&full[0], nfull); mbar_init(&full[1], nfull);
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1);
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(K::TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 2 * K::BUF / 16; i += 256) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the bulk copies (async proxy) overwrite these bytes
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------
    constexpr uint32_t IDESC = wgk_idesc();
    for (int i = 0; i < myTiles; ++i) {
      const int buf = i & 1;
      mbar_wait(&full[buf], (uint32_t)((i >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sDY = smem_u32(smem_raw + buf * K::BUF), sIN = sDY + K::DY_BYTES;
      // MN-major no-swizzle descriptors: LBO = 128 B between 8-pixel K groups, SBO = plane stride between 8-channel groups
      const uint64_t dAh = umma_desc(sDY, 128, K::DY_PLANE), dAl = umma_desc(sDY + 16 * K::DY_PLANE, 128, K::DY_PLANE);
      // B = 8 planes from hh (resp. hl) on: columns 0-31 = tap kx, columns 32-63 = the shifted copy = tap kx + 1
      const uint64_t dBh = umma_desc(sIN, 128, K::IN_PLANE), dBl = umma_desc(sIN + 8 * K::IN_PLANE, 128, K::IN_PLANE);
#pragma unroll 1
      for (int pr = 0; pr < K::NPAIR; ++pr) {
        const uint32_t dcol = tmem_base + (uint32_t)(pr * 64);
#pragma unroll 2
        for (int ks = 0; ks < K::KSTEPS; ++ks) {
          // K-step = 16 consecutive pixels of one tile row: delta compact (pitch 32), h padded (pitch WP) shifted by kx
          const uint32_t ao = (uint32_t)(ks * 16), bo = (uint32_t)((ks >> 1) * K::WP + (ks & 1) * 16 + 2 * pr);
          if (elect_one()) {
            umma_bf16(dcol, dAh + ao, dBh + bo, IDESC, (i | ks) ? 1u : 0u);
            if (!a.single) {   // conv_mode = bf16 keeps the leading product only
              umma_bf16(dcol, dAh + ao, dBl + bo, IDESC, 1u);
              umma_bf16(dcol, dAl + ao, dBh + bo, IDESC, 1u);
            }
          }
        }
      }
      __syncwarp();
      if (elect_one()) umma_commit(&empty[buf]);
    }
    __syncwarp();
    if (elect_one()) umma_commit(accum_bar);
  } else {
    // ---- staging warps (0, 2-7) -------------------------------------------------------------------------------
    const int stid = warp == 0 ? lane : tid - 32;
    for (int i = 0; i < myTiles; ++i) {
      const int buf = i & 1;
      if (i >= 2) mbar_wait(&empty[buf], (uint32_t)(((i >> 1) - 1) & 1));
      const int tile = r0 + i * R;
      const int b = tile / (tilesX * tilesY), trem = tile - b * (tilesX * tilesY), y0 = (trem / tilesX) * K::TH, x0 = (trem % tilesX) * K::TW;
      uint8_t* sDY = smem_raw + buf * K::BUF;
      uint8_t* sIN = sDY + K::DY_BYTES;
      // one list of 16-byte-vector items: delta (planes 0..ceil(Cout/8)-1, 4 rows x 8 four-pixel groups; channels >= Cout
      // zero) followed by the interior of the h tile (4 planes; rows shifted by ky - PAD; zero outside the image / beyond
      // Cin).  WGK_NITEM items per pass, all loads requested before the first conversion (L2-latency bound).
      const float* dyb = a.dy + (size_t)b * a.Cout * HW;
      const float* inb = a.in + (size_t)b * (a.in_bs ? (size_t)a.in_bs : (size_t)a.Cin * HW);
      // item -> (plane group of 4, row) x lane; inside a warp the lane picks (plane q = 4*group + lane%4, 4-pixel group g):
      // the 8 lanes of a quarter warp then hit 8 different 16-byte bank groups with their STS.128 (plane strides are
      // 16 resp. 48 mod 128 bytes, the two g of a quarter warp are 64 bytes apart) and still read whole 32-byte sectors.
      const int nDY = ((a.Cout + 31) / 32) * K::TH * 32, nIt = nDY + K::TH * 32;
      if (a.dyc != nullptr && stid == 0) {    // pre-split delta tile: one bulk copy, completes on the same barrier
        mbar_expect_tx(&full[buf], K::DY_BYTES);
        bulk_g2s(sDY, reinterpret_cast<const uint8_t*>(a.dyc) + (size_t)tile * K::DY_BYTES, K::DY_BYTES, &full[buf]);
      }
      for (int it0 = (a.dyc ? nDY : 0) + stid; it0 < nIt; it0 += WGK_NITEM * WGK_NSTAGE_THREADS) {
        float4 v[WGK_NITEM][8];
        int dst[WGK_NITEM];               // byte offset of the item's first pixel inside its first plane, -1 = no item
        bool isin[WGK_NITEM];
#pragma unroll
        for (int r = 0; r < WGK_NITEM; ++r) {
          const int it = it0 + WGK_NSTAGE_THREADS * r;
          dst[r] = -1; isin[r] = false;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[r][e] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int l5 = it & 31, g = ((l5 >> 3) << 1) | ((l5 >> 2) & 1);
          if (it < nDY) {
            const int grp = it >> 5, row = grp % K::TH, q = (grp / K::TH) * 4 + (l5 & 3);
            if (q * 8 < a.Cout) dst[r] = q * K::DY_PLANE + (row * K::TW + 4 * g) * 16;   // planes past Cout stay zero
            const float* src = dyb + (size_t)(q * 8) * HW + (size_t)(y0 + row) * W + x0 + 4 * g;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (q * 8 + e < a.Cout) v[r][e] = *reinterpret_cast<const float4*>(src + (size_t)e * HW);
          } else if (it < nIt) {
            const int row = (it - nDY) >> 5, q = l5 & 3;
            const int gy = y0 + row + ky - K::PAD;
            isin[r] = true;
            dst[r] = q * K::IN_PLANE + (row * K::WP + K::PAD + 4 * g) * 16;
            if (gy >= 0 && gy < H) {
              const float* src = inb + (size_t)(q * 8) * HW + (size_t)gy * W + x0 + 4 * g;
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (q * 8 + e < a.Cin) v[r][e] = *reinterpret_cast<const float4*>(src + (size_t)e * HW);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < WGK_NITEM; ++r) {
          if (dst[r] < 0) continue;
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              split_bf16x2(reinterpret_cast<const float*>(&v[r][2 * e])[px], reinterpret_cast<const float*>(&v[r][2 * e + 1])[px], hi[e], lo[e]);
            const uint4 uh = make_uint4(hi[0], hi[1], hi[2], hi[3]), ul = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            if (!isin[r]) {
              *reinterpret_cast<uint4*>(sDY + dst[r] + px * 16) = uh;
              *reinterpret_cast<uint4*>(sDY + 16 * K::DY_PLANE + dst[r] + px * 16) = ul;
            } else {     // plane sets: hh, hh shifted (pixel p lands at p - 1; PAD >= 1 so p >= 1), hl, hl shifted
              *reinterpret_cast<uint4*>(sIN + dst[r] + px * 16) = uh;
              *reinterpret_cast<uint4*>(sIN + 4 * K::IN_PLANE + dst[r] + px * 16 - 16) = uh;
              *reinterpret_cast<uint4*>(sIN + 8 * K::IN_PLANE + dst[r] + px * 16) = ul;
              *reinterpret_cast<uint4*>(sIN + 12 * K::IN_PLANE + dst[r] + px * 16 - 16) = ul;
            }
          }
        }
      }
      // halo columns of the h tile (always outside the image when W == 32)
      for (int it = stid; it < 4 * K::TH * 2 * K::PAD; it += WGK_NSTAGE_THREADS) {
        const int q = it / (K::TH * 2 * K::PAD), rem = it - q * (K::TH * 2 * K::PAD), row = rem / (2 * K::PAD), hc = rem - row * (2 * K::PAD);
        const int c = hc < K::PAD ? hc : (K::TW + hc);
        const int gy = y0 + row + ky - K::PAD, gx = x0 + c - K::PAD;
        const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v0 = 0.f, v1 = 0.f;
          if (ok && q * 8 + 2 * e < a.Cin) v0 = inb[(size_t)(q * 8 + 2 * e) * HW + (size_t)gy * W + gx];
          if (ok && q * 8 + 2 * e + 1 < a.Cin) v1 = inb[(size_t)(q * 8 + 2 * e + 1) * HW + (size_t)gy * W + gx];
          split_bf16x2(v0, v1, hi[e], lo[e]);
        }
        const int p = row * K::WP + c;
        const uint4 uh = make_uint4(hi[0], hi[1], hi[2], hi[3]), ul = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(sIN + q * K::IN_PLANE + p * 16) = uh;
        *reinterpret_cast<uint4*>(sIN + 8 * K::IN_PLANE + q * K::IN_PLANE + p * 16) = ul;
        if (p >= 1) {
          *reinterpret_cast<uint4*>(sIN + (4 + q) * K::IN_PLANE + (p - 1) * 16) = uh;
          *reinterpret_cast<uint4*>(sIN + (12 + q) * K::IN_PLANE + (p - 1) * 16) = ul;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&full[buf]);
    }
  }
  __syncwarp();

  // ---- epilogue: the CTA's KS accumulators -> its slot [r0][ky*KS + kx][128][32] ----------------------------------
  if (myTiles > 0) {
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quad = warp & 3, par = warp >> 2;
    float* slot = a.slots + ((size_t)r0 * K::KSZ * K::KSZ + (size_t)ky * K::KSZ) * 4096 + (size_t)(quad * 32 + lane) * 32;
#pragma unroll 1
    for (int kx = par; kx < K::KSZ; kx += 2) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(kx * 32), v);
      float4* dst = reinterpret_cast<float4*>(slot + (size_t)kx * 4096);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 t = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        if (a.accumulate) { const float4 p = dst[i]; t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w; }
        dst[i] = t;
      }
    }
  } else if (!a.accumulate) {   // a CTA without tiles still owns a slot: define it
    float* slot = a.slots + ((size_t)r0 * K::KSZ * K::KSZ + (size_t)ky * K::KSZ) * 4096;
    for (int i = tid; i < K::KSZ * 4096; i += 256) slot[i] = 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(K::TCOLS));
}

// slots [R][taps][128][32] -> the `nbank` stacked [ocb][icb][taps] tensors inside dparams (double-precision fold, fixed order)
// icb = input channels covered by these slots (<= 32), placed at ic0 inside tensors with ic_total input channels
__global__ void fold_wgradk_kernel(const float* __restrict__ slots, int R, int taps, int nbank, int ocb, int icb, float* dst0,
                                   long long stride_between, int ic0, int ic_total) {
  const int total = nbank * ocb * icb * taps;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int tap = i % taps, ic = (i / taps) % icb, n = i / (taps * icb);
  double s = 0;
  for (int r = 0; r < R; ++r) s += slots[(((size_t)r * taps + tap) * 128 + n) * 32 + ic];
  dst0[(size_t)(n / ocb) * stride_between + ((size_t)(n % ocb) * ic_total + ic0 + ic) * taps + tap] = (float)s;
}

// fp32 [B,Cout,H,W] -> per tile [32 planes][128 px][8 ch] bf16 (hi planes 0-15, lo planes 16-31, WgK::DY_PLANE pitch)
__global__ void wgradk_prep_dy_kernel(const float* __restrict__ dy, int B, int H, int W, int Cout, uint8_t* dst, const float* skip) {
  if (skip != nullptr && skip[1] != 0.f) return;
  using K = WgK<7>;     // the delta-tile layout does not depend on the kernel size
  const int tilesX = W / K::TW, tilesY = H / K::TH, nTiles = B * tilesX * tilesY;
  const size_t HW = (size_t)H * W, total = (size_t)nTiles * 4 * K::TH * 32;     // items: (tile, plane group, row, lane)
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int l5 = (int)(e & 31), g = ((l5 >> 3) << 1) | ((l5 >> 2) & 1);
    const int grp = (int)((e >> 5) % (4 * K::TH)), tile = (int)(e / (4 * K::TH * 32));
    const int row = grp % K::TH, q = (grp / K::TH) * 4 + (l5 & 3);
    const int b = tile / (tilesX * tilesY), trem = tile - b * (tilesX * tilesY), y0 = (trem / tilesX) * K::TH, x0 = (trem % tilesX) * K::TW;
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q * 8 + c < Cout) v[c] = *reinterpret_cast<const float4*>(dy + ((size_t)b * Cout + q * 8 + c) * HW + (size_t)(y0 + row) * W + x0 + 4 * g);
    }
    uint8_t* o = dst + (size_t)tile * K::DY_BYTES + (size_t)q * K::DY_PLANE + (size_t)(row * K::TW + 4 * g) * 16;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        split_bf16x2(reinterpret_cast<const float*>(&v[2 * c])[px], reinterpret_cast<const float*>(&v[2 * c + 1])[px], hi[c], lo[c]);
      *reinterpret_cast<uint4*>(o + px * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(o + 16 * K::DY_PLANE + px * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}
size_t wgradk_tc_dyc_bytes(int B, int H, int W) { return (size_t)B * (H / 4) * (W / 32) * WgK<7>::DY_BYTES; }
int launch_wgradk_prep_dy(const float* dy, int B, int H, int W, int Cout, void* dyc, const float* skip, cudaStream_t st) {
  static_assert(WgK<7>::DY_BYTES == WgK<15>::DY_BYTES && WgK<7>::DY_PLANE == WgK<15>::DY_PLANE, "one delta-tile layout");
  if (H % 4 || W % 32 || Cout > 128) { set_error("wgradk_prep_dy: unsupported H=%d W=%d Cout=%d", H, W, Cout); return PTK_EINVAL; }
  wgradk_prep_dy_kernel<<<148 * 8, 256, 0, st>>>(dy, B, H, W, Cout, (uint8_t*)dyc, skip);
  PTK_LAUNCH_OK("wgradk_prep_dy_kernel");
  return PTK_OK;
}

int wgradk_tc_groups(int ks) { return 148 / ks; }
size_t wgradk_tc_slot_floats(int ks) { return (size_t)wgradk_tc_groups(ks) * ks * ks * 4096; }

template <int KSZ>
int launch_wgradk_tc(const WgradKArgs& a, cudaStream_t st) {
  using K = WgK<KSZ>;
  if (a.H % K::TH || a.W % K::TW || a.Cin > 32 || a.Cout > 128) {
    set_error("wgradk_tc: unsupported problem H=%d W=%d Cin=%d Cout=%d", a.H, a.W, a.Cin, a.Cout);
    return PTK_EINVAL;
  }
  PTK_CUDA_OK(ptk_func_smem(wgradk_tc_kernel<K>, K::SMEM));
  wgradk_tc_kernel<K><<<dim3(wgradk_tc_groups(KSZ), KSZ), 256, K::SMEM, st>>>(a);
  PTK_LAUNCH_OK("wgradk_tc_kernel");
  return PTK_OK;
}
template int launch_wgradk_tc<7>(const WgradKArgs&, cudaStream_t);
template int launch_wgradk_tc<15>(const WgradKArgs&, cudaStream_t);

int launch_fold_wgradk(const float* slots, int ks, int nbank, int ocb, int icb, float* dst0, long long stride_between, cudaStream_t st,
                       int ic0, int ic_total) {
  const int total = nbank * ocb * icb * ks * ks;
  fold_wgradk_kernel<<<(total + 255) / 256, 256, 0, st>>>(slots, wgradk_tc_groups(ks), ks * ks, nbank, ocb, icb, dst0, stride_between,
                                                          ic0, ic_total ? ic_total : icb);
  PTK_LAUNCH_OK("fold_wgradk_kernel");
  return PTK_OK;
}

}  // namespace ptk
