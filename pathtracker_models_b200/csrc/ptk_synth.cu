// On-device synthetic PathTracker clip generator (stands in for utils/TFRDataset.py:6-53; the
// reference ships no generator -- its clips are external TFRecords of uint8 [T,32,32,3] + label).
// Data contract kept: x [B,3,T,H,W] float in {0,1} (utils/engine.py:222-226 divides uint8 by 255),
// channel 2 of frame 0 = goal map (models/InT.py:236), y in {0,1}.
//
// Everything is integer arithmetic (splitmix64 counters, Q16 positions, "magic circle" rotation),
// so oracle/synth_oracle.py reproduces it bit-exactly on the CPU.
//   dot 0 = target, dots 1..D = distractors (at least one is simulated; it is a phantom, not drawn,
//   when D == 0).  ch0: all dots, every frame.  ch1: 3x3 start marker (target, t = 0), static.
//   ch2: 3x3 goal marker at the final position of the target (label 1) or of distractor 1 (label 0).
#include "ptk_common.cuh"

namespace ptk {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t rnd32(uint64_t key, uint32_t k) {
  return (uint32_t)(mix64(key + (uint64_t)k * 0x632BE59BD9B4E019ULL) >> 32);
}

// 16 unit vectors, Q16 (round(65536*cos), round(65536*sin)) of k*22.5 degrees
__constant__ int kDirX[16] = {65536, 60547, 46341, 25080, 0, -25080, -46341, -60547,
                              -65536, -60547, -46341, -25080, 0, 25080, 46341, 60547};
__constant__ int kDirY[16] = {0, 25080, 46341, 60547, 65536, 60547, 46341, 25080,
                              0, -25080, -46341, -60547, -65536, -60547, -46341, -25080};

__global__ void synth_pathtracker_kernel(uint64_t seed, uint64_t first_index, int B, int T, int H, int W,
                                         int D, int speed, float* __restrict__ x, float* __restrict__ y) {
  const int nd = (D > 0 ? D : 1) + 1;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= B * nd) return;
  const int b = gid / nd, dot = gid % nd;
  const uint64_t skey = mix64(seed ^ mix64(first_index + (uint64_t)b));
  const int label = (int)(rnd32(skey, 0) & 1u);
  const uint64_t key = mix64(skey + 0x1000003ULL * (uint64_t)(dot + 1));
  const size_t P = (size_t)H * W, plane = (size_t)T * P;
  float* xb = x + (size_t)b * 3 * plane;
  const int lo = 1 << 16, hix = (W - 2) << 16, hiy = (H - 2) << 16;
  int px = lo + (int)(rnd32(key, 1) % (uint32_t)(hix - lo + 1));
  int py = lo + (int)(rnd32(key, 2) % (uint32_t)(hiy - lo + 1));
  const int dir = (int)(rnd32(key, 3) & 15u);
  int vx = kDirX[dir] * speed, vy = kDirY[dir] * speed;
  int curv = (int)(rnd32(key, 4) % 513u) - 256;
  const bool drawn = (dot == 0) || (dot <= D);
  const int sx = (px + 32768) >> 16, sy = (py + 32768) >> 16;
  for (int t = 0; t < T; ++t) {
    if (t > 0) {
      curv += (int)(rnd32(key, 8 + t) % 257u) - 128;
      curv = curv > 1024 ? 1024 : (curv < -1024 ? -1024 : curv);
      vx -= (int)(((long long)vy * curv) >> 12);
      vy += (int)(((long long)vx * curv) >> 12);
      px += vx;
      py += vy;
      if (px < lo) { px = 2 * lo - px; vx = -vx; }
      if (px > hix) { px = 2 * hix - px; vx = -vx; }
      if (py < lo) { py = 2 * lo - py; vy = -vy; }
      if (py > hiy) { py = 2 * hiy - py; vy = -vy; }
      px = px < lo ? lo : (px > hix ? hix : px);
      py = py < lo ? lo : (py > hiy ? hiy : py);
    }
    if (drawn) {
      const int ix = (px + 32768) >> 16, iy = (py + 32768) >> 16;
      xb[(size_t)t * P + (size_t)iy * W + ix] = 1.0f;
    }
  }
  const int fx = (px + 32768) >> 16, fy = (py + 32768) >> 16;
  if (dot == 0) {
    y[b] = (float)label;
    for (int t = 0; t < T; ++t)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) xb[plane + (size_t)t * P + (size_t)(sy + dy) * W + sx + dx] = 1.0f;
  }
  if ((dot == 0 && label == 1) || (dot == 1 && label == 0)) {
    for (int t = 0; t < T; ++t)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) xb[2 * plane + (size_t)t * P + (size_t)(fy + dy) * W + fx + dx] = 1.0f;
  }
}

}  // namespace ptk

extern "C" int ptk_synth_pathtracker(uint64_t seed, uint64_t first_index, int32_t B, int32_t T, int32_t H,
                                     int32_t W, int32_t n_distractors, int32_t speed, float* x, float* y,
                                     void* stream) {
  if (B < 1 || T < 1 || H < 8 || W < 8 || n_distractors < 0 || speed < 1 || speed > 4 || !x || !y) {
    ptk::set_error("ptk_synth_pathtracker: bad argument (B=%d T=%d H=%d W=%d dist=%d speed=%d)", B, T, H, W,
                   n_distractors, speed);
    return PTK_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PTK_CUDA_OK(cudaMemsetAsync(x, 0, (size_t)B * 3 * T * H * W * sizeof(float), st));
  const int nd = (n_distractors > 0 ? n_distractors : 1) + 1;
  const int n = B * nd;
  ptk::synth_pathtracker_kernel<<<(n + 127) / 128, 128, 0, st>>>(seed, first_index, B, T, H, W, n_distractors,
                                                                speed, x, y);
  PTK_LAUNCH_OK("synth_pathtracker_kernel");
  return PTK_OK;
}
