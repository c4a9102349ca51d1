// Shared device/host helpers for libptk_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ptk_b200.h"

namespace ptk {

constexpr int C = 32;        // hidden channels (utils/engine.py:75 dimensions=32)
constexpr int KS = 7;        // horizontal kernel (utils/engine.py:75 fb_kernel_size=7)
constexpr int TAPS = KS * KS;
constexpr int PT = 128;      // pixels per point-wise tile (one thread per pixel)
constexpr float BN_EPS = 1e-3f;  // models/InT.py:102

// ---- error plumbing (thread-local message, see ptk_last_error) ----------------------------
void set_error(const char* fmt, ...);
#define PTK_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ptk::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                  \
                     cudaGetErrorString(_e));                                              \
      return PTK_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)
#define PTK_LAUNCH_OK(name)                                                                \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ptk::set_error("launch of %s failed at %s:%d: %s", name, __FILE__, __LINE__,         \
                     cudaGetErrorString(_e));                                              \
      return PTK_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

// ---- math ----------------------------------------------------------------------------------
// Precise transcendental maps are deliberately NOT inlined: expf/log1pf expand to ~100 instructions and
// the point-wise kernels evaluate several of them per element; inlining made each kernel > 60 KB of SASS
// (instruction-cache misses were the top stall reason, profiles/r1b_pointwise_wgrad_full.txt).
static __device__ __noinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }
static __device__ __noinline__ float softplus_precise(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
static __device__ __noinline__ float tanh_precise(float x) { return tanhf(x); }
// MUFU-based variants (ex2.approx / lg2.approx / rcp.approx): used by the split-bf16 tensor-core mode, whose
// convolutions are already ~1e-5 accurate; same parity gate (tests/test_parity_gpu.py).  Accuracy: sigmoid ~2 ulp.
// softplus_fast is accurate in ABSOLUTE terms only (~6e-8): for x < -10 the sum 1 + e^x rounds in fp32, so the
// RELATIVE error of the tiny result grows to 1e-3 and reaches 100 % below x ~ -16 (the result flushes to 0 where the
// exact value is < 1.2e-7).  Every use adds the result to O(1) quantities, so absolute accuracy is what the gate needs.
static __device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
static __device__ __forceinline__ float softplus_fast(float x) {
  return fmaxf(x, 0.0f) + __logf(1.0f + __expf(-fabsf(x)));   // = log(1 + e^x), no overflow, threshold-free
}

// Activation codes: PTK_ACT_SOFTPLUS (0), PTK_ACT_TANH (1) as in the C ABI; +2 selects the fast variants.
constexpr int ACT_FAST = 2;
template <int ACT>
__device__ __forceinline__ float sigm(float x) {
  if (ACT & ACT_FAST) return sigmoid_fast(x);
  return sigmoid_precise(x);
}
// activation `nl` of InT.__init__ (models/InT.py:184,208): softplus(beta=1, threshold=20) or tanh
template <int ACT>
__device__ __forceinline__ float actf(float x) {
  if ((ACT & 1) == PTK_ACT_SOFTPLUS) return (ACT & ACT_FAST) ? softplus_fast(x) : softplus_precise(x);
  return tanh_precise(x);
}
template <int ACT>
__device__ __forceinline__ float dactf(float x) {
  if ((ACT & 1) == PTK_ACT_SOFTPLUS) return (ACT & ACT_FAST) ? sigmoid_fast(x) : (x > 20.0f ? 1.0f : sigmoid_precise(x));
  const float t = tanh_precise(x);
  return 1.0f - t * t;
}

// acc[o] += sum_i sW[i*32 + o] * v[i]; sW in shared memory ([in][out]); every lane reads the
// same address, so each float4 is one broadcast wavefront.
__device__ __forceinline__ void matvec32(float (&acc)[C], const float* __restrict__ sW,
                                         const float (&v)[C]) {
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float4* w4 = reinterpret_cast<const float4*>(sW + i * C);
    const float vi = v[i];
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const float4 w = w4[q];
      acc[4 * q + 0] = fmaf(w.x, vi, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(w.y, vi, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(w.z, vi, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(w.w, vi, acc[4 * q + 3]);
    }
  }
}

// ---- cp.async (LDGSTS) ---------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// 16-byte copy that zero-fills when !valid (src-size 0)
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async4_zfill(void* smem, const void* gmem, bool valid) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// The time loop is a chain of ~13 short kernels per step whose CTAs fill a whole SM (shared memory), so
// launch latency + prologue (barrier init, TMEM allocation, weight staging) of kernel N+1 would sit
// exposed behind the tail of kernel N.  Loop kernels are launched with the programmatic-stream-
// serialization attribute and follow one rule:
//     prologue touching only parameters / shared memory / TMEM  ->  pdl_wait()  ->  pdl_launch_dependents()
//     ->  everything that reads or writes activations.
// pdl_wait() returns when the preceding kernel has completed and flushed; triggering only after the
// wait means kernel N+2 can never start before kernel N has finished, so pre-wait code may read data
// produced two or more kernels back (prepared weights) but nothing newer.  Without the launch attribute
// both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// L2 prefetch of a 128-byte line: harmless before pdl_wait() (a hint, no data is consumed), used to pull the
// forward checkpoints a backward kernel is about to read out of HBM while the previous kernel drains.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
bool pdl_enabled();   // PTK_NO_PDL=1 disables the attribute (A/B measurements)
// Every kernel of the time loop asks for the maximum shared-memory carve-out, including the ones that use none
// (bn_apply): a kernel-to-kernel change of the L1/shared split forces the SMs to drain and reconfigure.
template <typename F>
inline cudaError_t ptk_func_smem(F* kern, int dyn_bytes) {
  cudaError_t e = cudaSuccess;
  if (dyn_bytes > 0) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- frame addressing: x is [B,3,T,H,W] with arbitrary element strides ---------------------
struct FrameView {
  const float* x;
  long long sb, sc, st, sy, sx;
  __device__ __forceinline__ float at(int b, int c, int t, int y, int xx) const {
    return __ldg(x + b * sb + c * sc + t * st + y * sy + xx * sx);
  }
};

// ---- flat parameter vector -----------------------------------------------------------------
struct ParamOffsets {
  int64_t o[PTK_INT_NPARAM + 1];
};
inline ParamOffsets param_offsets(int Cc, int K) {
  const int64_t sizes[PTK_INT_NPARAM] = {
      Cc * 3, Cc, (int64_t)Cc * Cc * K * K, (int64_t)Cc * Cc * K * K, Cc, Cc, Cc, Cc,
      Cc * Cc, Cc, Cc * Cc, Cc, Cc * Cc, Cc, Cc * Cc, Cc, Cc * Cc, Cc, Cc * Cc, Cc,
      Cc, Cc, Cc, Cc, Cc, 1, 50, 1, 1, 1};
  ParamOffsets r;
  int64_t a = 0;
  for (int i = 0; i < PTK_INT_NPARAM; ++i) {
    r.o[i] = a;
    a += sizes[i];
  }
  r.o[PTK_INT_NPARAM] = a;
  return r;
}

}  // namespace ptk
