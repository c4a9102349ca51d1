// The step either side of the recurrent path: loss + metrics on device, fused Adam.
//
//   ptk_bce_scores  : criterion = BCEWithLogitsLoss (mainclean.py:156,189) and acc_scores / metric_scores
//                     (utils/misc_functions.py:12-45) in ONE launch, no host round trips (the reference does
//                     six .item() syncs per iteration, mainclean.py:190-201).
//   ptk_adam_step   : torch.optim.Adam(lr) defaults (mainclean.py:157: betas (0.9, 0.999), eps 1e-8, no weight
//                     decay, no amsgrad) over ONE flat fp32 vector + optional zero_grad (mainclean.py:205-206).
#include "ptk_common.cuh"
#include "ptk_profile.h"

namespace ptk {

// out[0] = mean BCE-with-logits loss, out[1] = bal_acc*100, out[2] = precision, out[3] = recall, out[4] = f1;
// dlogits[i] = (sigmoid(z_i) - y_i) / B  (gradient of the mean loss).  One CTA; B is a batch size.
__global__ void __launch_bounds__(256) bce_scores_kernel(int B, const float* __restrict__ z, const float* __restrict__ y,
                                                         float* __restrict__ dlogits, float* __restrict__ out) {
  __shared__ double sL[256];
  __shared__ int sC[256], sTp[256], sPp[256];
  double l = 0;
  int correct = 0, tp = 0, pp = 0;
  for (int i = threadIdx.x; i < B; i += 256) {
    const float zi = z[i], yi = y[i];
    // max(z,0) - z*y + log1p(exp(-|z|)): the numerically stable form PyTorch uses
    l += (double)(fmaxf(zi, 0.f) - zi * yi + log1pf(expf(-fabsf(zi))));
    if (dlogits) dlogits[i] = (1.0f / (1.0f + expf(-zi)) - yi) / (float)B;
    const int pred = zi > 0.5f;                 // misc_functions.py:40: the LOGIT is compared with 0.5
    const int tgt = (int)(unsigned char)yi;     // target.byte()
    correct += (pred == tgt);
    tp += (pred == tgt) && (tgt == 1);
    pp += pred;
  }
  sL[threadIdx.x] = l; sC[threadIdx.x] = correct; sTp[threadIdx.x] = tp; sPp[threadIdx.x] = pp;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sL[threadIdx.x] += sL[threadIdx.x + s]; sC[threadIdx.x] += sC[threadIdx.x + s];
      sTp[threadIdx.x] += sTp[threadIdx.x + s]; sPp[threadIdx.x] += sPp[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float P = (float)B, tpf = (float)sTp[0];
    const float tpfp = sPp[0] == 0 ? 1e-6f : (float)sPp[0];   // misc_functions.py:21-22
    out[0] = (float)(sL[0] / B);
    out[1] = 100.0f * (float)sC[0] / P;                       // :27, :45
    out[2] = tpf / tpfp;                                      // :24
    out[3] = tpf / P;                                         // :23 (recall over the batch size, as written)
    out[4] = 2.0f * tpf / (P + tpfp);                         // :28
  }
}

// Adam, PyTorch single-tensor semantics (torch/optim/adam.py, _single_tensor_adam, amsgrad False, maximize False):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float lr, float b1,
                                                   float b2, float eps, float step_size, float inv_sqrt_bc2, int zero_grad) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] -= step_size * (mi / denom);
    if (zero_grad) g[i] = 0.f;
  }
}

}  // namespace ptk

using namespace ptk;

extern "C" {

PTK_API int ptk_bce_scores(int32_t B, const float* logits, const float* target, float* dlogits, float* out5, void* stream) {
  if (B < 1 || !logits || !target || !out5) { set_error("ptk_bce_scores: bad arguments (B=%d)", B); return PTK_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope ps(PTK_PROF_MISC, st);
  bce_scores_kernel<<<1, 256, 0, st>>>(B, logits, target, dlogits, out5);
  PTK_LAUNCH_OK("bce_scores_kernel");
  return PTK_OK;
}

PTK_API int ptk_adam_step(int64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, float lr, float beta1,
                          float beta2, float eps, int64_t step, int32_t zero_grad, void* stream) {
  if (n < 1 || step < 1 || !params || !grads || !exp_avg || !exp_avg_sq) {
    set_error("ptk_adam_step: bad arguments (n=%lld step=%lld)", (long long)n, (long long)step);
    return PTK_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  ProfileScope ps(PTK_PROF_MISC, st);
  adam_kernel<<<grid, 256, 0, st>>>((long long)n, params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step_size,
                                    inv_sqrt_bc2, zero_grad);
  PTK_LAUNCH_OK("adam_kernel");
  return PTK_OK;
}

}  // extern "C"
