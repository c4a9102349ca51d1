#pragma once
#include <cuda_runtime.h>

// kernel classes reported by ptk_profile_stop
enum {
  PTK_PROF_CONV_FWD = 0, PTK_PROF_CONV_DGRAD = 1, PTK_PROF_WGRAD = 2, PTK_PROF_POINT_FWD = 3,
  PTK_PROF_POINT_BWD = 4, PTK_PROF_BN_APPLY = 5, PTK_PROF_MISC = 6, PTK_PROF_GATE_WGRAD = 7, PTK_PROF_NKINDS = 8
};

namespace ptk {
// RAII: counts `n_launches` kernel launches and, when profiling is on, brackets them with events.
class ProfileScope {
 public:
  ProfileScope(int kind, cudaStream_t st, int n_launches = 1);
  ~ProfileScope();
 private:
  int idx_;
  cudaStream_t st_;
};
}  // namespace ptk
