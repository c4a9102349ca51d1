// Optional in-library kernel timing + launch counting (used by bench.py, off by default).
// When enabled, every kernel-class launch is bracketed by cudaEvents recorded on the launching
// stream; ptk_profile_stop() synchronises those events and returns per-class totals.
#include <atomic>
#include <mutex>
#include <vector>

#include "ptk_common.cuh"
#include "ptk_profile.h"

namespace ptk {

static std::mutex g_mu;
static std::atomic<bool> g_on{false};
static std::atomic<long long> g_launches{0};   // the non-profiling path takes no lock
struct Span { cudaEvent_t a, b; int kind; };
static std::vector<Span> g_spans;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

ProfileScope::ProfileScope(int kind, cudaStream_t st, int n_launches) : idx_(-1), st_(st) {
  g_launches.fetch_add(n_launches, std::memory_order_relaxed);
  if (!g_on.load(std::memory_order_acquire)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_on.load(std::memory_order_relaxed)) return;
  Span s{get_event(), get_event(), kind};
  cudaEventRecord(s.a, st);
  g_spans.push_back(s);
  idx_ = (int)g_spans.size() - 1;
}
ProfileScope::~ProfileScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (idx_ < (int)g_spans.size()) cudaEventRecord(g_spans[idx_].b, st_);
}

}  // namespace ptk

using namespace ptk;

extern "C" {

PTK_API void ptk_profile_start(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& s : g_spans) { g_pool.push_back(s.a); g_pool.push_back(s.b); }
  g_spans.clear();
  g_on = true;
}

// ms[k], count[k] for k < PTK_PROF_NKINDS; returns number of spans folded
PTK_API int ptk_profile_stop(double* ms, long long* count) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = false;
  for (int k = 0; k < PTK_PROF_NKINDS; ++k) { ms[k] = 0; count[k] = 0; }
  int n = 0;
  for (auto& s : g_spans) {
    float t = 0.f;
    if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
      ms[s.kind] += t;
      count[s.kind] += 1;
      ++n;
    }
    g_pool.push_back(s.a);
    g_pool.push_back(s.b);
  }
  g_spans.clear();
  return n;
}

PTK_API long long ptk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
