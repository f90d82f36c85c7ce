// Optional per-kernel timing with CUDA events on the caller's stream (used by bench.py for the roofline numbers).
// Disabled by default: no events are created or recorded unless psn_profile_enable(1) was called.
#include <mutex>
#include <vector>

#include "prof.cuh"

namespace psn {

struct ProfRec { int tag; long long rows; cudaEvent_t a, b; };
static bool g_on = false;
static std::mutex g_mu;
static std::vector<ProfRec> g_recs;

ProfScope::ProfScope(int tag, long long rows, cudaStream_t st) : idx_(-1), st_(st) {
  if (!g_on) return;
  ProfRec r;
  r.tag = tag;
  r.rows = rows;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(r);
  idx_ = (int)g_recs.size() - 1;
}
ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEventRecord(g_recs[idx_].b, st_);
}

}  // namespace psn

using namespace psn;

extern "C" int psn_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_recs.clear();
  g_on = on != 0;
  return PSN_OK;
}

extern "C" int psn_profile_collect(int n_tags, int64_t* launches, double* ms, double* rows) {
  PSN_REQUIRE(launches && ms && rows && n_tags > 0, PSN_ERR_ARG, "psn_profile_collect: null argument");
  std::lock_guard<std::mutex> lk(g_mu);
  for (int t = 0; t < n_tags; ++t) { launches[t] = 0; ms[t] = 0; rows[t] = 0; }
  for (auto& r : g_recs) {
    PSN_CUDA_CHECK(cudaEventSynchronize(r.b));
    float e = 0.f;
    PSN_CUDA_CHECK(cudaEventElapsedTime(&e, r.a, r.b));
    if (r.tag >= 0 && r.tag < n_tags) {
      launches[r.tag] += 1;
      ms[r.tag] += e;
      if (r.rows > 0) rows[r.tag] += (double)r.rows;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  return PSN_OK;
}
