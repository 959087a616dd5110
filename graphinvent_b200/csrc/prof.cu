#include <vector>

#include "../../include/gib200.h"
#include "prof.cuh"
#include "common.cuh"

namespace gib {

bool g_prof_on = false;
long long g_launch_count = 0;

struct Rec { cudaEvent_t a, b; int cls; double work; };
static std::vector<Rec> g_pool;   // event pairs, reused across collections
static size_t g_used = 0;

void prof_begin(int cls, double work, cudaStream_t st) {
  if (g_used == g_pool.size()) {
    Rec r{};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    g_pool.push_back(r);
  }
  Rec& r = g_pool[g_used];
  r.cls = cls;
  r.work = work;
  cudaEventRecord(r.a, st);
}
void prof_end(cudaStream_t st) {
  cudaEventRecord(g_pool[g_used].b, st);
  ++g_used;
}

}  // namespace gib

using namespace gib;

extern "C" {

long long gib_launch_count(void) { return g_launch_count; }

void gib_profile_enable(int on) {
  g_prof_on = on != 0;
  if (on) g_used = 0;
}

// Per-launch records since gib_profile_enable(1), in launch order (call before gib_profile_collect, which clears them):
// fills at most cap entries and returns the number of records (or a negative CUDA error).
int gib_profile_records(double* ms, double* work, int* cls, int cap) {
  for (size_t i = 0; i < g_used && (int)i < cap; ++i) {
    float t = 0.f;
    cudaError_t e = cudaEventElapsedTime(&t, g_pool[i].a, g_pool[i].b);
    if (e != cudaSuccess) return -(int)e;
    ms[i] = t; work[i] = g_pool[i].work; cls[i] = g_pool[i].cls;
  }
  return (int)g_used;
}

// Call after synchronising the stream.  For each class c: ms[c] = summed event time,
// work[c] = summed algorithmic FLOPs (GEMM classes) or bytes (scatter), count[c] = launches.
int gib_profile_collect(double* ms, double* work, long long* count) {
  for (int c = 0; c < PROF_NCLASS; ++c) { ms[c] = 0; work[c] = 0; count[c] = 0; }
  for (size_t i = 0; i < g_used; ++i) {
    float t = 0.f;
    cudaError_t e = cudaEventElapsedTime(&t, g_pool[i].a, g_pool[i].b);
    if (e != cudaSuccess) return (int)e;
    ms[g_pool[i].cls] += t;
    work[g_pool[i].cls] += g_pool[i].work;
    count[g_pool[i].cls] += 1;
  }
  g_used = 0;
  return 0;
}

}  // extern "C"
