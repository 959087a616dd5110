// Optional per-kernel-class timing with CUDA events on the launching stream (bench.py uses it to
// report roofline.achieved from live launches inside the timed region).  Off by default: the
// hooks then cost one predictable branch.
#pragma once
#include <cuda_runtime.h>

namespace gib {

// 0 / 1: launches of the tcgen05 kernels (forward + dX, weight gradients); 3 / 4: the same contracts on the fp32 SIMT
// kernels (narrow / tiny problems, and everything when tensor cores are off)
enum ProfClass : int { PROF_GEMM_NT = 0, PROF_GEMM_DW = 1, PROF_SCATTER = 2, PROF_GEMM_NT_SIMT = 3, PROF_GEMM_DW_SIMT = 4,
                       PROF_NCLASS = 5 };

extern bool g_prof_on;
void prof_begin(int cls, double work, cudaStream_t st);
void prof_end(cudaStream_t st);

struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int cls, double work, cudaStream_t s) : st(s), on(g_prof_on) {
    if (on) prof_begin(cls, work, st);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

}  // namespace gib
