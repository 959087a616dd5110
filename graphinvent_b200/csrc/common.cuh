// Shared helpers for the GraphINVENT-B200 hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace gib {

constexpr int kTileRows = 128;  // row granularity of a bond-type segment (GEMM M tile)

__host__ __device__ inline int pad16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// SELU constants (torch.nn.SELU; reference gnn/modules.py:127)
#define GIB_SELU_SCALE 1.0507009873554804934193349852946f
#define GIB_SELU_ALPHA 1.6732632423543772848170429916717f

enum Act : int { ACT_NONE = 0, ACT_SELU = 1, ACT_TANH = 2 };

__device__ __forceinline__ float selu_f(float x) {
  return x > 0.f ? GIB_SELU_SCALE * x : (GIB_SELU_SCALE * GIB_SELU_ALPHA) * expm1f(x);
}
// derivative of SELU expressed through its OUTPUT y (y > 0 <=> x > 0)
__device__ __forceinline__ float dselu_from_out(float y) {
  return y > 0.f ? GIB_SELU_SCALE : y + GIB_SELU_SCALE * GIB_SELU_ALPHA;
}
__device__ __forceinline__ float act_f(float x, int act) {
  if (act == ACT_SELU) return selu_f(x);
  if (act == ACT_TANH) return tanhf(x);
  return x;
}
__device__ __forceinline__ float dact_from_out(float y, int act) {
  if (act == ACT_SELU) return dselu_from_out(y);
  if (act == ACT_TANH) return 1.f - y * y;
  return 1.f;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// error plumbing: kernels are launched through GIB_LAUNCH_CHECK so that a bad launch
// configuration is reported at the C-ABI boundary as a cudaError_t (>0).
#define GIB_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return (int)_e;                   \
  } while (0)
extern long long g_launch_count;  // kernels launched by this library (bench.py reports it)
#define GIB_LAUNCH_CHECK()                 \
  do {                                     \
    ++::gib::g_launch_count;               \
    GIB_CUDA_TRY(cudaGetLastError());      \
  } while (0)
#define GIB_TRY(expr)                                        \
  do {                                                       \
    int _r = (expr);                                         \
    if (_r != 0) return _r;                                  \
  } while (0)

void set_error(const char* fmt, ...);

}  // namespace gib
