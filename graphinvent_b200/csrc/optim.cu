// Flat-bucket Adam step: one launch over the contiguous parameter / gradient /
// moment buffers, replacing the per-tensor update of torch.optim.Adam at Workflow.py:191,221,245
// (stepped at Workflow.py:795-796).  HBM-bound: 4 streams read, 3 written, 28 B per parameter.
#include <math.h>

#include "../../include/gib200.h"
#include "common.cuh"
#include "gemm.cuh"

namespace gib {

struct AdamScalars {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay;
  float step_size;      // lr / (1 - beta1^t)
  float bc2_sqrt;       // sqrt(1 - beta2^t)
  float grad_scale;     // applied to the gradient first (1 / world size when the bucket holds an all-reduce SUM)
};

// same operation order as torch/optim/adam.py::_single_tensor_adam (non-amsgrad, L2 weight decay):
//   g += wd * p;  m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g;
//   p -= step_size * m / (sqrt(v) / sqrt(bc2) + eps)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& s) {
  g *= s.grad_scale;
  g = fmaf(s.weight_decay, p, g);
  m = fmaf(g - m, s.one_minus_beta1, m);
  v = fmaf(s.one_minus_beta2 * g, g, v * s.beta2);
  const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
  p = fmaf(-s.step_size, m / denom, p);
}

__global__ void __launch_bounds__(256, 6) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        long long head, AdamScalars s) {
  // [0, head) scalar prologue up to 16-byte alignment, then float4 body, then scalar tail
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nvec = (n - head) >> 2;
  float4* p4 = reinterpret_cast<float4*>(p + head);
  const float4* g4 = reinterpret_cast<const float4*>(g + head);
  float4* m4 = reinterpret_cast<float4*>(m + head);
  float4* v4 = reinterpret_cast<float4*>(v + head);
  for (long long i = tid; i < nvec; i += nthreads) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldg(g4 + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, s);
    adam_one(pp.y, gg.y, mm.y, vv.y, s);
    adam_one(pp.z, gg.z, mm.z, vv.z, s);
    adam_one(pp.w, gg.w, mm.w, vv.w, s);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  const long long tail0 = head + (nvec << 2);
  const long long nscalar = head + (n - tail0);
  for (long long i = tid; i < nscalar; i += nthreads) {
    const long long j = i < head ? i : tail0 + (i - head);
    float pp = p[j], mm = m[j], vv = v[j];
    adam_one(pp, g[j], mm, vv, s);
    p[j] = pp; m[j] = mm; v[j] = vv;
  }
}

}  // namespace gib

extern "C" int gib_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                             long long step, double lr, double beta1, double beta2, double eps,
                             double weight_decay, double grad_scale, gib_stream stream) {
  using namespace gib;
  if (n < 0 || step < 1) { set_error("gib_adam_step: n >= 0 and step >= 1 required (n=%lld step=%lld)", n, step); return -2; }
  if (n == 0) return 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq) { set_error("gib_adam_step: null buffer"); return -2; }
  const uintptr_t a = reinterpret_cast<uintptr_t>(params);
  if ((a & 3) || ((reinterpret_cast<uintptr_t>(grads) ^ a) & 15) || ((reinterpret_cast<uintptr_t>(exp_avg) ^ a) & 15) ||
      ((reinterpret_cast<uintptr_t>(exp_avg_sq) ^ a) & 15)) {
    set_error("gib_adam_step: the four buffers must share their alignment modulo 16 bytes");
    return -2;
  }
  AdamScalars s;
  // bias corrections in double on the host, as the reference optimizer computes them in Python floats
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  s.beta1 = (float)beta1; s.beta2 = (float)beta2;
  s.one_minus_beta1 = (float)(1.0 - beta1);
  s.one_minus_beta2 = (float)(1.0 - beta2);
  s.eps = (float)eps; s.weight_decay = (float)weight_decay;
  s.step_size = (float)(lr / bc1);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.grad_scale = (float)grad_scale;
  long long head = ((16 - (a & 15)) & 15) >> 2;
  if (head > n) head = n;
  const long long work = (n + 3) / 4;
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)gib::device_sm_count() * 6;   // one wave: 6 resident CTAs of 256 threads per SM (40 registers)
  if (blocks > cap) blocks = cap;
  adam_flat_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg,
                                                                                         exp_avg_sq, n, head, s);
  GIB_LAUNCH_CHECK();
  return 0;
}
