// extern "C" boundary of libgib200.so (see include/gib200.h for the contract).
#include <stdarg.h>
#include <string.h>

#include "../../include/gib200.h"
#include "gemm.cuh"
#include "model.cuh"
#include "ops.cuh"

namespace gib {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- KL-divergence loss + gradient (Workflow.py:833-860), one CTA per molecule ---------
template <int NT>
__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  float r = sm[0];
  for (int i = 1; i < NT / 32; ++i) r = is_max ? fmaxf(r, sm[i]) : r + sm[i];
  return r;
}

__global__ void __launch_bounds__(256) kl_loss_kernel(const float* __restrict__ out, const float* __restrict__ target,
                                                      int apd, float grad_scale, float* __restrict__ loss_rows,
                                                      float* __restrict__ dout) {
  __shared__ float sm[8];
  const int b = blockIdx.x;
  const float* o = out + (size_t)b * apd;
  const float* t = target + (size_t)b * apd;
  float mx = -INFINITY, ts = 0.f;
  for (int k = threadIdx.x; k < apd; k += 256) { mx = fmaxf(mx, o[k]); ts += t[k]; }
  mx = block_reduce<256>(mx, sm, true);
  ts = block_reduce<256>(ts, sm, false);
  float se = 0.f;
  for (int k = threadIdx.x; k < apd; k += 256) se += expf(o[k] - mx);
  se = block_reduce<256>(se, sm, false);
  const float lse = mx + logf(se);
  float acc = 0.f;
  for (int k = threadIdx.x; k < apd; k += 256) {
    const float th = t[k] / ts;                 // Workflow.py:854 (NaN for an all-zero target row, as in the reference)
    const float logp = o[k] - lse;
    acc += (th > 0.f ? th * logf(th) : (th == 0.f ? 0.f : th)) - th * logp;   // xlogy(t,t) - t*logp
    if (dout) dout[(size_t)b * apd + k] = (expf(logp) - th) * grad_scale;
  }
  acc = block_reduce<256>(acc, sm, false);
  if (threadIdx.x == 0 && loss_rows) loss_rows[b] = acc;
}

// ---- validation NLL per sub-graph (Analyzer.get_validation_likelihood, Analyzer.py:744-758):
//      nll[b] = -log( sum_k softmax(out[b])_k * target[b,k] / sum(target[b]) );  an all-zero target row gives NaN,
//      which the reference filters out afterwards (Analyzer.py:756) -----------------------------------------------
__global__ void __launch_bounds__(256) validation_nll_kernel(const float* __restrict__ out,
                                                             const float* __restrict__ target, int apd,
                                                             float* __restrict__ nll) {
  __shared__ float sm[8];
  const int b = blockIdx.x;
  const float* o = out + (size_t)b * apd;
  const float* t = target + (size_t)b * apd;
  float mx = -INFINITY, ts = 0.f;
  for (int k = threadIdx.x; k < apd; k += 256) { mx = fmaxf(mx, o[k]); ts += t[k]; }
  mx = block_reduce<256>(mx, sm, true);
  ts = block_reduce<256>(ts, sm, false);
  float se = 0.f, dot = 0.f;
  for (int k = threadIdx.x; k < apd; k += 256) {
    const float e = expf(o[k] - mx);
    se += e;
    dot += e * t[k];
  }
  se = block_reduce<256>(se, sm, false);
  dot = block_reduce<256>(dot, sm, false);
  if (threadIdx.x == 0) nll[b] = -logf((dot / se) / ts);     // ts == 0 -> 0/0 = NaN like target/sum(target)
}

// ---- categorical sampling of one action per molecule (GraphGenerator.py:121, 535-542) ----
__global__ void __launch_bounds__(256) sample_actions_kernel(const float* __restrict__ out, int apd,
                                                             const float* __restrict__ uniforms,
                                                             int* __restrict__ action, float* __restrict__ lik) {
  __shared__ float sm[8];
  __shared__ float pre[257];
  const int b = blockIdx.x;
  const float* o = out + (size_t)b * apd;
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < apd; k += 256) mx = fmaxf(mx, o[k]);
  mx = block_reduce<256>(mx, sm, true);
  const int L = ceil_div(apd, 256);
  const int lo = min(apd, (int)threadIdx.x * L), hi = min(apd, lo + L);
  float s = 0.f;
  for (int k = lo; k < hi; ++k) s += expf(o[k] - mx);
  pre[threadIdx.x + 1] = s;
  if (threadIdx.x == 0) pre[0] = 0.f;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 1; i <= 256; ++i) pre[i] += pre[i - 1];   // fixed-order prefix: deterministic
  __syncthreads();
  const float total = pre[256];
  const float target = uniforms[b] * total;
  // the owner thread is the first chunk whose inclusive prefix exceeds the target
  const bool owner = (pre[threadIdx.x] <= target && target < pre[threadIdx.x + 1]) ||
                     (threadIdx.x == 255 && target >= pre[256]);
  if (owner && hi > lo) {
    float run = pre[threadIdx.x];
    int pick = hi - 1;
    for (int k = lo; k < hi; ++k) {
      run += expf(o[k] - mx);
      if (target < run) { pick = k; break; }
    }
    action[b] = pick;
    lik[b] = expf(o[pick] - mx) / total;
  } else if (owner) {  // empty tail chunk: fall back to the last element
    action[b] = apd - 1;
    lik[b] = expf(o[apd - 1] - mx) / total;
  }
  // a row with NaN / Inf logits has no owner (every comparison with NaN is false): defined outputs instead of the
  // caller's uninitialised memory -- "terminate" with a NaN likelihood, which the caller can detect
  if (threadIdx.x == 0 && !(total == total && total > 0.f && total < INFINITY && target == target)) {
    action[b] = apd - 1;
    lik[b] = NAN;
  }
}

}  // namespace gib

using namespace gib;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

const char* gib_last_error(void) { return g_err; }
int gib_version(void) { return 201; }   // 200: capacity mode, int8 inputs, second-generation tcgen05 GEMM, grouped dW; 201: 5 profile classes
void gib_set_tensor_cores(int on) { g_use_tc = on != 0; }
int gib_get_tensor_cores(void) { return g_use_tc ? 1 : 0; }
void gib_tc_debug(int mode) { g_tc_debug = mode; }
int gib_device_sm_count(void) { return device_sm_count(); }
void gib_scatter_variant(int v) { g_scatter_variant = v; }
void gib_tc_trace(long long* device_buf, int tiles) { tc3_set_trace(device_buf, tiles); }

static int groups_of(const gib_dims* d) { return d->model == GIB_EMN ? 1 : d->Ef; }
static bool is_cap(const int* hdr) { return hdr[HDR_CAPACITY] != 0; }

size_t gib_graph_count_ws_bytes(const gib_dims* d) {
  return graph_count_ws_ints(d->B, groups_of(d)) * sizeof(int);
}
int gib_graph_count(const gib_dims* d, const void* edges, void* count_ws, gib_stream stream) {
  return graph_count(edges, d->in_dtype, d->B, d->N, d->Ef, d->model != GIB_EMN, reinterpret_cast<int*>(count_ws),
                     ST(stream));
}
int gib_graph_header_capacity(const gib_dims* d, int entry_capacity, const void* count_ws, int* hdr) {
  if (entry_capacity < 1 || !count_ws) { set_error("gib_graph_header_capacity: bad capacity / workspace"); return -1; }
  const int G = groups_of(d);
  memset(hdr, 0, HDR_INTS * sizeof(int));
  hdr[HDR_E] = entry_capacity;
  // every type group is padded to a multiple of 128 rows: at most 127 pad rows per group
  hdr[HDR_P] = ceil_div(entry_capacity, kTileRows) * kTileRows + G * kTileRows;
  hdr[HDR_CAPACITY] = 1;
  const unsigned long long a = reinterpret_cast<unsigned long long>(count_ws);
  hdr[HDR_DEV_LO] = (int)(unsigned)(a & 0xffffffffull);
  hdr[HDR_DEV_HI] = (int)(unsigned)(a >> 32);
  return 0;
}
size_t gib_graph_bytes(const gib_dims* d, const int* hdr) {
  return graph_buf_ints((long long)d->B * d->N, hdr[HDR_E], hdr[HDR_P]) * sizeof(int);
}
int gib_graph_fill(const gib_dims* d, const void* edges, void* count_ws, const int* hdr, void* graph_buf,
                   gib_stream stream) {
  GraphArrays ga = graph_arrays(graph_buf, (long long)d->B * d->N, hdr[HDR_E], hdr[HDR_P]);
  return graph_fill(edges, d->in_dtype, d->B, d->N, d->Ef, d->model != GIB_EMN, reinterpret_cast<int*>(count_ws), ga,
                    is_cap(hdr) ? hdr[HDR_E] : 0, is_cap(hdr) ? hdr[HDR_P] : 0, ST(stream));
}
void* gib_graph_array(const gib_dims* d, const int* hdr, void* graph_buf, int which) {
  GraphArrays ga = graph_arrays(graph_buf, (long long)d->B * d->N, hdr[HDR_E], hdr[HDR_P]);
  switch (which) {
    case 0: return ga.ent_src;
    case 1: return ga.ent_dst;
    case 2: return ga.ent_w;
    case 3: return ga.dst_ptr;
    case 4: return ga.dst_ent;
    case 5: return ga.src_ptr;
    case 6: return ga.src_ent;
  }
  return nullptr;
}

int gib_model_num_params(const gib_dims* d) {
  Plan pl;
  if (build_plan(*d, pl)) return -1;
  return (int)pl.param_numel.size();
}
long long gib_model_param_numel(const gib_dims* d, int index) {
  Plan pl;
  if (build_plan(*d, pl) || index < 0 || index >= (int)pl.param_numel.size()) return -1;
  return pl.param_numel[index];
}
size_t gib_model_packed_bytes(const gib_dims* d) {
  Plan pl;
  if (build_plan(*d, pl)) return 0;
  return pl.packed_floats * sizeof(float);
}
int gib_model_pack(const gib_dims* d, const float* const* params, void* packed, gib_stream stream) {
  Plan pl;
  GIB_TRY(build_plan(*d, pl));
  return pack_params(pl, params, reinterpret_cast<float*>(packed), ST(stream));
}

size_t gib_model_workspace_bytes(const gib_dims* d, const int* hdr) {
  Run r;
  if (make_run(*d, hdr, r)) return 0;
  return r.L.total * sizeof(float) + 256;
}
int gib_model_forward(const gib_dims* d, const int* hdr, const void* nodes, const void* edges, const void* graph_buf,
                      const void* packed, void* workspace, float* out, gib_stream stream) {
  Run r;
  GIB_TRY(make_run(*d, hdr, r));
  r.nodes = nodes; r.edges = edges;
  r.ga = graph_arrays(const_cast<void*>(graph_buf), r.S, r.E, r.P);
  r.packed = reinterpret_cast<const float*>(packed);
  r.ws = reinterpret_cast<float*>(workspace);
  r.st = ST(stream);
  return model_forward(r, out);
}
size_t gib_model_bwd_scratch_bytes(const gib_dims* d, const int* hdr) {
  Run r;
  if (make_run(*d, hdr, r)) return 0;
  BwdBufs bb;
  make_bwd(r, bb);
  return bb.total * sizeof(float) + 256;
}
int gib_model_backward(const gib_dims* d, const int* hdr, const void* nodes, const void* edges, const void* graph_buf,
                       const void* packed, const void* workspace, const float* out, const float* dout,
                       float* const* grads, void* scratch, gib_stream stream) {
  return gib_model_backward_part(d, hdr, nodes, edges, graph_buf, packed, workspace, out, dout, grads, scratch, 0,
                                 stream);
}
int gib_model_backward_part(const gib_dims* d, const int* hdr, const void* nodes, const void* edges,
                            const void* graph_buf, const void* packed, const void* workspace, const float* out,
                            const float* dout, float* const* grads, void* scratch, int part, gib_stream stream) {
  Run r;
  GIB_TRY(make_run(*d, hdr, r));
  r.nodes = nodes; r.edges = edges;
  r.ga = graph_arrays(const_cast<void*>(graph_buf), r.S, r.E, r.P);
  r.packed = reinterpret_cast<const float*>(packed);
  r.ws = reinterpret_cast<float*>(const_cast<void*>(workspace));
  r.scratch = reinterpret_cast<float*>(scratch);
  r.grads = grads;
  r.st = ST(stream);
  BwdBufs bb;
  make_bwd(r, bb);
  return model_backward(r, bb, out, dout, part);
}

int gib_kl_loss_fwd_bwd(const float* out, const float* target, int B, int apd, float grad_scale, float* loss_rows,
                        float* dout, gib_stream stream) {
  if (B <= 0) return 0;
  kl_loss_kernel<<<B, 256, 0, ST(stream)>>>(out, target, apd, grad_scale, loss_rows, dout);
  GIB_LAUNCH_CHECK();
  return 0;
}

// out[0] = scale * sum_b rows[b], fixed order (one CTA): the batch-mean of the per-molecule losses without an ATen
// reduction inside a captured step
__global__ void __launch_bounds__(256) sum_scaled_kernel(const float* __restrict__ rows, int n, float scale,
                                                         float* __restrict__ out) {
  __shared__ float sm[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += rows[i];
  s = block_reduce<256>(s, sm, false);
  if (threadIdx.x == 0) out[0] = s * scale;
}
int gib_sum_scaled(const float* rows, int n, float scale, float* out, gib_stream stream) {
  sum_scaled_kernel<<<1, 256, 0, ST(stream)>>>(rows, n, scale, out);
  GIB_LAUNCH_CHECK();
  return 0;
}
int gib_fill_zero(void* ptr, size_t bytes, gib_stream stream) {
  if (bytes == 0) return 0;
  GIB_CUDA_TRY(cudaMemsetAsync(ptr, 0, bytes, ST(stream)));
  return 0;
}

int gib_validation_nll(const float* out, const float* target, int B, int apd, float* nll, gib_stream stream) {
  if (B <= 0) return 0;
  validation_nll_kernel<<<B, 256, 0, ST(stream)>>>(out, target, apd, nll);
  GIB_LAUNCH_CHECK();
  return 0;
}

int gib_sample_actions(const float* out, int B, int apd, const float* uniforms, int* action, float* likelihood,
                       gib_stream stream) {
  if (B <= 0) return 0;
  sample_actions_kernel<<<B, 256, 0, ST(stream)>>>(out, apd, uniforms, action, likelihood);
  GIB_LAUNCH_CHECK();
  return 0;
}

int gib_linear_fwd(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy, int M,
                   int N, int K, int act, gib_stream stream) {
  GemmNT p;
  p.A = X; p.lda = ldx; p.B = W; p.ldb = ldw; p.C = Y; p.ldc = ldy; p.M = M; p.N = N; p.K = K; p.bias = bias;
  p.act = act; p.mode = EPI_ACT; p.n_store = N; p.n_valid = N;
  return gemm_nt(p, ST(stream));
}
int gib_linear_fwd_tc(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy, int M,
                      int N, int K, int act, gib_stream stream) {
  GemmNT p;
  p.A = X; p.lda = ldx; p.B = W; p.ldb = ldw; p.C = Y; p.ldc = ldy; p.M = M; p.N = N; p.K = K; p.bias = bias;
  p.act = act; p.mode = EPI_ACT; p.n_store = N; p.n_valid = N;
  return gemm_nt_tc(p, ST(stream));
}
int gib_linear_fwd_tc_planes(const float* X, int ldx, const float* W_hi, const float* W_lo, int ldw, const float* bias,
                             float* Y, int ldy, int M, int N, int K, int act, const int* m_dev, const int* base_dev,
                             gib_stream stream) {
  GemmNT p;
  p.A = X; p.lda = ldx; p.B = W_hi; p.B_hi = W_hi; p.B_lo = W_lo; p.ldb = ldw; p.C = Y; p.ldc = ldy;
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.act = act; p.mode = EPI_ACT; p.n_store = N; p.n_valid = N;
  p.m_dev = m_dev; p.base_dev = base_dev;
  if (g_tc_debug & 1) {
    if (m_dev) { set_error("device-side row counts need the second-generation kernel"); return -2; }
    return gemm_nt_tc(p, ST(stream));
  }
  return gemm_nt_tc3_group(&p, 1, ST(stream));
}
__global__ void split_planes_kernel(const float* __restrict__ W, float* __restrict__ hi, float* __restrict__ lo,
                                    long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = W[i];
  unsigned h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - __uint_as_float(h)));
  hi[i] = __uint_as_float(h);
  lo[i] = __uint_as_float(l);
}
int gib_split_planes(const float* W, float* W_hi, float* W_lo, long long n, gib_stream stream) {
  if (n <= 0) return 0;
  split_planes_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, ST(stream)>>>(W, W_hi, W_lo, n);
  GIB_LAUNCH_CHECK();
  return 0;
}
size_t gib_dw_scratch_bytes(int M, int Nn, int Kk) { return gemm_dw_scratch_floats(M, Nn, Kk) * sizeof(float); }
int gib_linear_bwd_dw(const float* G, int ldg, int Nn, const float* X, int ldx, int Kk, int M, float* dW, float* dbias,
                      int R, int C, void* scratch, const int* m_dev, const int* base_dev, gib_stream stream) {
  GemmDW q;
  q.G = G; q.ldg = ldg; q.Nn = Nn; q.X = X; q.ldx = ldx; q.Kk = Kk; q.M = M; q.dW = dW; q.dbias = dbias;
  q.R = R; q.C = C; q.Rb = R; q.Rbp = Nn; q.rs = C; q.cs = 1; q.scratch = reinterpret_cast<float*>(scratch);
  q.half_floats = gemm_dw_half_floats(M, Nn, Kk);
  q.m_dev = m_dev; q.base_dev = base_dev;
  GIB_TRY(dw_begin());
  const int rc = gemm_dw(q, ST(stream));
  const int rj = dw_join(ST(stream));
  return rc ? rc : rj;
}
int gib_scatter_sum(float* out, const float* msg, int ld, const int* ptr, const int* ent, const float* w, long long S,
                    gib_stream stream) {
  if (ld & 3) { set_error("gib_scatter_sum: ld must be a multiple of 4"); return -2; }
  return scatter_sum(out, msg, ld, ptr, ent, w, 0, S, ST(stream));
}
int gib_seg_softmax(float* out, const float* EM, const float* EN, int ld, const int* ptr, const int* ent,
                    const float* w, long long S, gib_stream stream) {
  return seg_softmax_fwd(out, EM, EN, ld, ptr, ent, w, S, ST(stream));
}
int gib_gru_gates(float* hn, const float* gi, const float* gh, const float* h, int Hp, const int* ptr, long long S,
                  gib_stream stream) {
  return gru_fwd(hn, gi, gh, h, Hp, ptr, S, ST(stream));
}
int gib_graph_gather(float* g, float* att, const float* en, const float* em, int ld, const int* ptr, int N, int B,
                     float big, gib_stream stream) {
  return graph_gather_fwd(g, att, en, em, ld, ptr, N, B, big, ST(stream));
}

}  // extern "C"
