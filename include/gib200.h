/*
 * gib200 -- C-ABI of the B200-native GraphINVENT MPNN hot path (libgib200.so).
 *
 * The reference has no FFI layer: its boundary for this path is the Python
 * `torch.nn.Module` protocol (`gnn.mpnn.{MNN,GGNN,AttentionGGNN,EMN}(constants)`,
 * `forward(nodes, edges) -> logits`).  The drop-in modules in
 * `graphinvent_b200/gnn/` keep that protocol and bind the entry points below through
 * ctypes (INTEGRATION.md shows the stub).  Each entry point names the reference code
 * it replaces (paths relative to /root/reference/graphinvent/).
 *
 * Conventions
 *  - every pointer except `hdr_host`, `params` / `grads` (host arrays of device pointers)
 *    and `dims` is a DEVICE pointer owned by the caller (PyTorch caching allocator:
 *    `tensor.data_ptr()`); the library never allocates or frees device memory -- sizes come
 *    from the *_bytes() queries;
 *  - `stream` is a cudaStream_t (`torch.cuda.current_stream().cuda_stream`); every call is
 *    asynchronous on it and performs no host synchronisation (the exact-size mode needs ONE 64-byte header read
 *    by the caller between gib_graph_count and gib_graph_fill; capacity mode needs none);
 *  - return value: 0 ok, < 0 invalid argument (see gib_last_error()), > 0 a cudaError_t;
 *  - nothing is thrown across the boundary.  Library state: the thread-local error string, a per-device helper
 *    stream for the split reductions (joined back into `stream` before a call returns) and per-device caches of
 *    function attributes / TMA descriptors (mutex-guarded).  One host thread drives one device's model at a time
 *    (the reference's threading model); different devices are independent;
 *  - all reductions run in a fixed order (no float atomics): results are bit-stable.
 *
 * Tensor layouts (reference `BlockDatasetLoader.py:135-143`):
 *    nodes  float32 [B, N, F]        dense, zero padded
 *    edges  float32 [B, N, N, Ef]    dense, zero padded, dst = row i, src = column j
 *    out    float32 [B, N*f_add + N*f_conn + 1]   un-normalised (SELU-activated) APD logits
 */
#ifndef GIB200_H
#define GIB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gib_stream; /* cudaStream_t */

enum { GIB_GGNN = 0, GIB_MNN = 1, GIB_ATTGGNN = 2, GIB_EMN = 3 };

/* Hyper-parameters the reference reads from `constants`. */
typedef struct gib_dims {
  int model;                       /* GIB_* */
  int B, N, F, Ef;                 /* batch, max_n_nodes, n_node_features, n_edge_features */
  int H, M, T;                     /* hidden_node_features, message_size, message_passes
                                      (EMN: H = M = edge_emb_size) */
  int msg_hidden, msg_depth;       /* GGNN: enn_*;  AttGGNN / EMN: msg_* */
  int att_hidden, att_depth;       /* AttGGNN / EMN: att_* */
  int eemb_hidden, eemb_depth;     /* EMN: edge_emb_* */
  int gather_width, gatt_hidden, gatt_depth, gemb_hidden, gemb_depth;
  int mlp1_hidden, mlp1_depth, mlp2_hidden, mlp2_depth;
  int f_add, f_conn;               /* len_f_add_per_node, len_f_conn_per_node */
  float big;                       /* constants.big_positive (1e6) */
  int in_dtype;                    /* element type of `nodes` / `edges`: 0 = float32 (BlockDatasetLoader.py:139-143),
                                      1 = int8, the reference's on-disk type (DataProcesser.py:157-161), read
                                      directly by K0 and the first-layer kernels */
} gib_dims;

/* Graph header: 16 ints written on the device by gib_graph_count() at the start of its workspace.
 * Exact mode: the caller copies them to the host (the one D2H read of a forward) and passes them back as `hdr_host`;
 * buffers are then sized exactly.  Capacity mode: the caller never reads them -- gib_graph_header_capacity() builds a
 * host header that carries static capacities and the ADDRESS of the device header, every kernel whose extent depends
 * on the batch content (the per-bond-type GEMM row ranges, the split counts of the weight-gradient reductions) reads
 * the live counts from device memory, and no launch parameter depends on the batch: a step has no host
 * synchronisation and is capturable in a CUDA graph.  Index meaning: */
enum {
  GIB_HDR_E = 0,          /* bond entries (non-zero elements of `edges`) */
  GIB_HDR_P = 1,          /* rows of the type-grouped entry arrays (groups padded to 128) */
  GIB_HDR_TYPE_COUNT = 2, /* [4] */
  GIB_HDR_TYPE_BASE = 6,  /* [5] */
  GIB_HDR_FLAGS = 11,     /* bit0: a bond with >1 non-zero type; bit1: a bond value != 1; bit2: the batch exceeds the
                             capacity (capacity mode; results of that step are invalid, nothing is written out of bounds) */
  GIB_HDR_CAPACITY = 12,  /* host header only: != 0 = capacity header (E, P are capacities) */
  GIB_HDR_DEV_LO = 13,    /* host header only: address of the device header, low / high 32 bits */
  GIB_HDR_DEV_HI = 14,
  GIB_HDR_INTS = 16
};

const char* gib_last_error(void);
int gib_version(void);
/* tcgen05 3xTF32 GEMM path on (default) / off (fp32 SIMT GEMMs only); process-wide switch */
void gib_set_tensor_cores(int on);
int gib_get_tensor_cores(void);
/* bit 2: narrow outputs (N < 48, the APD heads) on the tensor-core kernel too (default: fp32 SIMT, see gemm_simt.cu).
 * bit 1: no dependent-chain launches (every MLP layer its own launch).
 * bit 0: route the dense GEMMs to the first-generation tcgen05 kernel (operand split through shared memory,
 * gemm_tc.cu) instead of the default second-generation one (activation operand through tensor memory, gemm_tc3.cu);
 * process-wide, A/B measurements only.  Capacity mode needs the default. */
void gib_tc_debug(int mode);
/* K2 scatter-aggregate variant (A/B measurements): bit 0 = two slots per thread, bit 1 = streaming cache hints;
 * default 2 (the fastest at the C4 shape).  Results are identical across variants. */
void gib_scatter_variant(int v);
/* diagnosis: while device_buf != NULL, CTA 0 of every second-generation GEMM launch writes clock64 stamps of its first
 * `tiles` work items into device_buf[tile][16] (int64): 0 MMA start, 1 MMA last issue, 2 / 3 MMA cycles waiting for the
 * TMA tiles / the split operand, 4 epilogue sees the accumulator, 5 accumulator drained, 6 tile stored, 8 / 9 splitter
 * cycles waiting / working (tools/tc3_trace.py) */
void gib_tc_trace(long long* device_buf, int tiles);
/* streaming multiprocessors of the current device (grid sizing of the persistent kernels) */
int gib_device_sm_count(void);

/* ---- K0: edges -> bond entries + CSR.  Replaces summation_mpnn.py:102-118,
 *      aggregation_mpnn.py:105-148, edge_mpnn.py:104-173. ------------------------------- */
size_t gib_graph_count_ws_bytes(const gib_dims* d);
int gib_graph_count(const gib_dims* d, const void* edges, void* count_ws, gib_stream stream);
/* capacity mode: host header for `entry_capacity` bond entries (no device access; valid as long as count_ws lives) */
int gib_graph_header_capacity(const gib_dims* d, int entry_capacity, const void* count_ws, int* hdr_host_out);
size_t gib_graph_bytes(const gib_dims* d, const int* hdr_host);
int gib_graph_fill(const gib_dims* d, const void* edges, void* count_ws, const int* hdr_host, void* graph_buf,
                   gib_stream stream);
/* device addresses of the arrays inside graph_buf, for tests / standalone kernel calls:
 * which = 0 ent_src, 1 ent_dst, 2 ent_w, 3 dst_ptr, 4 dst_ent, 5 src_ptr, 6 src_ent */
void* gib_graph_array(const gib_dims* d, const int* hdr_host, void* graph_buf, int which);

/* ---- parameters: state_dict order of the reference ------------- */
int gib_model_num_params(const gib_dims* d);
long long gib_model_param_numel(const gib_dims* d, int index);
size_t gib_model_packed_bytes(const gib_dims* d);
/* zero-padded + transposed copies of every weight (and 3xTF32 splits when enabled) */
int gib_model_pack(const gib_dims* d, const float* const* params, void* packed, gib_stream stream);

/* ---- whole-model forward / backward.  Replaces SummationMPNN.forward
 *      (summation_mpnn.py:80-149), AggregationMPNN.forward (aggregation_mpnn.py:83-168),
 *      EdgeMPNN.forward (edge_mpnn.py:82-192), the model bodies in mpnn.py and
 *      GraphGather / GlobalReadout (modules.py:39-52, 237-281), and their autograd. ------- */
size_t gib_model_workspace_bytes(const gib_dims* d, const int* hdr_host);
int gib_model_forward(const gib_dims* d, const int* hdr_host, const void* nodes, const void* edges,
                      const void* graph_buf, const void* packed, void* workspace, float* out,
                      gib_stream stream);
size_t gib_model_bwd_scratch_bytes(const gib_dims* d, const int* hdr_host);
/* The backward in two parts on the same scratch: part 1 = readout only -- afterwards the gradients of the gather.* and
 * APDReadout.* parameters (the tail of the parameter order, 79 % of the bytes) are final and a data-parallel caller
 * can start their all-reduce; part 2 = the message passes; part 0 = both (== gib_model_backward). */
int gib_model_backward_part(const gib_dims* d, const int* hdr_host, const void* nodes, const void* edges,
                            const void* graph_buf, const void* packed, const void* workspace, const float* out,
                            const float* dout, float* const* grads, void* scratch, int part, gib_stream stream);
/* grads[i] (same order / shapes as params) are ACCUMULATED into (+=). */
int gib_model_backward(const gib_dims* d, const int* hdr_host, const void* nodes, const void* edges,
                       const void* graph_buf, const void* packed, const void* workspace,
                       const float* out, const float* dout, float* const* grads, void* scratch,
                       gib_stream stream);

/* ---- call-site post-ops (Workflow.py:833-860): KLDivLoss(batchmean)(log_softmax(out),
 *      target / sum(target)) and its gradient w.r.t. `out`, one kernel.
 *      loss_rows[b] = per-molecule KL (sum and divide by B on the caller side or pass
 *      inv_scale); dout = (softmax(out) - t_hat) * grad_scale. ---------------------------- */
int gib_kl_loss_fwd_bwd(const float* out, const float* target, int B, int apd, float grad_scale,
                        float* loss_rows, float* dout, gib_stream stream);

/* out[0] = scale * sum(rows[0..n)) in a fixed order (the batch mean of loss_rows), and a plain asynchronous memset:
 * the two non-model operations of a captured training step (graphinvent_b200/graphed.py) */
int gib_sum_scaled(const float* rows, int n, float scale, float* out, gib_stream stream);
int gib_fill_zero(void* ptr, size_t bytes, gib_stream stream);

/* ---- validation NLL of the "correct" actions (Analyzer.get_validation_likelihood, Analyzer.py:744-758), one kernel:
 *      nll[b] = -log( sum_k softmax(out[b])_k * target[b,k] / sum_k target[b,k] ).  Rows with an all-zero target
 *      give NaN (the reference drops them afterwards, Analyzer.py:756). -------------------------------------- */
int gib_validation_nll(const float* out, const float* target, int B, int apd, float* nll, gib_stream stream);

/* ---- flat-bucket Adam step: replaces torch.optim.Adam.step() on the model parameters (constructed at
 *      Workflow.py:191,221,245, stepped at Workflow.py:795-796; same update rule, L2 weight decay, no amsgrad)
 *      with ONE launch over contiguous params / grads / exp_avg / exp_avg_sq of n floats.  `step` is the
 *      1-based update count (bias corrections are evaluated in double on the host, as the reference does in
 *      Python floats); grads are multiplied by grad_scale first (1/world when the bucket holds an all-reduce
 *      sum).  The four buffers must share their address modulo 16 bytes. ------------------------------------ */
int gib_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                  long long step, double lr, double beta1, double beta2, double eps, double weight_decay,
                  double grad_scale, gib_stream stream);

/* ---- single kernels (unit tests, ncu evidence, reuse) --------------------------------- */
/* Y = act(X W^T + b); X [M, ldx], W packed [Np, Kp] (ldw), Y [M, ldy]; act 0 none / 1 selu / 2 tanh */
int gib_linear_fwd(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy,
                   int M, int N, int K, int act, gib_stream stream);
/* same contract, forced onto the tcgen05 3xTF32 kernel regardless of the size heuristics */
int gib_linear_fwd_tc(const float* X, int ldx, const float* W, int ldw, const float* bias, float* Y, int ldy,
                      int M, int N, int K, int act, gib_stream stream);
/* the model's own call pattern: W arrives as its TF32 hi / lo planes (hi = rna(W), lo = rna(W - hi), as
 * gib_model_pack lays them out), only X is split in the kernel.  m_dev / base_dev (device ints, may be NULL): the live
 * row range [*base_dev, *base_dev + *m_dev) inside buffers of M rows (capacity mode). */
int gib_linear_fwd_tc_planes(const float* X, int ldx, const float* W_hi, const float* W_lo, int ldw,
                             const float* bias, float* Y, int ldy, int M, int N, int K, int act,
                             const int* m_dev, const int* base_dev, gib_stream stream);
/* TF32 hi / lo planes of a row-major matrix (round-to-nearest split), for callers of the entry above */
int gib_split_planes(const float* W, float* W_hi, float* W_lo, long long n, gib_stream stream);
/* dW[R,C] += G^T X, dbias[R] += colsum(G); G [M, ldg], X [M, ldx]; scratch from gib_dw_scratch_bytes */
size_t gib_dw_scratch_bytes(int M, int Nn, int Kk);
int gib_linear_bwd_dw(const float* G, int ldg, int Nn, const float* X, int ldx, int Kk, int M, float* dW,
                      float* dbias, int R, int C, void* scratch, const int* m_dev, const int* base_dev,
                      gib_stream stream);
/* K2 scatter-aggregate: out[s,:] = sum_{q in [ptr[s],ptr[s+1])} w[ent[q]] * msg[ent[q],:]   (w may be NULL) */
int gib_scatter_sum(float* out, const float* msg, int ld, const int* ptr, const int* ent, const float* w,
                    long long S, gib_stream stream);
/* K2' segmented softmax-aggregate (AttentionGGNN) */
int gib_seg_softmax(float* out, const float* EM, const float* EN, int ld, const int* ptr, const int* ent,
                    const float* w, long long S, gib_stream stream);
/* GRU gates: hn = GRU(gi, gh, h) on rows whose CSR segment is non-empty (ptr may be NULL) */
int gib_gru_gates(float* hn, const float* gi, const float* gh, const float* h, int Hp, const int* ptr,
                  long long S, gib_stream stream);
/* GraphGather softmax readout */
int gib_graph_gather(float* g, float* att, const float* en, const float* em, int ld, const int* ptr, int N,
                     int B, float big, gib_stream stream);

/* ---- generation round post-processing: softmax + categorical sample of one
 *      action per molecule from the APD logits, inverse-CDF on a caller-provided uniform. ---- */
int gib_sample_actions(const float* out, int B, int apd, const float* uniforms, int* action,
                       float* likelihood, gib_stream stream);

/* ---- one round of batched graph generation on the device: decode the sampled flat APD
 *      index per slot, validity rules, copy terminated graphs out (pre-action state; terminate-sampled first, then
 *      invalid, ascending), apply add / connect, reset, re-stamp the dummy graph in slot 0.  Replaces
 *      GraphGenerator.get_actions / get_invalid_actions / copy_terminated_graphs / apply_actions / reset_graphs
 *      (GraphGenerator.py:467-657, 340-385, 211-338, 425-465).  6-tuple action layout (no implicit-H / chirality).
 *      State: nodes [B,N,F] f32, edges [B,N,N,Ef] f32, n_nodes [B] i32, likelihoods [B,2N] f32; outputs
 *      gen_* with `capacity` rows; counters[0] = n_generated (in/out), counters[1] = graphs written this round. ---- */
size_t gib_generation_scratch_bytes(int B);
int gib_generation_round(int B, int N, int F, int Ef, int n_atom_types, int n_charges, int round,
                         const int* action, const float* likelihood, float* nodes, float* edges, int* n_nodes,
                         float* likelihoods, float* gen_nodes, float* gen_edges, signed char* gen_n_nodes,
                         float* gen_likelihoods, signed char* properly_terminated, int capacity, int* counters,
                         void* scratch, gib_stream stream);

/* ---- measurement hooks: CUDA-event timing per kernel class on the launching stream.
 *      class 0 = forward/dX launches of the tcgen05 kernel, 1 = its weight-gradient launches, 2 = scatter-aggregate (K2),
 *      3 / 4 = forward/dX and weight-gradient GEMMs on the fp32 SIMT kernels.  GIB_PROFILE_CLASSES entries per array.
 *      work = algorithmic FLOPs (GEMM classes) or bytes (class 2).  Collect after a stream sync. -- */
#define GIB_PROFILE_CLASSES 5
void gib_profile_enable(int on);
long long gib_launch_count(void); /* kernels launched by this library since load */
int gib_profile_collect(double* ms, double* work, long long* count);
/* per-launch records in launch order (before gib_profile_collect, which clears them): returns their number */
int gib_profile_records(double* ms, double* work, int* cls, int cap);

#ifdef __cplusplus
}
#endif
#endif /* GIB200_H */
