// tcgen05 3xTF32 GEMM, second generation: the activation operand is split into its TF32 (hi, lo) pair in REGISTERS and
// handed to the tensor core through TENSOR MEMORY (tcgen05.st -> A-operand-in-TMEM MMA), so it never crosses shared
// memory a second time.  sm_100a.
//
//   NT :  C[M, :N]   = epi( A[M,K] * W[N,K]^T )     A = fp32 activations (row-major, K contiguous), W = a packed weight
//                                                    given as its TF32 hi / lo planes (gib_model_pack)
//   TN :  P[z][n][k] = sum_{m in chunk z} G[m,n] * X[m,k]   (weight-gradient partials; G, X fp32 row-major activations)
//         pb[z][n]   = sum_{m in chunk z} G[m,n]            (bias-gradient partials, free: the G tile is in registers)
//
// Why: with both operands in shared memory a 128x128x8 TF32 MMA reads 8 KB in its 64 cycles -- all of the SM's
// 128 B/clk -- and the first-generation kernel (gemm_tc.cu) added the TMA writes and an in-place hi/lo split on top:
// 192-224 KB of shared-memory traffic per 32-float k-block against 768 MMA cycles (measured: tensor pipe 25 % busy,
// profiles/r01_ncu_full_summary.md).  Here a k-block costs: TMA writes 48 KB (raw A tile, W hi, W lo), one 16 KB read
// by the splitter warps, and 48 KB of W reads by the MMAs = 112 KB = 896 cycles against the same 768.
//
// Roles (448 threads, one persistent CTA per SM, 4-stage ring; smem 3 x 16 KB per stage):
//   warp 0      TMA producer   raw A tile [128 x 32 fp32, 128B swizzle] + W_hi / W_lo tiles (TN: G and X tiles)
//   warps 2-5   splitters      thread = tile row: 8 x LDS.128 (swizzle decoded -> conflict free), hi = x rounded to
//                              nearest at 10 mantissa bits (integer add + mask), lo = x - hi (exact), two
//                              tcgen05.st.32x32b.x32 into the stage's 64 TMEM columns.
//                              TN: thread = column n of G (register transpose for free), masks rows beyond the valid
//                              count, accumulates the bias column sum; the X tile (MN-major B) is split in place by
//                              the epilogue warps, which idle during a work item's k-loop
//   warp 1      MMA issuer     per k-block 8 x tcgen05.mma.kind::tf32 with A in TMEM: a_hi * [W_hi | W_lo] as ONE N = 256
//                              instruction (hi*hi into the main accumulator, hi*lo into the second one) + a_lo * W_hi;
//                              the cross terms' roundings stay away from the large sum (summed in fp32 in the epilogue)
//   warps 6-13  epilogue       drain both accumulators into registers (then the MMA warp may start the next tile),
//                              XOR-swizzled per-warp smem transpose (conflict free, explicit LDS/STS), bias / SELU /
//                              dSELU / residual in the coalesced domain (branch-free), 128-bit global stores of full lines
// TMEM (512 columns): [0,128) main accumulator, [128,256) cross-term accumulator, [256,512) 4 x (A_hi 32 | A_lo 32).
//
// Dynamic row counts: a problem may name device ints (m_dev, base_dev) -- the bond-type group sizes written by K0 --
// instead of host values; tile counts are then computed on the device, so the launch needs no device->host read
// and can be captured in a CUDA graph (include/gib200.h, capacity mode).
#include <cuda.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace gib {

namespace tc3 {

using namespace tcptx;

constexpr int BM = 128, BN = 128, BKF = 32;
constexpr int STAGES = 4;
constexpr int TILE_BYTES = BM * BKF * 4;          // 16 KB
constexpr int STAGE_BYTES = 3 * TILE_BYTES;       // NT: A raw | W hi | W lo      TN: G raw | X raw -> hi | X lo
constexpr int SPL_WARPS = 4, EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * SPL_WARPS + 32 * EPI_WARPS;   // 448
constexpr int EPI_STG_BYTES = 32 * 32 * 4;        // per-warp 32 x 32 staging tile, 128 B rows, XOR-swizzled 16 B chunks
                                                  // (the generic epilogue uses it as 32 x 16)
constexpr int OFF_BARS = STAGES * STAGE_BYTES;
constexpr int OFF_SCHED = OFF_BARS + 256;
constexpr int OFF_STG = OFF_SCHED + 512;
constexpr int SMEM_BYTES = OFF_STG + EPI_WARPS * EPI_STG_BYTES + 1024 /*align slack*/;
constexpr int TMEM_COLS = 512;
constexpr uint32_t ACC_MAIN = 0, ACC_X = 128, A_BASE = 256, A_STAGE_COLS = 64;

constexpr int kMaxChunkRows = 4096;   // <= 512 truncating accumulations per TMEM accumulator and work item
constexpr int MAXP = kTc3MaxProblems;   // 16: e.g. the 5 layers x 3 bond types of a message MLP as one dependent chain

// instruction descriptor, kind::tf32: D = F32 (bits 4-5 = 1), A/B = TF32 (bits 7-9, 10-12 = 2), N >> 3 at bits 17-22,
// M >> 4 at bits 24-28; bit 16 = B is MN-major (TN mode).  A comes from TMEM (always K-major, bit 15 = 0).
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_TN = IDESC | (1u << 16);

struct Maps {   // TMA descriptors in kernel-parameter space
  CUtensorMap a[MAXP];      // NT: activations A (128B swizzle, box 128 x 32)    TN: G (no swizzle, box 32 x 32)
  CUtensorMap b[MAXP];      // NT: W hi plane (128B swizzle)                     TN: X (128B swizzle, 32B atoms, box 32 x 32)
  CUtensorMap b_lo[MAXP];   // NT: W lo plane
};

struct Params {
  GemmNT g[MAXP];           // TN: A = G, B = X, M = rows (capacity when m_dev is set), C = partials of the problem,
                            //     ldc = Kk, N = n_store = n_valid = Kk
  int n_tiles[MAXP];        // column tiles of the output (NT: ceil(N / 128), TN: ceil(Kk / 128))
  int k_blocks[MAXP];       // NT: ceil(K / 32)
  int tn_mt[MAXP];          // TN: ceil(Nn / 128)
  int tn_nn[MAXP];          // TN: Nn
  float* bias_part[MAXP];   // TN: [splits][Nn] partial column sums of G (nullptr: not wanted)
  // dependent chains (NT): problem p reads as its A operand what problem dep[p] writes (the next layer of an MLP);
  // a tile of p at row block i may start once all n_tiles[dep[p]] tiles of row block i of dep[p] are stored.
  // flags[flag_off[p] + i] counts the stored tiles of row block i of problem p (zeroed by the host before the launch).
  int dep[MAXP];
  int flag_off[MAXP];
  int* flags;
  int nprob;
  int chunk_rows;           // TN: reduction rows per work item (multiple of 32)
  long long* trace;         // diagnosis: per-tile clock64 stamps of CTA 0 ([tile][16] int64, gib_tc_trace), else nullptr
  int trace_tiles;
  int diag;                 // timing experiments (gib_tc_debug >> 8; results are wrong with most of them):
                            //   1 no global stores   2 no epilogue after the drain   4 no split / STTM   8 no W_lo tile + MMAs
                            //   16 no MMAs   32 no TMA loads   64 no accumulator drain   128 rotate the k-block order per CTA
                            //   256 interleave the MMAs of the two accumulators   512 round-to-nearest activation split
                            //   1024 TN: no proxy fence after the X split (timing only)
                            //   2048 three accumulators / 4096 N = 256 MMAs (timing only, with the MMA-only switches)
                            //   8192 epilogue without the smem transpose / 16384 without the activation math (timing only)
                            //   65536 splitters with a software pipeline: next operand loaded under the tcgen05.st (correct)
                            //   32768 twelve N = 128 MMAs per k-block instead of 4 x (N = 256 + N = 128) (correct)
                            //   (results stay correct with 128, 256, 512)
};
enum { DG_NO_STORE = 1, DG_NO_EPI = 2, DG_NO_SPLIT = 4, DG_NO_BLO = 8, DG_NO_MMA = 16, DG_NO_TMA = 32, DG_NO_DRAIN = 64,
       DG_ROTATE = 128, DG_INTERLEAVE = 256, DG_RNA_SPLIT = 512, DG_NO_PFENCE = 1024, DG_ACC3 = 2048, DG_N256 = 4096,
       DG_NO_STAGE = 8192, DG_NO_MATH = 16384, DG_MMA12 = 32768, DG_SPLIT_PIPE = 65536 };

struct Sched {              // computed once per CTA from host values or the device-side row counts
  int M[MAXP], base[MAXP], begin[MAXP + 1], splits[MAXP];
};

struct Item { int p, m0, n0, nkb, z, r0, rows; };

template <bool TN>
__device__ __forceinline__ Item decode_item(const Params& P, const Sched& S, int item) {
  Item it;
  int p = 0;
  while (p + 1 < P.nprob && item >= S.begin[p + 1]) ++p;
  const int local = item - S.begin[p];
  it.p = p;
  if constexpr (TN) {
    const int tiles_mn = P.tn_mt[p] * P.n_tiles[p];
    const int tile = local % tiles_mn;
    it.z = local / tiles_mn;
    it.m0 = (tile / P.n_tiles[p]) * BM;
    it.n0 = (tile % P.n_tiles[p]) * BN;
    it.r0 = it.z * P.chunk_rows;
    it.rows = min(S.M[p], it.r0 + P.chunk_rows) - it.r0;
    it.nkb = ceil_div(it.rows, BKF);
  } else {
    it.z = 0; it.r0 = 0; it.rows = 0;
    it.m0 = (local / P.n_tiles[p]) * BM;
    it.n0 = (local % P.n_tiles[p]) * BN;
    it.nkb = P.k_blocks[p];
  }
  return it;
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// chain hand-off: all 8 epilogue warps have stored their part of the tile -> one release increment of the row
// block's counter (named barrier 1 = the 256 epilogue threads)
__device__ __forceinline__ void signal_tile(int* flag) {
  fence_proxy_async_all();
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
  if (threadIdx.x == 32 * (2 + SPL_WARPS)) {
    __threadfence();
    atomicAdd(flag, 1);
  }
}

// round-to-nearest split (weights are split this way once per optimizer step, gib_model_pack): 7 ALU ops per element
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// truncation split, used for the weight-gradient mode's X tile (the B operand, which stays in shared memory): hi = the
// top 19 bits -- the raw tile already IS the hi operand, only lo is written -- lo = x - hi (exact in fp32; the tensor
// core reads its top 19 bits): 2 ALU ops per element.  Error of the pair <= 2^-21 |x|, one-sided.
__device__ __forceinline__ void split_tf32_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// Split for the operand that travels through registers into tensor memory: hi rounded to NEAREST (ties away, what
// cvt.rna does, as integer arithmetic: + half an ulp of the 10-bit mantissa, clear 13 bits; inf stays inf), lo = x - hi
// exact.  With hi = trunc(x), lo always carries the sign of x and the tensor core's truncation of lo is a systematic
// bias towards zero that adds up linearly over K (measured on the shipped checkpoint: logit error 7e-5 -> 1.1e-4 once
// the 500-term output layers ran here); with hi rounded to nearest the sign of lo is random and the truncation of lo
// is unbiased.  3 ALU ops per element instead of 2, no conversion-unit instruction.
__device__ __forceinline__ void split_tf32_rn(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
template <bool RNA> __device__ __forceinline__ void split_act(float x, uint32_t& hi, uint32_t& lo) {
  if (RNA) split_tf32(x, hi, lo);
  else split_tf32_fast(x, hi, lo);
}

// EPI: epilogue specialisation shared by every problem of the launch.  The per-element epilogue must stay a few
// instructions: a generic one (run-time mode / activation / alignment branches for each of the 64 outputs of a thread)
// compiled to ~80 KB of straight-line code, missed the instruction cache on every tile and made the 8 epilogue warps
// the bottleneck of the whole kernel (measured: 236 us with it, 108 us without, profiles/r02_tc3_probe.md).
//   EPI_SPEC_GENERIC   any mode / activation / alignment, chunk by chunk straight from TMEM (accumulators are held
//                      until the tile is stored: rare shapes only -- unaligned or narrow outputs)
//   EPI_SPEC_SELU      C = selu(acc + bias)          EPI_SPEC_LINEAR  C = acc (+ bias)
//   EPI_SPEC_DSELU     C = acc * selu'(aux)          EPI_SPEC_ADD     C = acc + aux
// The specialised ones need 16-byte aligned rows (ldc, ldaux multiples of 4 floats) and n_store == n_valid, a
// multiple of 4 -- the padded-layout contract of every internal buffer.
enum { EPI_SPEC_GENERIC = 0, EPI_SPEC_SELU = 1, EPI_SPEC_LINEAR = 2, EPI_SPEC_DSELU = 3, EPI_SPEC_ADD = 4 };

template <bool TN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc3_gemm_kernel(const __grid_constant__ Maps maps, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full = bars;                    // [STAGES] TMA -> splitters + MMA
  uint64_t* a_full = bars + STAGES;         // [STAGES] splitters -> MMA (A pair in TMEM; TN: X pair in smem too)
  uint64_t* empty = bars + 2 * STAGES;      // [STAGES] MMA -> TMA (smem stage and its TMEM columns are free)
  uint64_t* acc_full = bars + 3 * STAGES;   // MMA -> epilogue
  uint64_t* acc_empty = bars + 3 * STAGES + 1;   // epilogue (accumulators are in registers) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 2);
  Sched& S = *reinterpret_cast<Sched*>(smem + OFF_SCHED);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&a_full[s], TN ? SPL_WARPS + EPI_WARPS : SPL_WARPS);   // one arrival per warp (after __syncwarp)
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int total = 0;
    for (int p = 0; p < P.nprob; ++p) {
      const GemmNT& g = P.g[p];
      const int base = g.base_dev ? __ldg(g.base_dev) : 0;
      int M = g.M;
      if (g.m_dev) {                        // device-side row count inside a buffer of g.M rows
        M = __ldg(g.m_dev);
        if (M > g.M - base) M = g.M - base;
        if (M < 0) M = 0;
      }
      S.M[p] = M; S.base[p] = base; S.begin[p] = total;
      if constexpr (TN) {
        const int s = ceil_div(M, P.chunk_rows);
        S.splits[p] = s;
        total += P.tn_mt[p] * P.n_tiles[p] * s;
      } else {
        S.splits[p] = 1;
        total += ceil_div(M, BM) * P.n_tiles[p];
      }
    }
    for (int p = P.nprob; p <= MAXP; ++p) S.begin[p] = total;
  }
  if (warp == 1) {  // TMEM allocation is a whole-warp operation; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_items = S.begin[MAXP];

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item w = decode_item<TN>(P, S, item);
        const CUtensorMap* map_a = &maps.a[w.p];
        const CUtensorMap* map_b = &maps.b[w.p];
        const int base = S.base[w.p];
        if constexpr (!TN) {
          if (P.flags && P.dep[w.p] >= 0) {       // chain: the row block of the layer below must be complete
            const int d = P.dep[w.p];
            const int* f = P.flags + P.flag_off[d] + w.m0 / BM;
            const int need = P.n_tiles[d];
            const long long t0 = clock64();
            int have;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(f) : "memory");
              if (have < need) {
                __nanosleep(64);
                if (clock64() - t0 > 20000000000LL) __trap();   // ~10 s: a dead-lock, not contention
              }
            } while (have < need);
            fence_proxy_async_all();               // other SMs' generic-proxy stores -> this SM's async-proxy (TMA) reads
          }
        }
        const int rot = (P.diag & DG_ROTATE) ? (int)(blockIdx.x % (unsigned)w.nkb) : 0;
        for (int kb0 = 0; kb0 < w.nkb; ++kb0) {
          const int kb = (kb0 + rot) % w.nkb;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          if (P.diag & DG_NO_TMA) {
            mbar_arrive(&full[stage]);
          } else if constexpr (!TN) {
            const bool blo = !(P.diag & DG_NO_BLO);
            mbar_arrive_expect_tx(&full[stage], (blo ? 3 : 2) * TILE_BYTES);
            tma_load_2d(map_a, &full[stage], st, kb * BKF, base + w.m0);
            tma_load_2d(map_b, &full[stage], st + TILE_BYTES, kb * BKF, w.n0);
            if (blo) tma_load_2d(&maps.b_lo[w.p], &full[stage], st + 2 * TILE_BYTES, kb * BKF, w.n0);
          } else {
            mbar_arrive_expect_tx(&full[stage], 2 * TILE_BYTES);
            const int row = base + w.r0 + kb * BKF;   // 32 reduction rows per stage
#pragma unroll
            for (int j = 0; j < 4; ++j) {             // four 32-float column groups per operand
              tma_load_2d(map_a, &full[stage], st + j * 4096, w.m0 + 32 * j, row);
              tma_load_2d(map_b, &full[stage], st + TILE_BYTES + j * 4096, w.n0 + 32 * j, row);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t tmem_d = tmem_base + ACC_MAIN;   // sum of hi*hi
      const uint32_t tmem_x = tmem_base + ACC_X;      // sum of lo*hi + hi*lo
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int nkb = decode_item<TN>(P, S, item).nkb;
        mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);  // the epilogue holds the previous tile in registers
        tc_fence_after();
        const bool tr = P.trace && blockIdx.x == 0 && it < P.trace_tiles;
        long long w_full = 0, w_afull = 0;
        if (tr) P.trace[it * 16 + 0] = clock64();
        for (int kb = 0; kb < nkb; ++kb) {
          if (tr) {
            const long long t0 = clock64();
            mbar_wait(&full[stage], phase);
            const long long t1 = clock64();
            mbar_wait(&a_full[stage], phase);
            w_full += t1 - t0; w_afull += clock64() - t1;
          } else {
            mbar_wait(&full[stage], phase);              // W (TN: X) tiles landed (async proxy)
            mbar_wait(&a_full[stage], phase);            // A pair is in TMEM (TN: X pair is split in smem)
          }
          tc_fence_after();
          const uint32_t sb = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t a_hi = tmem_base + A_BASE + stage * A_STAGE_COLS, a_lo = a_hi + 32;
          uint64_t d_bhi, d_blo, kstep;
          uint32_t idesc;
          if constexpr (!TN) {
            d_bhi = make_desc(sb + TILE_BYTES); d_blo = make_desc(sb + 2 * TILE_BYTES);
            kstep = 32 >> 4;       // 8 tf32 = 32 B further along K, in 16 B units
            idesc = IDESC;
          } else {
            d_bhi = make_desc_mn(sb + TILE_BYTES); d_blo = make_desc_mn(sb + 2 * TILE_BYTES);
            kstep = 1024 >> 4;     // 8 reduction rows = one 1024 B swizzle atom further
            idesc = IDESC_TN;
          }
          // grouped by accumulator so that consecutive MMAs chain on the same TMEM tile
          if (P.diag & DG_ACC3) {                // experiment (garbage operands): one accumulator per product type
#pragma unroll
            for (int k = 0; k < BKF / 8; ++k) {
              umma_tf32_ts(tmem_base + 0, tmem_base + 384 + 8 * k, d_bhi + k * kstep, idesc, (kb | k) != 0);
              umma_tf32_ts(tmem_base + 128, tmem_base + 416 + 8 * k, d_bhi + k * kstep, idesc, (kb | k) != 0);
              umma_tf32_ts(tmem_base + 256, tmem_base + 384 + 8 * k, d_blo + k * kstep, idesc, (kb | k) != 0);
            }
          } else if (P.diag & DG_N256) {         // experiment (garbage operands): the same FLOPs as 6 N = 256 MMAs
            const uint32_t idesc256 = (idesc & ~(0x3fu << 17)) | ((uint32_t)(256 >> 3) << 17);
#pragma unroll
            for (int k = 0; k < BKF / 8; ++k) {
              umma_tf32_ts(tmem_base + 0, a_hi + 8 * k, d_bhi + k * kstep, idesc256, (kb | k) != 0);
              if (k & 1) umma_tf32_ts(tmem_base + 0, a_lo + 8 * k, d_bhi + k * kstep, idesc256, 1);
            }
          } else if (P.diag & DG_INTERLEAVE) {   // experiment: alternate the two accumulators k-step by k-step
#pragma unroll
            for (int k = 0; k < BKF / 8; ++k) {
              umma_tf32_ts(tmem_d, a_hi + 8 * k, d_bhi + k * kstep, idesc, (kb | k) != 0);
              umma_tf32_ts(tmem_x, a_lo + 8 * k, d_bhi + k * kstep, idesc, (kb | k) != 0);
              umma_tf32_ts(tmem_x, a_hi + 8 * k, d_blo + k * kstep, idesc, 1);
            }
          } else if (!(P.diag & (DG_NO_MMA | DG_MMA12))) {
            // 8 instead of 12 instructions per k-block: the W_hi and W_lo tiles are adjacent in the stage, so ONE N = 256
            // MMA computes a_hi * [W_hi | W_lo]: hi*hi into the main accumulator (columns 0-127) and hi*lo into the
            // cross-term accumulator (columns 128-255); a_lo * W_hi follows as an N = 128 MMA.  A tf32 MMA carries a
            // fixed cost of ~40 cycles per instruction (measured: N = 128: 108 cycles, N = 256: 182).
            const uint32_t idesc256 = (idesc & ~(0x3fu << 17)) | ((uint32_t)(256 >> 3) << 17);
#pragma unroll
            for (int k = 0; k < BKF / 8; ++k) {
              umma_tf32_ts(tmem_d, a_hi + 8 * k, d_bhi + k * kstep, idesc256, (kb | k) != 0);
              umma_tf32_ts(tmem_x, a_lo + 8 * k, d_bhi + k * kstep, idesc, 1);
            }
          } else if (!(P.diag & DG_NO_MMA)) {
#pragma unroll
            for (int k = 0; k < BKF / 8; ++k)
              umma_tf32_ts(tmem_d, a_hi + 8 * k, d_bhi + k * kstep, idesc, (kb | k) != 0);
#pragma unroll
            for (int k = 0; k < BKF / 8; ++k) {
              umma_tf32_ts(tmem_x, a_lo + 8 * k, d_bhi + k * kstep, idesc, (kb | k) != 0);
              if (!(P.diag & DG_NO_BLO)) umma_tf32_ts(tmem_x, a_hi + 8 * k, d_blo + k * kstep, idesc, 1);
            }
          }
          umma_commit(&empty[stage]);                  // frees the smem stage + its TMEM columns when the MMAs retire
          if (kb == nkb - 1) umma_commit(acc_full);    // accumulators complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (tr) { P.trace[it * 16 + 1] = clock64(); P.trace[it * 16 + 2] = w_full; P.trace[it * 16 + 3] = w_afull; }
      }
    }
  } else if (warp < 2 + SPL_WARPS) {
    // ================= splitters: fp32 -> (hi, lo) TF32 pairs, registers -> TMEM =================
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                  // tile row (NT) / G column (TN) this thread owns
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + A_BASE;
    const bool rna = (P.diag & DG_RNA_SPLIT) != 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item w = decode_item<TN>(P, S, item);
      float colsum = 0.f;
      const int itn = (item - (int)blockIdx.x) / (int)gridDim.x;
      const bool tr = P.trace && blockIdx.x == 0 && itn < P.trace_tiles && warp == 2 && lane == 0;
      long long s_wait = 0, s_t0 = 0;
      if (tr) s_t0 = clock64();
      // optional software pipeline (diag 65536): the raw operand of k-block kb + 1 is loaded while the two tcgen05.st of
      // k-block kb are in flight.  The warp handles its k-blocks strictly one after the other, so its per-k-block
      // LATENCY must stay below the MMA time of a k-block; measured, the plain order already does.
      float v[32];
      auto load_raw = [&](int stg_i, uint32_t ph, int kb) {
        if (tr) {
          const long long t0 = clock64();
          mbar_wait(&full[stg_i], ph);
          s_wait += clock64() - t0;
        } else {
          mbar_wait(&full[stg_i], ph);
        }
        tc_fence_after();
        if (P.diag & DG_NO_SPLIT) return;
        const uint32_t sb = smem_u32(smem + stg_i * STAGE_BYTES);
        if constexpr (!TN) {
          // row r of the 128B-swizzled tile: logical 16-byte chunk j sits at chunk (j ^ (r & 7))
          const uint32_t rowaddr = sb + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q4 = lds128(rowaddr + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4));
            v[4 * j + 0] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
          }
        } else {
          // G tile: four unswizzled [32 rows x 32 floats] boxes; this thread reads column r: a free transpose
          const int valid = w.rows - kb * BKF;          // reduction rows of this k-block that exist (>= 1)
          const uint32_t col = sb + (uint32_t)(r >> 5) * 4096 + (uint32_t)(r & 31) * 4;
#pragma unroll
          for (int m = 0; m < 32; ++m) v[m] = lds32(col + m * 128);
          if (valid < 32) {                             // rows past the chunk / the row count may hold anything
#pragma unroll
            for (int m = 0; m < 32; ++m)
              if (m >= valid) v[m] = 0.f;
          }
        }
      };
      const bool pipe = (P.diag & DG_SPLIT_PIPE) != 0;   // measured: NT equal, TN 9 % slower with it (r02_tc3_probe_stage4.txt)
      if (pipe && w.nkb > 0) load_raw(stage, phase, 0);
      for (int kb = 0; kb < w.nkb; ++kb) {
        if (!pipe) load_raw(stage, phase, kb);
        if (P.diag & DG_NO_SPLIT) {
          if (pipe && kb + 1 < w.nkb) load_raw(stage + 1 == STAGES ? 0 : stage + 1, stage + 1 == STAGES ? phase ^ 1 : phase, kb + 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
        uint32_t hi[32], lo[32];
        if constexpr (TN) {
#pragma unroll
          for (int m = 0; m < 32; ++m) colsum += v[m];
        }
        if (rna) {                             // one branch per k-block: per-element selects made ptxas run BOTH splits
#pragma unroll
          for (int m = 0; m < 32; ++m) split_act<true>(v[m], hi[m], lo[m]);
        } else {
#pragma unroll
          for (int m = 0; m < 32; ++m) split_tf32_rn(v[m], hi[m], lo[m]);
        }
        tmem_st32(trow + stage * A_STAGE_COLS, hi);
        tmem_st32(trow + stage * A_STAGE_COLS + 32, lo);
        if (pipe && kb + 1 < w.nkb)
          load_raw(stage + 1 == STAGES ? 0 : stage + 1, stage + 1 == STAGES ? phase ^ 1 : phase, kb + 1);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();                          // every lane's tcgen05.st has completed and is fenced:
        if (lane == 0) mbar_arrive(&a_full[stage]);   // one arrival per warp (128 arrivals per k-block serialise on the barrier)
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if constexpr (TN) {
        float* bp = P.bias_part[w.p];
        if (bp && w.n0 == 0 && w.m0 + r < P.tn_nn[w.p]) bp[(size_t)w.z * P.tn_nn[w.p] + w.m0 + r] = colsum;
      }
      if (tr) { P.trace[itn * 16 + 8] = s_wait; P.trace[itn * 16 + 9] = clock64() - s_t0 - s_wait; }
    }
  } else {
    // ================= epilogue: TMEM -> registers (frees the accumulators) -> smem transpose -> global ============
    const int ew = warp - (2 + SPL_WARPS);     // 0..7
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = ew >> 2;                  // which 64-column half of the tile this warp drains
    const uint32_t stg = smem_u32(smem + OFF_STG + ew * EPI_STG_BYTES);
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const int rr = lane >> 2, c4 = lane & 3;
    const uint32_t st_wr = stg + lane * 64;                          // this lane's staging row (written)
    const uint32_t swz_wr = ((uint32_t)lane >> 1) & 3u;
    int it = 0;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const Item w = decode_item<TN>(P, S, item);
      if constexpr (TN) {
        // the epilogue warps idle during a work item's k-loop: they split the X tile (MN-major B operand) of every
        // k-block -- raw -> hi in place, lo into the sibling tile at the same (swizzled) offset
        const int et = threadIdx.x - 32 * (2 + SPL_WARPS);           // 0..255
        const bool trx = P.trace && blockIdx.x == 0 && it < P.trace_tiles && ew == 0 && lane == 0;
        long long x_wait = 0;
        const long long x_t0 = trx ? clock64() : 0;
        for (int kb = 0; kb < w.nkb; ++kb) {
          if (trx) {
            const long long t0 = clock64();
            mbar_wait(&full[stage], phase);
            x_wait += clock64() - t0;
          } else {
            mbar_wait(&full[stage], phase);
          }
          if (!(P.diag & DG_NO_SPLIT)) {
            const int valid = w.rows - kb * BKF;
            const uint32_t xb = smem_u32(smem + stage * STAGE_BYTES) + TILE_BYTES;
#pragma unroll
            for (int i = 0; i < TILE_BYTES / 16 / (32 * EPI_WARPS); ++i) {
              const uint32_t c = (uint32_t)(et + i * 32 * EPI_WARPS);   // 16-byte chunk index; 8 chunks per 128 B row
              const int m = (int)((c & 255u) >> 3);                     // reduction row inside its 4 KB box
              float4 v = lds128(xb + c * 16);
              if (m >= valid) v = make_float4(0.f, 0.f, 0.f, 0.f);      // rows past the chunk / the row count
              uint32_t h[4], l[4];
              if (P.diag & DG_RNA_SPLIT) {
                split_act<true>(v.x, h[0], l[0]); split_act<true>(v.y, h[1], l[1]);
                split_act<true>(v.z, h[2], l[2]); split_act<true>(v.w, h[3], l[3]);
              } else {
                split_act<false>(v.x, h[0], l[0]); split_act<false>(v.y, h[1], l[1]);
                split_act<false>(v.z, h[2], l[2]); split_act<false>(v.w, h[3], l[3]);
              }
              // truncation split: the raw tile already IS the hi operand (the tensor core ignores the 13 low bits)
              if ((P.diag & DG_RNA_SPLIT) || m >= valid)
                sts128(xb + c * 16, make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]),
                                                __uint_as_float(h[3])));
              sts128(xb + TILE_BYTES + c * 16, make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]),
                                                           __uint_as_float(l[2]), __uint_as_float(l[3])));
            }
            if (!(P.diag & DG_NO_PFENCE))
              fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor-core (async) proxy
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (trx) { P.trace[it * 16 + 10] = x_wait; P.trace[it * 16 + 11] = clock64() - x_t0 - x_wait; }
      }
      const GemmNT& g = P.g[w.p];
      const int Mrows = TN ? P.tn_nn[w.p] : S.M[w.p];
      const size_t row_base = TN ? (size_t)0 : (size_t)S.base[w.p];
      float* const Cbase = g.C + (TN ? (size_t)w.z * P.tn_nn[w.p] * g.ldc : row_base * g.ldc);
      const float* const Xbase = g.aux ? g.aux + row_base * g.ldaux : nullptr;
      const int ldc = g.ldc, ldaux = g.ldaux, n_store = g.n_store;
      const int mrow0 = w.m0 + q * 32 + rr;                          // first of this lane's 4 output rows (stride 8)
      const int ncol0 = w.n0 + half * 64 + c4 * 4;                   // first of this lane's 4 column groups (stride 16)
      mbar_wait(acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      const bool tr = P.trace && blockIdx.x == 0 && it < P.trace_tiles && ew == 0 && lane == 0;
      if (tr) P.trace[it * 16 + 4] = clock64();
      if constexpr (EPI != EPI_SPEC_GENERIC) {
        float acc[64];
        if (P.diag & DG_NO_DRAIN) {
#pragma unroll
          for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        } else {
#pragma unroll
          for (int chunk = 0; chunk < 4; ++chunk) {
            uint32_t r1[16], r2[16];
            tmem_ld16(tq + ACC_MAIN + half * 64 + chunk * 16, r1);
            tmem_ld16(tq + ACC_X + half * 64 + chunk * 16, r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[chunk * 16 + i] = __uint_as_float(r1[i]) + __uint_as_float(r2[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);   // the MMA warp may overwrite the accumulators now
        if (tr) P.trace[it * 16 + 5] = clock64();
        const float* const bias = g.bias;
        // 32-column chunks: a staging row is one 128-byte line, so every 128-bit store instruction of the warp writes
        // 4 rows x 128 contiguous bytes = 4 full lines (16-column chunks wrote 8 half lines per instruction)
        const int rr8 = lane >> 3, c8 = lane & 7;
        const uint32_t st_wr32 = stg + lane * 128;
        const uint32_t swz32 = (uint32_t)lane & 7u;
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
          if (P.diag & DG_NO_EPI) break;
          // the auxiliary operand (layer output / residual) of the 8 rows this lane finishes: all 8 loads are issued
          // before the staging round trip -- inside the store loop each load sat behind the previous row's store
          // (possible alias) and the warp paid a full global-memory latency per row
          float4 xx[8];
          if constexpr (EPI == EPI_SPEC_DSELU || EPI == EPI_SPEC_ADD) {
            const int n = w.n0 + half * 64 + chunk * 32 + c8 * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int m = w.m0 + q * 32 + i * 4 + rr8;
              xx[i] = (n < n_store && m < Mrows) ? __ldg(reinterpret_cast<const float4*>(Xbase + (size_t)m * ldaux + n))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          __syncwarp();
          // lane = tile row; 16-byte chunk j of row `lane` sits at chunk (j ^ (lane & 7)) -- the 128B-swizzle pattern:
          // the 8 rows of a write phase and the 8 chunks of a read phase both cover all 32 banks exactly once
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8)
            if (!(P.diag & DG_NO_STAGE)) sts128(st_wr32 + (((uint32_t)j8 ^ swz32) << 4),
                   make_float4(acc[chunk * 32 + j8 * 4 + 0], acc[chunk * 32 + j8 * 4 + 1],
                               acc[chunk * 32 + j8 * 4 + 2], acc[chunk * 32 + j8 * 4 + 3]));
          __syncwarp();
          const int n = w.n0 + half * 64 + chunk * 32 + c8 * 4;
          if (n < n_store) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (EPI == EPI_SPEC_SELU || EPI == EPI_SPEC_LINEAR)
              if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
            float4 vv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {      // all 8 staging reads in flight before any arithmetic
              const int row = i * 4 + rr8;
              if (P.diag & DG_NO_STAGE) vv[i] = make_float4(acc[chunk * 32 + 4 * i], acc[chunk * 32 + 4 * i + 1],
                                                            acc[chunk * 32 + 4 * i + 2], acc[chunk * 32 + 4 * i + 3]);
              else vv[i] = lds128(stg + row * 128 + (((uint32_t)c8 ^ ((uint32_t)row & 7u)) << 4));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + rr8;
              const int m = w.m0 + q * 32 + row;
              if (m < Mrows) {
                float4 v = vv[i];
                if (P.diag & DG_NO_MATH) {
                } else if constexpr (EPI == EPI_SPEC_SELU) {
                  v.x = act_fast(v.x + b4.x, ACT_SELU); v.y = act_fast(v.y + b4.y, ACT_SELU);
                  v.z = act_fast(v.z + b4.z, ACT_SELU); v.w = act_fast(v.w + b4.w, ACT_SELU);
                } else if constexpr (EPI == EPI_SPEC_LINEAR) {
                  v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                } else {
                  const float4 x = xx[i];
                  if constexpr (EPI == EPI_SPEC_DSELU) {
                    v.x *= dselu_from_out(x.x); v.y *= dselu_from_out(x.y);
                    v.z *= dselu_from_out(x.z); v.w *= dselu_from_out(x.w);
                  } else {
                    v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
                  }
                }
                if (!(P.diag & DG_NO_STORE)) *reinterpret_cast<float4*>(Cbase + (size_t)m * ldc + n) = v;
              }
            }
          }
        }
      } else {
        // generic: one 16-column chunk at a time straight from TMEM (no 64-register drain, compact code)
        const bool vec_c = (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
        const bool vec_x = g.aux && (ldaux & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.aux) & 15) == 0);
        const int mode = g.mode, act = g.act, n_valid = g.n_valid, Ncols = g.N;
#pragma unroll 1
        for (int chunk = 0; chunk < 4; ++chunk) {
          uint32_t r1[16], r2[16];
          tmem_ld16(tq + ACC_MAIN + half * 64 + chunk * 16, r1);
          tmem_ld16(tq + ACC_X + half * 64 + chunk * 16, r2);
          tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4)
            sts128(st_wr + (((uint32_t)j4 ^ swz_wr) << 4),
                   make_float4(__uint_as_float(r1[j4 * 4 + 0]) + __uint_as_float(r2[j4 * 4 + 0]),
                               __uint_as_float(r1[j4 * 4 + 1]) + __uint_as_float(r2[j4 * 4 + 1]),
                               __uint_as_float(r1[j4 * 4 + 2]) + __uint_as_float(r2[j4 * 4 + 2]),
                               __uint_as_float(r1[j4 * 4 + 3]) + __uint_as_float(r2[j4 * 4 + 3])));
          __syncwarp();
          const int n = ncol0 + chunk * 16;
          if (n < n_store && !(P.diag & DG_NO_EPI)) {
            float bj[4] = {0.f, 0.f, 0.f, 0.f};
            if (mode == EPI_ACT && g.bias) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n + j < Ncols) bj[j] = __ldg(g.bias + n + j);
            }
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int row = i * 8 + rr;
              const int m = mrow0 + i * 8;
              if (m >= Mrows) continue;
              const float4 a4 = lds128(stg + row * 64 + (((uint32_t)c4 ^ (((uint32_t)row >> 1) & 3u)) << 4));
              float v[4] = {a4.x, a4.y, a4.z, a4.w};
              float x[4] = {0.f, 0.f, 0.f, 0.f};
              if (mode != EPI_ACT) {
                const float* ax = Xbase + (size_t)m * ldaux + n;
                if (vec_x && n + 3 < n_store) {
                  const float4 t4 = *reinterpret_cast<const float4*>(ax);
                  x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w;
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (n + j < n_store) x[j] = ax[j];
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (mode == EPI_ACT) v[j] = act_fast(v[j] + bj[j], act);
                else if (mode == EPI_MUL_DACT) v[j] = v[j] * dact_from_out(x[j], act);
                else v[j] = v[j] + x[j];
                if (n + j >= n_valid) v[j] = 0.f;
              }
              float* dst = Cbase + (size_t)m * ldc + n;
              if (vec_c && n + 3 < n_store) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (n + j < n_store) dst[j] = v[j];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
      if constexpr (!TN)
        if (P.flags) signal_tile(P.flags + P.flag_off[w.p] + w.m0 / BM);
      if (tr) P.trace[it * 16 + 6] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---- fixed-order reduction of the split partials into the gradient tensors, all problems of a group in one launch.
//      blocks [blk_w[p], blk_b[p]):    dW_p[r*rs + c*cs] += sum_z ws_p[z][prow(r)][c]   (one thread per element)
//      blocks [blk_b[p], blk_w[p+1]):  db_p[r]           += sum_z wsb_p[z][prow(r)]     (32 rows per block)
//      z runs over the problem's split count, recomputed here from the same host / device row count the GEMM used.
struct ReduceProb {
  const float* ws; const float* wsb; float* dW; float* db;
  const int* m_dev; const int* base_dev; int M;
  int Nn, Kk, R, C, Rb, Rbp;
  long long rs, cs;
  int blk_w, blk_b;
};
struct ReduceTable { ReduceProb q[MAXP]; int n, total_blocks, chunk_rows; };

__global__ void __launch_bounds__(256) reduce_grads3_kernel(const ReduceTable T) {
  int p = 0;
  while (p + 1 < T.n && (int)blockIdx.x >= T.q[p + 1].blk_w) ++p;
  const ReduceProb& q = T.q[p];
  int M = q.M;
  if (q.m_dev) {
    const int base = q.base_dev ? __ldg(q.base_dev) : 0;
    M = __ldg(q.m_dev);
    if (M > q.M - base) M = q.M - base;
    if (M < 0) M = 0;
  }
  const int splits = ceil_div(M, T.chunk_rows);
  if ((int)blockIdx.x < q.blk_b) {
    const long long idx = (long long)((int)blockIdx.x - q.blk_w) * 256 + threadIdx.x;
    if (idx >= (long long)q.R * q.C) return;
    const int r = (int)(idx / q.C), c = (int)(idx % q.C);
    const int prow = (r / q.Rb) * q.Rbp + (r % q.Rb);
    const float* src = q.ws + (size_t)prow * q.Kk + c;
    const size_t stride = (size_t)q.Nn * q.Kk;
    float s = 0.f;
    int zi = 0;
    for (; zi + 8 <= splits; zi += 8) {      // 8 independent loads in flight, summed in ascending z
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(zi + u) * stride];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; zi < splits; ++zi) s += src[(size_t)zi * stride];
    q.dW[r * q.rs + c * q.cs] += s;
  } else {
    __shared__ float sm[8][33];
    const int r = ((int)blockIdx.x - q.blk_b) * 32 + (threadIdx.x & 31);
    const int wy = threadIdx.x >> 5;
    float s = 0.f;
    if (r < q.R) {
      const int prow = (r / q.Rb) * q.Rbp + (r % q.Rb);
      for (int zi = wy; zi < splits; zi += 8) s += q.wsb[(size_t)zi * q.Nn + prow];
    }
    sm[wy][threadIdx.x & 31] = s;
    __syncthreads();
    if (wy == 0 && r < q.R) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
      q.db[r] += t;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Encoded descriptors are pure functions of (base, rows, cols, ld, box, swizzle): the training step presents the same
// few hundred operands every iteration, so they are memoised (an encode costs ~1 us of host time, 3-12 per launch).
struct MapKey {
  const void* base; int rows, cols, ld, box_rows, swz;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && swz == o.swz;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    h = h * 1000003u ^ (size_t)k.rows;
    h = h * 1000003u ^ (size_t)k.cols;
    h = h * 1000003u ^ (size_t)k.ld;
    h = h * 1000003u ^ (size_t)(k.box_rows * 8 + k.swz);
    return h;
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

static int make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows, CUtensorMapSwizzle swz) {
  const MapKey key{base, rows, cols, ld, box_rows, (int)swz};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { memcpy(map, &it->second, sizeof(CUtensorMap)); return 0; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return -4; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BKF, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box_rows=%d swz=%d", (int)r, rows, cols, ld,
              box_rows, (int)swz);
    return -4;
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_map_cache.size() > 16384) g_map_cache.clear();   // bounded: a long generation run sees many batch sizes
  g_map_cache.emplace(key, *map);
  return 0;
}

long long* g_trace = nullptr;   // gib_tc_trace
int g_trace_tiles = 0;

struct DevInfo { int num_sms = 0; bool attr_done = false; };
static std::mutex g_dev_mu;
static DevInfo g_dev[64];

static int prepare(int* num_sms_out) {
  int dev = 0;
  GIB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return -4; }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevInfo& d = g_dev[dev];
  if (!d.attr_done) {   // function attributes are per device
    GIB_CUDA_TRY(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc3_gemm_kernel<false, EPI_SPEC_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc3_gemm_kernel<false, EPI_SPEC_SELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc3_gemm_kernel<false, EPI_SPEC_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc3_gemm_kernel<false, EPI_SPEC_DSELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc3_gemm_kernel<false, EPI_SPEC_ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc3_gemm_kernel<true, EPI_SPEC_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    d.attr_done = true;
  }
  *num_sms_out = d.num_sms;
  return 0;
}

// which compact epilogue a problem can use (EPI_SPEC_GENERIC: none)
static int epi_spec(const GemmNT& p) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if ((p.ldc & 3) || !al(p.C) || (p.n_store & 3) || p.n_valid != p.n_store) return EPI_SPEC_GENERIC;
  if (p.mode == EPI_ACT) {
    if (p.bias && (!al(p.bias) || p.N < p.n_store)) return EPI_SPEC_GENERIC;
    if (p.act == ACT_SELU) return EPI_SPEC_SELU;
    if (p.act == ACT_NONE) return EPI_SPEC_LINEAR;
    return EPI_SPEC_GENERIC;
  }
  if (!p.aux || (p.ldaux & 3) || !al(p.aux)) return EPI_SPEC_GENERIC;
  if (p.mode == EPI_MUL_DACT) return p.act == ACT_SELU ? EPI_SPEC_DSELU : EPI_SPEC_GENERIC;
  return EPI_SPEC_ADD;
}

}  // namespace tc3

void tc3_set_trace(long long* buf, int tiles) { tc3::g_trace = buf; tc3::g_trace_tiles = tiles; }

int device_sm_count() {
  int n = 0;
  if (tc3::prepare(&n) != 0 || n <= 0) n = 148;
  return n;
}

bool tc3_eligible(const GemmNT& p) {
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return p.M >= 1 && p.N >= 1 && p.K >= 16 && (p.K % 16) == 0 && (p.lda % 4) == 0 && (p.ldb % 4) == 0 && p.B_hi &&
         p.B_lo && al(p.A) && al(p.B_hi) && al(p.B_lo);
}

// up to MAXP NT problems in one persistent launch (weights as pre-split TF32 planes); dep == nullptr: independent
static int launch_nt(const GemmNT* ps, const int* dep, int n, int* flags, cudaStream_t st) {
  using namespace tc3;
  if (n < 1 || n > MAXP) { set_error("gemm_nt_tc3: %d problems (max %d)", n, MAXP); return -2; }
  int num_sms = 0;
  GIB_TRY(prepare(&num_sms));
  Maps maps;
  Params P;
  memset(&P, 0, sizeof(P));
  double work = 0;
  long long tiles = 0;
  int np = 0;
  int spec = -1;
  int slot[MAXP];                 // input index -> launch slot (-1: skipped, no rows)
  int flag_ints = 0;
  for (int i = 0; i < n; ++i) {
    const GemmNT& p = ps[i];
    slot[i] = -1;
    if (p.M <= 0 || p.N <= 0) continue;
    if (!tc3_eligible(p)) { set_error("gemm_nt_tc3: operands violate the TMA alignment / pre-split contract"); return -2; }
    const int sp = epi_spec(p);
    spec = (spec < 0 || spec == sp) ? sp : EPI_SPEC_GENERIC;     // one epilogue specialisation per launch
    GIB_TRY(make_map(&maps.a[np], p.A, p.M, p.K, p.lda, BM, CU_TENSOR_MAP_SWIZZLE_128B));
    GIB_TRY(make_map(&maps.b[np], p.B_hi, p.N, p.K, p.ldb, BN, CU_TENSOR_MAP_SWIZZLE_128B));
    GIB_TRY(make_map(&maps.b_lo[np], p.B_lo, p.N, p.K, p.ldb, BN, CU_TENSOR_MAP_SWIZZLE_128B));
    P.g[np] = p;
    P.n_tiles[np] = ceil_div(p.N, BN);
    P.k_blocks[np] = ceil_div(p.K, BKF);
    P.dep[np] = -1;
    P.flag_off[np] = flag_ints;
    flag_ints += ceil_div(p.M, BM) + 1;                         // row blocks of the (capacity) row count
    if (dep && dep[i] >= 0) {
      if (dep[i] >= i || slot[dep[i]] < 0) { set_error("gemm_nt_tc3_chain: problem %d depends on %d", i, dep[i]); return -2; }
      const GemmNT& d = ps[dep[i]];
      if (d.M != p.M || d.m_dev != p.m_dev || d.base_dev != p.base_dev || d.C != p.A) {
        set_error("gemm_nt_tc3_chain: problem %d does not consume the rows problem %d produces", i, dep[i]);
        return -2;
      }
      P.dep[np] = slot[dep[i]];
    }
    slot[i] = np;
    tiles += (long long)ceil_div(p.M, BM) * P.n_tiles[np];      // upper bound when the row count lives on the device
    work += p.work > 0 ? p.work : 2.0 * p.M * (double)p.N * p.K;
    ++np;
  }
  if (np == 0) return 0;
  P.nprob = np;
  P.diag = g_tc_debug >> 8;
  P.trace = g_trace; P.trace_tiles = g_trace_tiles;
  if (dep) {
    P.flags = flags;
    GIB_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)flag_ints * sizeof(int), st));
  }
  const int grid = (int)(tiles < num_sms ? tiles : num_sms);
  ProfScope prof(PROF_GEMM_NT, work, st);
  switch (spec) {
    case EPI_SPEC_SELU: tc3_gemm_kernel<false, EPI_SPEC_SELU><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, P); break;
    case EPI_SPEC_LINEAR: tc3_gemm_kernel<false, EPI_SPEC_LINEAR><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, P); break;
    case EPI_SPEC_DSELU: tc3_gemm_kernel<false, EPI_SPEC_DSELU><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, P); break;
    case EPI_SPEC_ADD: tc3_gemm_kernel<false, EPI_SPEC_ADD><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, P); break;
    default: tc3_gemm_kernel<false, EPI_SPEC_GENERIC><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, P); break;
  }
  GIB_LAUNCH_CHECK();
  return 0;
}

int gemm_nt_tc3_group(const GemmNT* ps, int n, cudaStream_t st) { return launch_nt(ps, nullptr, n, nullptr, st); }

size_t tc3_chain_flag_ints(const GemmNT* ps, int n) {
  size_t f = 0;
  for (int i = 0; i < n; ++i) f += (size_t)ceil_div(ps[i].M > 0 ? ps[i].M : 0, tc3::BM) + 1;
  return f;
}

int gemm_nt_tc3_chain(const GemmNT* ps, const int* dep, int n, int* flags, cudaStream_t st) {
  if (!flags) { set_error("gemm_nt_tc3_chain: no flag buffer"); return -2; }
  return launch_nt(ps, dep, n, flags, st);
}

// ---- weight gradients --------------------------------------------------------------------------------------------
bool tc3_dw_eligible(const GemmDW& q) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return q.dW && q.M >= 1 && q.Nn >= 16 && q.Kk >= 16 && (q.Nn % 4) == 0 && (q.Kk % 4) == 0 && (q.ldg % 4) == 0 &&
         (q.ldx % 4) == 0 && al(q.G) && al(q.X);
}

// Reduction rows per work item for a group: about one work item per SM, within [256, kMaxChunkRows].
int tc3_dw_chunk_rows(const GemmDW* qs, int n, long long plan_rows) {
  using namespace tc3;
  const int num_sms = device_sm_count();
  int max_tiles = 1;
  long long rows = 0;
  for (int i = 0; i < n; ++i) {
    if (qs[i].M <= 0) continue;
    max_tiles = std::max(max_tiles, ceil_div(qs[i].Nn, BM) * ceil_div(qs[i].Kk, BN));
    rows += qs[i].M;
  }
  if (plan_rows > 0) rows = plan_rows;
  // whole rounds of work items: r rounds of num_sms items, the fewest rounds whose chunk respects the cap (a capped
  // chunk that leaves a few items for an extra round would double the kernel time)
  long long c = 8 * BKF;
  for (int r = 1; r <= 64; ++r) {
    c = ceil_div_ll(rows * max_tiles, (long long)r * num_sms);
    c = ceil_div_ll(c, BKF) * BKF;
    if (c <= kMaxChunkRows) break;                    // the tensor core accumulates with truncation: bounded chains
  }
  if (c > kMaxChunkRows) c = kMaxChunkRows;
  if (c < 8 * BKF) c = 8 * BKF;                       // >= 256 reduction rows per item: amortise the 64 KB tile drain
  return (int)c;
}

// Scratch layout of a group: per problem [cap_splits][Nn][Kk] partial products, then [cap_splits][Nn] bias partials.
void tc3_dw_layout(const GemmDW* qs, int n, int chunk_rows, Dw3Layout* L) {
  using namespace tc3;
  L->chunk_rows = chunk_rows;
  size_t off = 0;
  for (int i = 0; i < MAXP; ++i) {
    if (i < n) {
      const int s = std::max(1, ceil_div(qs[i].M, L->chunk_rows));
      L->cap_splits[i] = s;
      L->part_off[i] = off; off += (size_t)s * qs[i].Nn * qs[i].Kk;
      L->bias_off[i] = off; off += (size_t)s * qs[i].Nn;
      off = (off + 31) & ~(size_t)31;
    } else {
      L->cap_splits[i] = 0; L->part_off[i] = L->bias_off[i] = off;
    }
  }
  L->floats = off;
}

// partial products + bias partials of up to MAXP weight-gradient problems in one launch (main stream)
int gemm_dw_tc3_partials(const GemmDW* qs, int n, const Dw3Layout& L, float* scratch, cudaStream_t st) {
  using namespace tc3;
  if (n < 1 || n > MAXP) { set_error("gemm_dw_tc3_partials: %d problems (max %d)", n, MAXP); return -2; }
  int num_sms = 0;
  GIB_TRY(prepare(&num_sms));
  Maps maps;
  Params P;
  memset(&P, 0, sizeof(P));
  long long items = 0;
  for (int i = 0; i < n; ++i) {
    const GemmDW& q = qs[i];
    if (!tc3_dw_eligible(q)) { set_error("gemm_dw_tc3: operands violate the TMA alignment contract"); return -2; }
    GIB_TRY(make_map(&maps.a[i], q.G, q.M, q.Nn, q.ldg, BKF, CU_TENSOR_MAP_SWIZZLE_NONE));          // 32 x 32 boxes
    GIB_TRY(make_map(&maps.b[i], q.X, q.M, q.Kk, q.ldx, BKF, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    GemmNT& g = P.g[i];
    g = GemmNT();
    g.A = q.G; g.lda = q.ldg; g.B = q.X; g.ldb = q.ldx;
    g.M = q.M; g.m_dev = q.m_dev; g.base_dev = q.base_dev;
    g.C = scratch + L.part_off[i]; g.ldc = q.Kk; g.N = q.Kk; g.n_store = q.Kk; g.n_valid = q.Kk;
    g.mode = EPI_ACT; g.act = ACT_NONE; g.bias = nullptr;
    P.n_tiles[i] = ceil_div(q.Kk, BN);
    P.tn_mt[i] = ceil_div(q.Nn, BM);
    P.tn_nn[i] = q.Nn;
    P.bias_part[i] = q.dbias ? scratch + L.bias_off[i] : nullptr;
    items += (long long)P.tn_mt[i] * P.n_tiles[i] * L.cap_splits[i];
  }
  P.nprob = n;
  P.chunk_rows = L.chunk_rows;
  P.diag = g_tc_debug >> 8;
  P.trace = g_trace; P.trace_tiles = g_trace_tiles;
  const int grid = (int)(items < num_sms ? items : num_sms);
  tc3_gemm_kernel<true, EPI_SPEC_LINEAR><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, P);
  GIB_LAUNCH_CHECK();
  return 0;
}

// fixed-order reduction of the group's partials into the gradient tensors (any stream ordered after the partials)
int gemm_dw_tc3_reduce(const GemmDW* qs, int n, const Dw3Layout& L, const float* scratch, cudaStream_t st) {
  using namespace tc3;
  ReduceTable T;
  memset(&T, 0, sizeof(T));
  int blk = 0;
  for (int i = 0; i < n; ++i) {
    const GemmDW& q = qs[i];
    ReduceProb& r = T.q[i];
    r.ws = scratch + L.part_off[i]; r.wsb = scratch + L.bias_off[i]; r.dW = q.dW; r.db = q.dbias;
    r.m_dev = q.m_dev; r.base_dev = q.base_dev; r.M = q.M;
    r.Nn = q.Nn; r.Kk = q.Kk; r.R = q.R; r.C = q.C; r.Rb = q.Rb; r.Rbp = q.Rbp; r.rs = q.rs; r.cs = q.cs;
    r.blk_w = blk; blk += q.dW ? (int)ceil_div_ll((long long)q.R * q.C, 256) : 0;
    r.blk_b = blk; blk += q.dbias ? ceil_div(q.R, 32) : 0;
  }
  T.n = n; T.total_blocks = blk; T.chunk_rows = L.chunk_rows;
  if (blk == 0) return 0;
  reduce_grads3_kernel<<<blk, 256, 0, st>>>(T);
  GIB_LAUNCH_CHECK();
  return 0;
}

}  // namespace gib
