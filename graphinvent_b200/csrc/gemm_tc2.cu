// tcgen05 GEMM, CTA-pair edition (cta_group::2): C[M, :N] = epi( A[M,K] * B[N,K]^T ), fp32-accurate 3xTF32.
//
// STATUS: candidate for the forward / dX GEMMs of gemm_tc.cu -- compiled into the library, selected only by
// gib_tc_debug bit 7 (128), never by default.  Written when no GPU time was left in round 1: NOT yet run on a B200.
//
// Why: the single-CTA kernel (gemm_tc.cu) is bounded by shared-memory traffic, not by the tensor pipe.
// A 128x128x32 k-block moves 192 KB through one SM's shared memory (TMA fill 48 KB, splitter 16 + 32 KB, three
// MMAs x (A 16 KB + B 16 KB)) = 1536 cycles at 128 B/clk against 768 MMA cycles.  With a CTA pair on one TPC the
// MMA is 256 x 128: each CTA keeps its own 128 rows of A and only HALF of the B tile (64 weight rows); the tensor
// cores of both SMs read each half once.  Per SM and k-block: TMA fill 32 KB, splitter 16 + 16 KB (with
// gib_tc_debug bit 6 the raw tile is the hi operand -- the MMA reads its top 19 bits -- and only the rounded
// remainder is written; otherwise hi is rounded in place, + 16 KB), MMA reads 3 x (16 + 8) KB: 136 KB = 1088 cycles,
// and the weight traffic L2 -> SM halves.  The freed 16 KB per stage buy a 4th pipeline stage.
//
// Roles per CTA (same thread layout as gemm_tc.cu): warp 0 TMA producer (own A rows + own half of B_hi / B_lo),
// warps 2-5 splitters (A only; weights arrive pre-split from the packed arena), warp 1 MMA issuer (leader CTA only;
// in the peer it just owns the TMEM allocation), warps 6-13 epilogue (each CTA drains its own 128 accumulator rows).
// Cross-CTA hand-offs:  splitters -> leader's full_split (remote mbarrier.arrive, one per warp);
//                       tcgen05.commit.cta_group::2 multicast -> both CTAs' empty[stage] / acc_full[acc];
//                       epilogue warps -> leader's acc_empty (remote arrive).
// Weight-gradient (TN) mode as in gemm_tc.cu: P[z][n][k] = sum_{m in chunk z} G[m,n] X[m,k]; the activations are the
// MN-major operands themselves, the pair covers 256 columns of G x 128 columns of X (each CTA: 128 / 64), both
// operands are split in the kernel (raw tile = hi, written: lo).
// Every wait is bounded (trap, never a hang).
#include <cuda.h>
#include <string.h>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace gib {

namespace tc2 {

using namespace tcptx;

constexpr int BM = 128;                  // accumulator rows per CTA (256 per pair)
constexpr int BN = 128;                  // output columns per tile
constexpr int BNH = BN / 2;              // weight rows each CTA stages
constexpr int BKF = 32;                  // floats of K per stage (one 128-byte swizzle row)
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BKF * 4;    // 16 KB
constexpr int BH_BYTES = BNH * BKF * 4;  //  8 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;   // A_hi(raw) | A_lo | B_hi half | B_lo half = 48 KB
constexpr int EPI_WARPS = 8;
constexpr int EPI_LD = 20;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + EPI_WARPS * 32 * EPI_LD * 4;
constexpr int NUM_THREADS = 192 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;           // 2 accumulator stages x (hi*hi + cross terms) x 128 columns
constexpr int MAXP = 4;

static_assert(SMEM_BYTES <= 232448, "dynamic shared memory budget of sm_100a");

// instruction descriptor as in gemm_tc.cu, M = 256 (the pair), N = 128, both operands K-major
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr uint32_t IDESC2_TN = IDESC2 | (1u << 15) | (1u << 16);   // A and B MN-major

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster (cutlass ClusterBarrier::arrive)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of the pair -> arrive on the barrier at this offset in both CTAs
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      :
      : "r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}

struct Maps {
  CUtensorMap a[MAXP];      // NT: raw fp32 activations, box 128 rows x 32 floats.  TN: G, box 32 rows x 32 floats
  CUtensorMap b_hi[MAXP];   // NT: TF32 hi plane of the packed weights, box 64 rows x 32 floats.  TN: X, box 32 x 32
  CUtensorMap b_lo[MAXP];   // NT: lo plane
};

struct Params {
  GemmNT g[MAXP];
  int m_pairs[MAXP], n_tiles[MAXP], k_blocks[MAXP], item_begin[MAXP + 1];
  int nprob;
  int debug;
  // TN (weight-gradient) mode: problem 0 only; g[0].C = split-K workspace [splits][tn_nn][tn_kk]
  int tn, splits, chunk_rows, tn_rows, tn_nn, tn_kk;
  long long* timing;   // TIMING build only: same [mode][cta][slot] accumulation table as gemm_tc.cu (gemm.cuh)
};

struct Item { int p, m0, n0, nkb, z; };   // m0 = first row of the PAIR's 256-row tile

__device__ __forceinline__ Item decode_item(const Params& P, int item) {
  if (P.tn) {
    const int tiles = P.m_pairs[0] * P.n_tiles[0];
    const int t = item % tiles;
    Item it;
    it.p = 0; it.z = item / tiles;
    it.m0 = (t / P.n_tiles[0]) * (2 * BM);
    it.n0 = (t % P.n_tiles[0]) * BN;
    const int r0 = it.z * P.chunk_rows;
    it.nkb = ceil_div(min(P.tn_rows, r0 + P.chunk_rows) - r0, BKF);
    return it;
  }
  int p = 0;
  while (p + 1 < P.nprob && item >= P.item_begin[p + 1]) ++p;
  const int t = item - P.item_begin[p];
  Item it;
  it.p = p; it.z = 0;
  it.m0 = (t / P.n_tiles[p]) * (2 * BM);
  it.n0 = (t % P.n_tiles[p]) * BN;
  it.nkb = P.k_blocks[p];
  return it;
}

// raw fp32 tile -> TF32 (hi, lo) pair, 128 threads.  SPLIT 1: hi = rna(x) written in place, lo = rna(x - hi);
// SPLIT 2: the raw tile stays (the MMA reads its top 19 bits = truncation), lo = rna(x - trunc(x)).
template <int SPLIT, int BYTES>
__device__ __forceinline__ void split_tile(uint8_t* hi_raw, uint8_t* lo_out, int t) {
  const uint32_t hi = smem_u32(hi_raw), lo = smem_u32(lo_out);
#pragma unroll
  for (int i = 0; i < BYTES / 16 / 128; ++i) {
    const uint32_t c = (uint32_t)(t + i * 128) * 16;
    const float4 v = lds128(hi + c);
    float4 h, l;
    if (SPLIT == 2) {
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    } else {
      h.x = __uint_as_float(to_tf32(v.x)); h.y = __uint_as_float(to_tf32(v.y));
      h.z = __uint_as_float(to_tf32(v.z)); h.w = __uint_as_float(to_tf32(v.w));
      sts128(hi + c, h);
    }
    l.x = __uint_as_float(to_tf32(v.x - h.x)); l.y = __uint_as_float(to_tf32(v.y - h.y));
    l.z = __uint_as_float(to_tf32(v.z - h.z)); l.w = __uint_as_float(to_tf32(v.w - h.w));
    sts128(lo + c, l);
  }
}

template <bool TIMING>
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc) {
  if constexpr (TIMING) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}
__device__ __forceinline__ long long* timing_row(const Params& P) {
  return P.timing + ((size_t)(P.tn ? TIMING_CTAS : 0) + blockIdx.x) * TIMING_SLOTS;
}

template <int SPLIT, bool TIMING = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc2_gemm_nt_kernel(const __grid_constant__ Maps maps, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  // identical carve-up in both CTAs: the pair addresses its peer's tiles and barriers by the same offsets
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_raw = bars;                       // [STAGES] local: TMA -> splitters
  uint64_t* full_split = bars + STAGES;            // [STAGES] used in the LEADER: splitter warps of both CTAs -> MMA
  uint64_t* empty = bars + 2 * STAGES;             // [STAGES] local: MMA commit (multicast) -> TMA
  uint64_t* acc_full = bars + 3 * STAGES;          // [2] local: MMA commit (multicast) -> epilogue
  uint64_t* acc_empty = bars + 3 * STAGES + 2;     // [2] used in the LEADER: epilogue warps of both CTAs -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  static_assert((3 * STAGES + 4) * 8 + 4 <= BAR_BYTES, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();         // 0 = leader (issues the MMAs), 1 = peer
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_items = P.tn ? P.m_pairs[0] * P.n_tiles[0] * P.splits : P.item_begin[P.nprob];
  long long t_kernel0 = 0;
  if constexpr (TIMING) t_kernel0 = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&full_split[s], 4 * 2);            // one arrival per splitter warp of each CTA
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], EPI_WARPS * 2);     // one arrival per epilogue warp of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // the same logical warp in both CTAs allocates (and later frees) the pair's tensor memory
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // the peer's barriers exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0, t_begin = 0;
      if constexpr (TIMING) t_begin = clock64();
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const Item w = decode_item(P, item);
        const int m0 = w.m0 + (int)rank * BM;      // this CTA's accumulator rows
        const int nb0 = w.n0 + (int)rank * BNH;    // this CTA's half of the weight rows
        for (int kb = 0; kb < w.nkb; ++kb) {
          mbar_wait_t<TIMING>(&empty[stage], phase ^ 1, t_wait);
          uint8_t* st = smem + stage * STAGE_BYTES;
          if (!P.tn) {
            mbar_arrive_expect_tx(&full_raw[stage], A_BYTES + 2 * BH_BYTES);
            tma_load_2d(&maps.a[w.p], &full_raw[stage], st, kb * BKF, m0);
            tma_load_2d(&maps.b_hi[w.p], &full_raw[stage], st + 2 * A_BYTES, kb * BKF, nb0);
            tma_load_2d(&maps.b_lo[w.p], &full_raw[stage], st + 2 * A_BYTES + BH_BYTES, kb * BKF, nb0);
          } else {
            mbar_arrive_expect_tx(&full_raw[stage], A_BYTES + BH_BYTES);
            const int row = w.z * P.chunk_rows + kb * BKF;      // 32 reduction rows per stage
#pragma unroll
            for (int j = 0; j < 4; ++j)                         // four 32-float column groups of G
              tma_load_2d(&maps.a[0], &full_raw[stage], st + j * 4096, m0 + 32 * j, row);
#pragma unroll
            for (int j = 0; j < 2; ++j)                         // two column groups of X (this CTA's half)
              tma_load_2d(&maps.b_hi[0], &full_raw[stage], st + 2 * A_BYTES + j * 4096, nb0 + 32 * j, row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (TIMING) {
        long long* T = timing_row(P);
        T[TS_TMA_WAIT_EMPTY] += t_wait;
        T[TS_TMA_TOTAL] += clock64() - t_begin;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long t_wait_split = 0, t_wait_acc = 0, t_begin = 0, n_kb = 0;
      if constexpr (TIMING) t_begin = clock64();
      for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
        const int nkb = decode_item(P, item).nkb;
        if constexpr (TIMING) n_kb += nkb;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_t<TIMING>(&acc_empty[acc], acc_phase ^ 1, t_wait_acc);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 2 * BN;   // sum of hi*hi   (same columns in both CTAs' TMEM)
        const uint32_t tmem_x = tmem_d + BN;                // sum of lo*hi + hi*lo
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_t<TIMING>(&full_split[stage], phase, t_wait_split);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + stage * STAGE_BYTES);
          uint64_t d_ahi, d_alo, d_bhi, d_blo, kstep;
          if (!P.tn) {
            d_ahi = make_desc(base); d_alo = make_desc(base + A_BYTES);
            d_bhi = make_desc(base + 2 * A_BYTES); d_blo = make_desc(base + 2 * A_BYTES + BH_BYTES);
            kstep = 32 >> 4;       // 8 tf32 = 32 B along K, in 16-byte units
          } else {
            d_ahi = make_desc_mn(base); d_alo = make_desc_mn(base + A_BYTES);
            d_bhi = make_desc_mn(base + 2 * A_BYTES); d_blo = make_desc_mn(base + 2 * A_BYTES + BH_BYTES);
            kstep = 1024 >> 4;     // 8 reduction rows = one 1024-byte swizzle atom
          }
          const uint32_t idesc = P.tn ? IDESC2_TN : IDESC2;
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k)
            umma2_tf32(tmem_d, d_ahi + k * kstep, d_bhi + k * kstep, idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k) {
            umma2_tf32(tmem_x, d_alo + k * kstep, d_bhi + k * kstep, idesc, (kb | k) != 0);
            umma2_tf32(tmem_x, d_ahi + k * kstep, d_blo + k * kstep, idesc, 1);
          }
          umma2_commit_both(&empty[stage]);                      // both CTAs may refill this stage
          if (kb == nkb - 1) umma2_commit_both(&acc_full[acc]);  // both epilogues may drain
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (TIMING) {
        long long* T = timing_row(P);
        T[TS_MMA_WAIT_SPLIT] += t_wait_split;
        T[TS_MMA_WAIT_ACC] += t_wait_acc;
        T[TS_MMA_TOTAL] += clock64() - t_begin;
        T[TS_ITEMS] += it;
        T[TS_KBLOCKS] += n_kb;
      }
    }
  } else if (warp < 6) {
    // ================= splitters (both CTAs) =================
    const int t = threadIdx.x - 64;   // 0..127
    int stage = 0;
    uint32_t phase = 0;
    long long t_wait = 0, t_work = 0, t_begin = 0;
    if constexpr (TIMING) t_begin = clock64();
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const int nkb = decode_item(P, item).nkb;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_t<TIMING>(&full_raw[stage], phase, t_wait);
        long long t_w0 = 0;
        if constexpr (TIMING) t_w0 = clock64();
        uint8_t* st = smem + stage * STAGE_BYTES;
        if (!(P.debug & 1)) {
          split_tile<SPLIT, A_BYTES>(st, st + A_BYTES, t);
          if (P.tn) split_tile<SPLIT, BH_BYTES>(st + 2 * A_BYTES, st + 2 * A_BYTES + BH_BYTES, t);   // X half is raw too
        }
        fence_proxy_async();            // this thread's generic-proxy writes -> visible to the tensor-core proxy
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&full_split[stage], 0);
        if constexpr (TIMING) t_work += clock64() - t_w0;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if constexpr (TIMING) {
      if (t == 0) {
        long long* T = timing_row(P);
        T[TS_SPL_WAIT_RAW] += t_wait;
        T[TS_SPL_WORK] += t_work;
        T[TS_SPL_TOTAL] += clock64() - t_begin;
      }
    }
  } else {
    // ================= epilogue (both CTAs; same data path as gemm_tc.cu) =================
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const uint32_t stg_s = smem_u32(smem + STAGES * STAGE_BYTES + BAR_BYTES) + (uint32_t)(warp - 6) * (32 * EPI_LD * 4);
    const int rr = lane >> 2, cc = (lane & 3) * 4;
    int it = 0;
    long long t_wait = 0, t_work = 0, t_begin = 0;
    if constexpr (TIMING) t_begin = clock64();
    for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
      const Item w = decode_item(P, item);
      const GemmNT& g = P.g[w.p];
      const bool vec_c = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
      const bool vec_x = g.aux && (g.ldaux & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.aux) & 15) == 0);
      float* const Cbase = g.C + (P.tn ? (size_t)w.z * P.tn_nn * P.tn_kk : (size_t)0);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = w.m0 + (int)rank * BM, n0 = w.n0;
      mbar_wait_t<TIMING>(&acc_full[acc], acc_phase, t_wait);
      long long t_w0 = 0;
      if constexpr (TIMING) t_w0 = clock64();
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        const int col0 = half * 64 + chunk * 16;
        uint32_t r[16], rx[16];
        __syncwarp();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN + col0, r);
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 2 * BN + BN + col0, rx);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          sts128(stg_s + (uint32_t)(lane * EPI_LD + j4 * 4) * 4,
                 make_float4(__uint_as_float(r[j4 * 4 + 0]) + __uint_as_float(rx[j4 * 4 + 0]),
                             __uint_as_float(r[j4 * 4 + 1]) + __uint_as_float(rx[j4 * 4 + 1]),
                             __uint_as_float(r[j4 * 4 + 2]) + __uint_as_float(rx[j4 * 4 + 2]),
                             __uint_as_float(r[j4 * 4 + 3]) + __uint_as_float(rx[j4 * 4 + 3])));
        __syncwarp();
        const int n = n0 + col0 + cc;
        if (n < g.n_store && !(P.debug & 2)) {
          float bj[4] = {0.f, 0.f, 0.f, 0.f};
          if (g.mode == EPI_ACT && g.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < g.N) bj[j] = __ldg(g.bias + n + j);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + rr;
            const int m = m0 + q * 32 + row;
            if (m >= g.M) continue;
            const float4 a4 = lds128(stg_s + (uint32_t)(row * EPI_LD + cc) * 4);
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
            float x[4] = {0.f, 0.f, 0.f, 0.f};
            if (g.mode != EPI_ACT) {
              const float* ax = g.aux + (size_t)m * g.ldaux + n;
              if (vec_x && n + 3 < g.n_store) {
                const float4 t4 = *reinterpret_cast<const float4*>(ax);
                x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w;
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (n + j < g.n_store) x[j] = ax[j];
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (g.mode == EPI_ACT) v[j] = act_fast(v[j] + bj[j], g.act);
              else if (g.mode == EPI_MUL_DACT) v[j] = v[j] * dact_from_out(x[j], g.act);
              else v[j] = v[j] + x[j];
              if (n + j >= g.n_valid) v[j] = 0.f;
            }
            float* dst = Cbase + (size_t)m * g.ldc + n;
            if (vec_c && n + 3 < g.n_store) {
              *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n + j < g.n_store) dst[j] = v[j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
      if constexpr (TIMING) t_work += clock64() - t_w0;
    }
    if constexpr (TIMING) {
      if (warp == 6 && lane == 0) {
        long long* T = timing_row(P);
        T[TS_EPI_WAIT_ACC] += t_wait;
        T[TS_EPI_WORK] += t_work;
        T[TS_EPI_TOTAL] += clock64() - t_begin;
      }
    }
  }

  // nobody leaves while the peer may still signal this CTA's barriers or the pair's MMAs read its shared memory
  tc_fence_before();
  __syncthreads();
  if constexpr (TIMING) {
    if (threadIdx.x == 0) {
      long long* T = timing_row(P);
      T[TS_KERNEL_TOTAL] += clock64() - t_kernel0;
      T[TS_LAUNCHES] += 1;
    }
  }
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

}  // namespace tc2

bool tc2_eligible(const GemmNT& p) {
  return tc_eligible(p) && p.B_hi && p.B_lo && (reinterpret_cast<uintptr_t>(p.B_hi) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.B_lo) & 15) == 0;
}

// cluster launch of CTA pairs; the number of co-resident pairs is queried once (pairs must sit on one TPC, so it can
// be below SMs / 2)
static int pair_launch(const tc2::Maps& maps, const tc2::Params& P, int items, cudaStream_t st) {
  using namespace tc2;
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters == 0) {
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc2_gemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc2_gemm_nt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc2_gemm_nt_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc2_gemm_nt_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int dev = 0, sms = 0;
    GIB_CUDA_TRY(cudaGetDevice(&dev));
    GIB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cfg.gridDim = dim3(sms & ~1, 1, 1);
    int nc = 0;
    GIB_CUDA_TRY(cudaOccupancyMaxActiveClusters(&nc, tc2_gemm_nt_kernel<1>, &cfg));
    if (nc < 1) { set_error("gemm_tc2: no CTA pair fits on this device"); return -4; }
    max_clusters = nc < sms / 2 ? nc : sms / 2;
  }
  const int clusters = items < max_clusters ? items : max_clusters;
  cfg.gridDim = dim3(2 * clusters, 1, 1);
  tc2::Params Q = P;
  Q.timing = g_tc_timing;
  const bool raw_hi = (g_tc_debug & 64) != 0;
  if (Q.timing) {
    if (raw_hi) GIB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc2_gemm_nt_kernel<2, true>, maps, Q));
    else GIB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc2_gemm_nt_kernel<1, true>, maps, Q));
  } else if (raw_hi) {
    GIB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc2_gemm_nt_kernel<2>, maps, Q));
  } else {
    GIB_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc2_gemm_nt_kernel<1>, maps, Q));
  }
  ++g_launch_count;
  return 0;
}
static int pair_capacity() { return 74; }   // planning figure for split counts (148 SMs); the launch clamps to the real one

// up to MAXP independent NT problems (pre-split weights) in one persistent launch of CTA pairs
int gemm_nt_tc2_group(const GemmNT* ps, int n, cudaStream_t st) {
  using namespace tc2;
  if (n < 1 || n > MAXP) { set_error("gemm_nt_tc2_group: %d problems (max %d)", n, MAXP); return -2; }
  Maps maps;
  Params P;
  memset(&P, 0, sizeof(P));
  double work = 0;
  int items = 0, np = 0;
  for (int i = 0; i < n; ++i) {
    const GemmNT& p = ps[i];
    if (p.M <= 0 || p.N <= 0) continue;
    if (!tc2_eligible(p)) { set_error("gemm_nt_tc2: needs TMA-aligned operands and pre-split weight planes"); return -2; }
    GIB_TRY(tc_make_map(&maps.a[np], p.A, p.M, p.K, p.lda, BM, 0));
    GIB_TRY(tc_make_map(&maps.b_hi[np], p.B_hi, p.N, p.K, p.ldb, BNH, 0));
    GIB_TRY(tc_make_map(&maps.b_lo[np], p.B_lo, p.N, p.K, p.ldb, BNH, 0));
    P.g[np] = p;
    P.m_pairs[np] = ceil_div(p.M, 2 * BM);
    P.n_tiles[np] = ceil_div(p.N, BN);
    P.k_blocks[np] = ceil_div(p.K, BKF);
    P.item_begin[np] = items;
    items += P.m_pairs[np] * P.n_tiles[np];
    work += p.work > 0 ? p.work : 2.0 * p.M * (double)p.N * p.K;
    ++np;
  }
  if (np == 0) return 0;
  for (int i = np; i <= MAXP; ++i) P.item_begin[i] = items;
  P.nprob = np;
  P.debug = g_tc_debug;
  ProfScope prof(PROF_GEMM_NT, work, st);
  return pair_launch(maps, P, items, st);
}

// weight-gradient partial products into q.scratch ([splits][Nn][Kk]); the caller reduces them (reduce_grads_kernel).
// Never more splits than tc_dw_plan() chose: the scratch halves are sized for that plan.
int gemm_dw_tc2_partials(const GemmDW& q, int* splits_out, cudaStream_t st) {
  using namespace tc2;
  int plan_splits, plan_chunk;
  tc_dw_plan(q.M, q.Nn, q.Kk, &plan_splits, &plan_chunk);
  const int tiles = ceil_div(q.Nn, 2 * BM) * ceil_div(q.Kk, BN);
  int s = ceil_div(pair_capacity(), tiles);
  if (s > plan_splits) s = plan_splits;
  if (s < 1) s = 1;
  int chunk = ceil_div(ceil_div(q.M, s), BKF) * BKF;
  if (chunk < 8 * BKF) chunk = 8 * BKF;
  if (chunk < plan_chunk) chunk = plan_chunk;
  const int splits = ceil_div(q.M, chunk);
  Maps maps;
  GIB_TRY(tc_make_map(&maps.a[0], q.G, q.M, q.Nn, q.ldg, BKF, 1));     // 32 reduction rows x 32 floats
  GIB_TRY(tc_make_map(&maps.b_hi[0], q.X, q.M, q.Kk, q.ldx, BKF, 1));
  Params P;
  memset(&P, 0, sizeof(P));
  GemmNT& g = P.g[0];
  g = GemmNT();
  g.C = q.scratch; g.ldc = q.Kk; g.M = q.Nn; g.N = q.Kk; g.n_store = q.Kk; g.n_valid = q.Kk;
  g.mode = EPI_ACT; g.act = ACT_NONE; g.bias = nullptr;
  P.m_pairs[0] = ceil_div(q.Nn, 2 * BM);
  P.n_tiles[0] = ceil_div(q.Kk, BN);
  P.nprob = 1;
  P.tn = 1; P.splits = splits; P.chunk_rows = chunk; P.tn_rows = q.M; P.tn_nn = q.Nn; P.tn_kk = q.Kk;
  P.debug = g_tc_debug;
  GIB_TRY(pair_launch(maps, P, P.m_pairs[0] * P.n_tiles[0] * splits, st));
  *splits_out = splits;
  return 0;
}

}  // namespace gib
