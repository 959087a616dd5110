// tcgen05 (5th-gen tensor core) GEMM with fp32-accurate 3xTF32 operand splitting, sm_100a.
//
//   C[M, :N] = epi( A[M,K] * B[N,K]^T ),   A, B fp32 row-major, K-contiguous (same contract as gemm_nt)
//
// Why 3xTF32: the parity bar is 1e-4 on the logits; one TF32 pass is 1.8e-2 off.  Each
// operand x is split into hi = tf32(x) and lo = tf32(x - hi); the product is accumulated as
// hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator (the dropped lo*lo term is ~2^-22 relative).
//
// Structure (one persistent CTA per SM, warp-specialised, 3-stage smem ring, 2 TMEM accumulator stages):
//   warp 0      TMA producer   : cp.async.bulk.tensor (128B swizzle) raw fp32 A/B k-blocks -> smem
//   warps 2-5   splitters      : in-place raw -> hi, plus lo into a sibling tile (same swizzled address,
//                                 so the swizzle pattern never has to be decoded), fence.proxy.async
//   warp 1      MMA issuer     : 12 tcgen05.mma.kind::tf32 per k-block (4 k-steps x 3 products), accumulators
//                                 in TMEM; tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 6-13  epilogue       : tcgen05.ld TMEM -> registers -> smem transpose -> bias / SELU / dSELU / add ->
//                                 coalesced global stores (two warps per TMEM lane quarter)
// Every mbarrier wait is bounded: a dead-lock turns into a trap (error at the next API call), never a hang.
#include <cuda.h>
#include <string.h>

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace gib {

namespace tc {

using namespace tcptx;

constexpr int BM = 128, BN = 128, BKF = 32;       // tile: 128 x 128 outputs, 32 floats (128 B) of K per stage
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BKF * 4;          // 16 KB per operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi | A_lo | B_hi | B_lo
constexpr int EPI_WARPS = 8;                       // two warps per TMEM lane quarter, each owning 64 of the 128 columns
constexpr int EPI_LD = 20;                         // padded row (floats) of the per-warp 32x16 epilogue staging tile
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_WARPS * 32 * EPI_LD * 4;
constexpr int NUM_THREADS = 192 + 32 * 8;          // TMA + MMA + 4 splitter + 8 epilogue warps
constexpr int TMEM_COLS = 512;                    // 2 stages x (main hi*hi + cross-term accumulator) x 128 fp32 columns

// mbar_wait that also accumulates the cycles spent waiting (TIMING instantiation)
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

// instruction descriptor, kind::tf32: D = F32 (bits 4-5 = 1), A/B = TF32 (bits 7-9, 10-12 = 2), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.  (InstrDescriptor in the same header)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_TN = IDESC | (1u << 15) | (1u << 16);   // A and B MN-major (bits 15, 16)

constexpr int MAXP = 4;   // independent GEMM problems per launch (grouped NT launches)

struct Maps {   // TMA descriptors live in kernel-parameter space (__grid_constant__)
  CUtensorMap a[MAXP];
  CUtensorMap b[MAXP];      // raw fp32 B, or its TF32 hi plane when the caller supplies pre-split weights
  CUtensorMap b_lo[MAXP];   // lo plane (pre-split weights only)
};

struct Params {
  // NT mode: up to MAXP independent problems C_p = epi(A_p B_p^T) share one persistent launch; work items are their
  // output tiles, concatenated (tile_begin[p] .. tile_begin[p+1]).  Used for the per-bond-type message MLPs (same
  // layer, different weights and row ranges) and for sibling MLPs of the readout.
  GemmNT g[MAXP];
  int m_tiles[MAXP], n_tiles[MAXP], k_blocks[MAXP], tile_begin[MAXP + 1];
  int bsplit[MAXP];   // 1: B arrives as (hi, lo) planes from the packed arena -> only A is split in the kernel
  int nprob;
  // TN (weight-gradient) mode: P[z][n][k] = sum_{m in chunk z} G[m,n] X[m,k]; operands are the row-major
  // activations themselves (MN-major for the MMA), problem 0 only, g[0].C = split-K workspace
  int tn, splits, chunk_rows, tn_rows, tn_nn, tn_kk;
  int debug;   // bit0: skip the hi/lo split (timing experiments only), bit1: skip the epilogue stores
  // TIMING instantiation only: [2 (NT, TN)][TIMING_CTAS][TIMING_SLOTS] clock64 totals per role, ACCUMULATED over
  // launches (each CTA index owns its row; tcgen05 launches are stream-ordered)
  long long* timing;
};


__device__ __forceinline__ long long* timing_row(const Params& P) {
  return P.timing + ((size_t)(P.tn ? TIMING_CTAS : 0) + blockIdx.x) * TIMING_SLOTS;
}

struct Item { int p, m0, n0, nkb, z; };

__device__ __forceinline__ Item decode_item(const Params& P, int item) {
  Item it;
  if (P.tn) {
    const int tiles_mn = P.m_tiles[0] * P.n_tiles[0];
    const int tile = item % tiles_mn;
    it.p = 0; it.z = item / tiles_mn;
    it.m0 = (tile / P.n_tiles[0]) * BM; it.n0 = (tile % P.n_tiles[0]) * BN;
    const int r0 = it.z * P.chunk_rows;
    it.nkb = ceil_div(min(P.tn_rows, r0 + P.chunk_rows) - r0, BKF);
  } else {
    int p = 0;
    while (p + 1 < P.nprob && item >= P.tile_begin[p + 1]) ++p;
    const int tile = item - P.tile_begin[p];
    it.p = p; it.z = 0;
    it.m0 = (tile / P.n_tiles[p]) * BM; it.n0 = (tile % P.n_tiles[p]) * BN;
    it.nkb = P.k_blocks[p];
  }
  return it;
}

// SPLIT 0: hi = truncation, lo = exact remainder;  1: hi, lo both round-to-nearest (cvt.rna);
//       2: hi = the raw value (the MMA reads its top 19 bits = truncation, nothing is written back), lo = rna(x - trunc(x))
//       3: SPLIT 1 with explicit shared-window loads / stores in the splitters and the epilogue staging
// TIMING: accumulate per-role wait / work cycles into P.timing (diagnosis builds; the product uses <1, false>)
template <int SPLIT, bool TIMING = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_nt_kernel(const __grid_constant__ Maps maps, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_raw = bars;                 // [STAGES] TMA -> splitters
  uint64_t* full_split = bars + STAGES;      // [STAGES] splitters -> MMA
  uint64_t* empty = bars + 2 * STAGES;       // [STAGES] MMA -> TMA
  uint64_t* acc_full = bars + 3 * STAGES;    // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 3 * STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t_kernel0 = 0;
  if constexpr (TIMING) t_kernel0 = clock64();
  // work items: NT = output tiles of all problems; TN = (output tile, reduction chunk) pairs
  const int num_tiles = P.tn ? P.m_tiles[0] * P.n_tiles[0] * P.splits : P.tile_begin[P.nprob];

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&full_split[s], 128);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 32 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0, t_begin = 0;
      if constexpr (TIMING) t_begin = clock64();
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        const Item w = decode_item(P, item);
        const CUtensorMap* map_a = &maps.a[w.p];
        const CUtensorMap* map_b = &maps.b[w.p];
        const int m0 = w.m0, n0 = w.n0, nkb = w.nkb;
        const int r0 = w.z * P.chunk_rows;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_t<TIMING>(&empty[stage], phase ^ 1, t_wait);
          uint8_t* st = smem + stage * STAGE_BYTES;
          const bool presplit = !P.tn && P.bsplit[w.p];
          mbar_arrive_expect_tx(&full_raw[stage], (presplit ? 3 : 2) * TILE_BYTES);
          if (!P.tn) {
            tma_load_2d(map_a, &full_raw[stage], st, kb * BKF, m0);
            tma_load_2d(map_b, &full_raw[stage], st + 2 * TILE_BYTES, kb * BKF, n0);
            if (presplit) tma_load_2d(&maps.b_lo[w.p], &full_raw[stage], st + 3 * TILE_BYTES, kb * BKF, n0);
          } else {
            const int row = r0 + kb * BKF;           // 32 reduction rows per stage
#pragma unroll
            for (int j = 0; j < 4; ++j) {            // four 32-float column groups per operand
              tma_load_2d(map_a, &full_raw[stage], st + j * 4096, m0 + 32 * j, row);
              tma_load_2d(map_b, &full_raw[stage], st + 2 * TILE_BYTES + j * 4096, n0 + 32 * j, row);
            }
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
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long t_wait_split = 0, t_wait_acc = 0, t_begin = 0, n_kb = 0;
      if constexpr (TIMING) t_begin = clock64();
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x, ++it) {
        const int nkb = decode_item(P, item).nkb;
        if constexpr (TIMING) n_kb += nkb;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_t<TIMING>(&acc_empty[acc], acc_phase ^ 1, t_wait_acc);
        tc_fence_after();
        // Two accumulators per tile.  The tensor core adds into TMEM with truncation, one rounding per MMA; the
        // hi*lo / lo*hi products are 2^-11 of the hi*hi ones, so giving them their own accumulator keeps their
        // 2/3 of the roundings away from the large running sum (measured: error / 3, back to fp32-FMA level).
        const uint32_t tmem_d = tmem_base + acc * 2 * BN;        // sum of hi*hi
        const uint32_t tmem_x = tmem_d + BN;                     // sum of hi*lo + lo*hi
        const uint32_t idesc = P.tn ? IDESC_TN : IDESC;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_t<TIMING>(&full_split[stage], phase, t_wait_split);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * STAGE_BYTES);
          uint64_t d_ahi, d_alo, d_bhi, d_blo, kstep;
          if (!P.tn) {
            d_ahi = make_desc(a_hi); d_alo = make_desc(a_hi + TILE_BYTES);
            d_bhi = make_desc(a_hi + 2 * TILE_BYTES); d_blo = make_desc(a_hi + 3 * TILE_BYTES);
            kstep = 32 >> 4;       // 8 tf32 = 32 B further along K, in 16 B units
          } else {
            d_ahi = make_desc_mn(a_hi); d_alo = make_desc_mn(a_hi + TILE_BYTES);
            d_bhi = make_desc_mn(a_hi + 2 * TILE_BYTES); d_blo = make_desc_mn(a_hi + 3 * TILE_BYTES);
            kstep = 1024 >> 4;     // 8 reduction rows = one 1024 B swizzle atom further
          }
          // grouped by accumulator so that consecutive MMAs chain on the same TMEM tile
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k)
            umma_tf32(tmem_d, d_ahi + k * kstep, d_bhi + k * kstep, idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k) {
            umma_tf32(tmem_x, d_alo + k * kstep, d_bhi + k * kstep, idesc, (kb | k) != 0);
            umma_tf32(tmem_x, d_ahi + k * kstep, d_blo + k * kstep, idesc, 1);
          }
          umma_commit(&empty[stage]);                        // frees the smem stage when the MMAs retire
          if (kb == nkb - 1) umma_commit(&acc_full[acc]);    // accumulators complete
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
    // ================= splitters: raw fp32 -> (hi, lo) TF32 pairs, in place =================
    const int t = threadIdx.x - 64;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    long long t_wait = 0, t_work = 0, t_begin = 0;
    if constexpr (TIMING) t_begin = clock64();
    for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
      const Item wi = decode_item(P, item);
      const int nkb = wi.nkb;
      const int nops = (!P.tn && P.bsplit[wi.p]) ? 1 : 2;   // pre-split weights: only the A tile needs splitting
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_t<TIMING>(&full_raw[stage], phase, t_wait);
        long long t_w0 = 0;
        if constexpr (TIMING) t_w0 = clock64();
        uint8_t* st = smem + stage * STAGE_BYTES;
        if (SPLIT == 3) {
          // same arithmetic as SPLIT 1 through explicit LDS.128 / STS.128 (the pointer form below compiles to
          // generic LD.E / ST.E): isolates the cost of the generic address path
          for (int op = 0; op < nops; ++op) {
            const uint32_t hi_s = smem_u32(st + op * 2 * TILE_BYTES), lo_s = hi_s + TILE_BYTES;
#pragma unroll
            for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
              const uint32_t c = (uint32_t)(t + i * 128) * 16;
              const float4 v = lds128(hi_s + c);
              float4 h, l;
              h.x = __uint_as_float(to_tf32(v.x)); l.x = __uint_as_float(to_tf32(v.x - h.x));
              h.y = __uint_as_float(to_tf32(v.y)); l.y = __uint_as_float(to_tf32(v.y - h.y));
              h.z = __uint_as_float(to_tf32(v.z)); l.z = __uint_as_float(to_tf32(v.z - h.z));
              h.w = __uint_as_float(to_tf32(v.w)); l.w = __uint_as_float(to_tf32(v.w - h.w));
              sts128(hi_s + c, h);
              sts128(lo_s + c, l);
            }
          }
        } else if (!(P.debug & 1)) {
          auto split_tile = [&](int op) {
            float4* hi = reinterpret_cast<float4*>(st + op * 2 * TILE_BYTES);
            float4* lo = reinterpret_cast<float4*>(st + op * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
            for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
              const int c = t + i * 128;
              const float4 v = hi[c];
              float4 h, l;
              if (SPLIT == 2) {    // hi stays the raw tile (not written back); only the rounded remainder is stored
                l.x = __uint_as_float(to_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)));
                l.y = __uint_as_float(to_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)));
                l.z = __uint_as_float(to_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)));
                l.w = __uint_as_float(to_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)));
              } else if (SPLIT == 1) {    // round-to-nearest split: 3 conversions per element, smallest error
                h.x = __uint_as_float(to_tf32(v.x)); l.x = __uint_as_float(to_tf32(v.x - h.x));
                h.y = __uint_as_float(to_tf32(v.y)); l.y = __uint_as_float(to_tf32(v.y - h.y));
                h.z = __uint_as_float(to_tf32(v.z)); l.z = __uint_as_float(to_tf32(v.z - h.z));
                h.w = __uint_as_float(to_tf32(v.w)); l.w = __uint_as_float(to_tf32(v.w - h.w));
              } else {             // hi = top 19 bits (exact tf32), lo = x - hi (exact in fp32; the MMA reads its top 19 bits)
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
              }
              if (SPLIT != 2) hi[c] = h;
              lo[c] = l;
            }
          };
          split_tile(0);
          if (nops == 2) split_tile(1);
        }
        fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        mbar_arrive(&full_split[stage]);
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
    // ================= epilogue: TMEM -> registers -> smem transpose -> coalesced global =================
    // tcgen05.ld hands each lane one accumulator ROW; storing that directly would scatter 16-byte pieces over 32
    // rows per instruction (measured: 0.6 TB/s).  Each warp therefore bounces 32x32 chunks through a private
    // padded smem tile and writes 4 rows x 128 contiguous bytes per instruction; bias / activation / aux math runs
    // in that coalesced domain, so the aux operand (dSELU source or residual) is read coalesced as well.
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - 6) >> 2;          // which 64-column half of the tile this warp drains
    float* stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256) + (warp - 6) * (32 * EPI_LD);
    const int rr = lane >> 2, cc = (lane & 3) * 4;
    int it = 0;
    long long t_wait = 0, t_work = 0, t_begin = 0;
    if constexpr (TIMING) t_begin = clock64();
    for (int item = blockIdx.x; item < num_tiles; item += gridDim.x, ++it) {
      const Item w = decode_item(P, item);
      const GemmNT& g = P.g[w.p];
      const bool vec_c = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
      const bool vec_x = g.aux && (g.ldaux & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.aux) & 15) == 0);
      float* const Cbase = g.C + (P.tn ? (size_t)w.z * P.tn_nn * P.tn_kk : (size_t)0);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = w.m0, n0 = w.n0;
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
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 sum4 = make_float4(__uint_as_float(r[j4 * 4 + 0]) + __uint_as_float(rx[j4 * 4 + 0]),
                                          __uint_as_float(r[j4 * 4 + 1]) + __uint_as_float(rx[j4 * 4 + 1]),
                                          __uint_as_float(r[j4 * 4 + 2]) + __uint_as_float(rx[j4 * 4 + 2]),
                                          __uint_as_float(r[j4 * 4 + 3]) + __uint_as_float(rx[j4 * 4 + 3]));
          if constexpr (SPLIT == 3) sts128(smem_u32(stg + lane * EPI_LD + j4 * 4), sum4);
          else *reinterpret_cast<float4*>(stg + lane * EPI_LD + j4 * 4) = sum4;
        }
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
            float4 a4;
            if constexpr (SPLIT == 3) a4 = lds128(smem_u32(stg + row * EPI_LD + cc));
            else a4 = *reinterpret_cast<const float4*>(stg + row * EPI_LD + cc);
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
      mbar_arrive(&acc_empty[acc]);
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

  tc_fence_before();
  __syncthreads();
  if constexpr (TIMING) {
    if (threadIdx.x == 0) {
      long long* T = timing_row(P);
      T[TS_KERNEL_TOTAL] += clock64() - t_kernel0;
      T[TS_LAUNCHES] += 1;
    }
  }
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows = BM,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return -4; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BKF, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld); return -4; }
  return 0;
}

}  // namespace tc

int tc_make_map(void* map, const float* base, int rows, int cols, int ld, int box_rows, int mn_major) {
  return tc::make_map(reinterpret_cast<CUtensorMap*>(map), base, rows, cols, ld, box_rows,
                      mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

bool g_use_tc = true;
int g_tc_debug = 0;               // bit 0: route every GEMM to this first-generation kernel instead of gemm_tc3.cu

bool tc_eligible(const GemmNT& p) {
  return p.M >= 1 && p.N >= 1 && p.K >= 16 && (p.K % 16) == 0 && (p.lda % 4) == 0 && (p.ldb % 4) == 0 &&
         (reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.B) & 15) == 0;
}

static int tc_prepare(int* num_sms_out) {
  using namespace tc;
  static int num_sms_dev[64] = {0};
  static bool attr_done_dev[64] = {false};
  int dev = 0;
  GIB_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return -4; }
  int& num_sms = num_sms_dev[dev];
  bool& attr_done = attr_done_dev[dev];
  if (!attr_done) {   // function attributes are per device
    GIB_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    GIB_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  *num_sms_out = num_sms;
  return 0;
}

// product instantiation: <1> = round-to-nearest split (the other SPLIT / TIMING template branches are not instantiated)
static void launch_tc(int grid, const tc::Maps& maps, tc::Params& P, cudaStream_t st) {
  using namespace tc;
  P.timing = nullptr;
  tc_gemm_nt_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, P);
}

// up to MAXP independent NT problems in one persistent launch
int gemm_nt_tc_group(const GemmNT* ps, int n, cudaStream_t st) {
  using namespace tc;
  if (n < 1 || n > MAXP) { set_error("gemm_nt_tc_group: %d problems (max %d)", n, MAXP); return -2; }
  int num_sms = 0;
  GIB_TRY(tc_prepare(&num_sms));
  Maps maps;
  Params P;
  memset(&P, 0, sizeof(P));
  double work = 0;
  int tiles = 0, np = 0;
  for (int i = 0; i < n; ++i) {
    const GemmNT& p = ps[i];
    if (p.M <= 0 || p.N <= 0) continue;
    if (!tc_eligible(p)) { set_error("gemm_nt_tc: operands violate the TMA alignment contract"); return -2; }
    GIB_TRY(make_map(&maps.a[np], p.A, p.M, p.K, p.lda));
    const bool presplit = p.B_hi && p.B_lo && (reinterpret_cast<uintptr_t>(p.B_hi) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(p.B_lo) & 15) == 0;
    GIB_TRY(make_map(&maps.b[np], presplit ? p.B_hi : p.B, p.N, p.K, p.ldb));
    if (presplit) GIB_TRY(make_map(&maps.b_lo[np], p.B_lo, p.N, p.K, p.ldb));
    P.bsplit[np] = presplit ? 1 : 0;
    P.g[np] = p;
    P.m_tiles[np] = ceil_div(p.M, BM);
    P.n_tiles[np] = ceil_div(p.N, BN);
    P.k_blocks[np] = ceil_div(p.K, BKF);
    P.tile_begin[np] = tiles;
    tiles += P.m_tiles[np] * P.n_tiles[np];
    work += p.work > 0 ? p.work : 2.0 * p.M * (double)p.N * p.K;
    ++np;
  }
  if (np == 0) return 0;
  for (int i = np; i <= MAXP; ++i) P.tile_begin[i] = tiles;
  P.nprob = np;
  P.debug = 0;
  const int grid = tiles < num_sms ? tiles : num_sms;
  ProfScope prof(PROF_GEMM_NT, work, st);
  launch_tc(grid, maps, P, st);
  GIB_LAUNCH_CHECK();
  return 0;
}

int gemm_nt_tc(const GemmNT& p, cudaStream_t st) { return gemm_nt_tc_group(&p, 1, st); }

// ---- weight-gradient GEMM on the tensor cores --------------------------------------------------------------
void tc_dw_plan(int M, int Nn, int Kk, int* splits, int* chunk) {
  const int tiles = ceil_div(Nn, tc::BM) * ceil_div(Kk, tc::BN);
  int s = ceil_div(device_sm_count(), tiles);         // about one work item per SM
  int c = ceil_div(ceil_div(M, s), tc::BKF) * tc::BKF;
  if (c < 8 * tc::BKF) c = 8 * tc::BKF;               // >= 256 reduction rows per item: amortise the 64 KB tile drain
  s = ceil_div(M, c);
  if (s < 1) s = 1;
  *splits = s;
  *chunk = c;
}

bool tc_dw_eligible(const GemmDW& q) {
  return q.M >= 2048 && q.Nn >= 32 && q.Kk >= 32 && (q.ldg % 4) == 0 && (q.ldx % 4) == 0 &&
         (reinterpret_cast<uintptr_t>(q.G) & 15) == 0 && (reinterpret_cast<uintptr_t>(q.X) & 15) == 0;
}

// partial products into q.scratch ([splits][Nn][Kk]); the caller reduces them (reduce_grads_kernel)
int gemm_dw_tc_partials(const GemmDW& q, int* splits_out, cudaStream_t st) {
  using namespace tc;
  int num_sms = 0;
  GIB_TRY(tc_prepare(&num_sms));
  int splits, chunk;
  tc_dw_plan(q.M, q.Nn, q.Kk, &splits, &chunk);
  Maps maps;
  GIB_TRY(make_map(&maps.a[0], q.G, q.M, q.Nn, q.ldg, BKF, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));   // 32 floats x 32 rows
  GIB_TRY(make_map(&maps.b[0], q.X, q.M, q.Kk, q.ldx, BKF, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  Params P;
  memset(&P, 0, sizeof(P));
  GemmNT& g = P.g[0];
  g = GemmNT();
  g.C = q.scratch; g.ldc = q.Kk; g.M = q.Nn; g.N = q.Kk; g.n_store = q.Kk; g.n_valid = q.Kk;
  g.mode = EPI_ACT; g.act = ACT_NONE; g.bias = nullptr;
  P.m_tiles[0] = ceil_div(q.Nn, BM);
  P.n_tiles[0] = ceil_div(q.Kk, BN);
  P.nprob = 1;
  P.tn = 1; P.splits = splits; P.chunk_rows = chunk; P.tn_rows = q.M; P.tn_nn = q.Nn; P.tn_kk = q.Kk;
  P.debug = 0;
  const int items = P.m_tiles[0] * P.n_tiles[0] * splits;
  const int grid = items < num_sms ? items : num_sms;
  launch_tc(grid, maps, P, st);
  GIB_LAUNCH_CHECK();
  *splits_out = splits;
  return 0;
}

}  // namespace gib
