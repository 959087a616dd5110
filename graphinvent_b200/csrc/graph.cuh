// Device-side molecular-graph structure produced by K0 (graph_build.cu).
#pragma once
#include "common.cuh"

namespace gib {

// header ints written by the scan kernel and read back by the host (16 ints = 64 B)
enum : int {
  HDR_E = 0,            // number of bond entries (non-zero elements of `edges`)
  HDR_P = 1,            // rows of the type-grouped entry arrays (each group padded to 128)
  HDR_TYPE_COUNT = 2,   // [4] entries per bond type
  HDR_TYPE_BASE = 6,    // [5] first row of each type group (multiples of 128); [G] == P
  HDR_FLAGS = 11,
  HDR_CAPACITY = 12,    // host header only: != 0 -> E / P are capacities, the live header stays on the device ...
  HDR_DEV_LO = 13,      // ... at this address (low / high 32 bits)
  HDR_DEV_HI = 14,
  HDR_INTS = 16
};
enum : int {
  GRAPH_FLAG_MULTITYPE = 1,  // some (b,i,j) carries more than one non-zero bond type
  GRAPH_FLAG_NONBINARY = 2,  // some non-zero bond value differs from 1
  GRAPH_FLAG_OVERFLOW = 4    // capacity mode: the batch holds more bond entries than the capacity (results invalid)
};

struct GraphArrays {
  int* ent_src;    // [P] source slot (b*N + j) of the entry, -1 on pad rows
  int* ent_dst;    // [P] destination slot (b*N + i), -1 on pad rows
  float* ent_w;    // [P] bond value edges[b,i,j,t] (1 for one-hot), 0 on pad rows
  int* dst_ptr;    // [S+1] CSR over destination slots
  int* dst_ent;    // [E]   entry rows, ordered (b, i, j, t)  == reference nonzero() order
  int* src_ptr;    // [S+1] CSR over source slots
  int* src_ent;    // [E]   entry rows, ordered (b, j, i, t)
};

size_t graph_count_ws_ints(int B, int G);
// `edges` is float32 (in_dtype 0) or int8 / uint8 (in_dtype 1: the reference's on-disk format, DataProcesser.py:157-161)
int graph_count(const void* edges, int in_dtype, int B, int N, int Ef, int by_type, int* ws, cudaStream_t st);
// cap_E / cap_P > 0: capacity mode -- arrays hold cap_E entries / cap_P rows; writes beyond are dropped, the tail rows
// [P, cap_P) are padded, and GRAPH_FLAG_OVERFLOW is raised in the device header when the batch does not fit
int graph_fill(const void* edges, int in_dtype, int B, int N, int Ef, int by_type, int* ws, GraphArrays ga, int cap_E,
               int cap_P, cudaStream_t st);

}  // namespace gib
