// graph.h - the opaque graph handle of gnnfp.h
#pragma once
#include <vector>

#include "common.cuh"

struct gnnfp_graph {
  int N = 0, A = 0, G = 0, n_types = 0, mode = 0, mask_len = 0, M = 0;
  int* src = nullptr;          // [A] arcs[:,0]
  int* dst = nullptr;          // [A] arcs[:,1]
  int* dst_rowptr = nullptr;   // [N+1] destination-grouped CSR
  int* dst_arc = nullptr;      // [A]  arc id per entry (ascending inside a row)
  int* dst_src = nullptr;      // [A]  source node per entry
  float* dst_w = nullptr;      // [A]  value per entry
  int* src_rowptr = nullptr;   // [N+1] source-grouped CSR
  int* src_arc = nullptr;
  int* src_dst = nullptr;
  float* src_w = nullptr;
  float* arc_val = nullptr;    // [A] in arc order
  uint8_t* mask = nullptr;     // [mask_len] set_mask & output_mask
  int* mask_idx = nullptr;     // [M]
  uint8_t* type_mask = nullptr;             // [n_types, N]
  int* type_rows[GNNFP_MAX_TYPES] = {};     // rows of each type, ascending
  int type_count[GNNFP_MAX_TYPES] = {};
  float* typed_w[GNNFP_MAX_TYPES] = {};     // dst-CSR weights of CompositeAdjacencies[t]
  int types_ok = 1;                         // every node in exactly one type
  // tile-local view of the dst-CSR for the fused forward iteration (rows_tma.cu): 128-row tiles
  short* tile_lidx = nullptr;  // [A + pad] source row inside the destination's tile (dst_src - 128 * (row / 128)), -1 = outside
  int* tile_arc0 = nullptr;    // [ceil(N/128) + 2] dst_rowptr[128 i]
  int* bnd_rows = nullptr;     // rows with an in-arc from outside their tile (or whose tile holds too many arcs), ascending
  int* bnd_count = nullptr;    // device scalar: entries of bnd_rows
  int* node2graph = nullptr;   // [N]
  float* ng_val = nullptr;     // [N]
  int* graph_ptr = nullptr;    // [G+1]
  int* d_bad = nullptr;        // device word of the id validation (0 ok, 1 id out of range, 2 nodes of a graph not contiguous)
  cudaStream_t stream = nullptr;   // stream the device arrays were allocated on (stream-ordered pool)
  std::vector<void*> allocs;
  size_t device_bytes = 0;
};
