// graph.cu - device-side build of the integer structures the loop consumes.
//
// Replaces (reference): GraphObject.buildArcNode / buildAdjacency / buildNodeGraph
// (GNN/graph_class.py:82-138), CompositeGraphObject.buildCompositeAdjacency / buildArcNode
// (GNN/composite_graph_class.py:57-103) and the tensorisation COO2SparseTensor + tf.sparse.reorder
// (graph_class.py:551-560).  Everything here is integer / bit-exact against oracle/structures.py:
//   dst-CSR  : rows = destination node, entries in increasing arc id (stable radix sort) - the order
//              in which TF-CPU SparseTensorDenseMatMul(adjoint_a=True) accumulates;
//   src-CSR  : rows = source node (backward: (Adj x)[i] = sum_{a: src_a = i} v_a x[dst_a]);
//   values   : ArcNode.data == Adjacency.data by aggregation_mode, float32((double)1/count);
//   mask list: rows with set_mask & output_mask (tf.boolean_mask order), type row lists
//              (tf.where(type_mask[t]) order), graph_ptr (nodes of a graph are contiguous after merge).
// CUB (shipped with the CUDA toolkit) provides the scan / stable radix sort / select primitives;
// the build is once per batch and not part of the per-iteration hot loop.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <vector>

#include "graph.h"

static __global__ void k_iota(int* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
static __global__ void k_count(const int* key, int n, int n_rows, int* cnt, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int k = key[i];
    if (k < 0 || k >= n_rows) { *bad = 1; return; }
    atomicAdd(cnt + k, 1);
  }
}
static __global__ void k_gather_i(const int* src, const int* idx, int* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}
static __global__ void k_gather_f(const float* src, const int* idx, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}
// ArcNode values in arc order (graph_class.py:107-121): float32 of a float64 quotient
static __global__ void k_values(int mode, int n_arcs, const int* dst, const int* indeg, const float* explicit_v,
                                float* val) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_arcs) return;
  float v = 1.0f;
  if (mode == GNNFP_AGG_EXPLICIT) v = explicit_v[a];
  else if (mode == GNNFP_AGG_NORMALIZED) v = (float)(1.0 * (1.0 / (double)n_arcs));
  else if (mode == GNNFP_AGG_AVERAGE) v = (float)(1.0 / (double)indeg[dst[a]]);
  val[a] = v;
}
// composite_average (composite_graph_class.py:92-99): for each type t in order, arcs whose source is of
// type t are divided by the number of such arcs into the same destination.
static __global__ void k_count_typed(const int* src, const int* dst, const uint8_t* tmask_t, int n_arcs, int* cnt) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n_arcs && tmask_t[src[a]]) atomicAdd(cnt + dst[a], 1);
}
static __global__ void k_div_typed(const int* src, const int* dst, const uint8_t* tmask_t, int n_arcs,
                                   const int* cnt, float* val) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n_arcs && tmask_t[src[a]]) val[a] = (float)((double)val[a] / (double)cnt[dst[a]]);
}
// CompositeAdjacencies[t] (composite_graph_class.py:57-70): Adjacency entries whose source is type t
// (zero-valued entries are eliminated, which is the same as weight 0 here), in dst-CSR order.
static __global__ void k_typed_weights(const int* csr_src, const float* csr_w, const uint8_t* tmask_t, int n_arcs,
                                       float* out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_arcs) out[p] = tmask_t[csr_src[p]] ? csr_w[p] : 0.0f;
}
static __global__ void k_mask_and(const uint8_t* a, const uint8_t* b, uint8_t* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint8_t)((a ? a[i] != 0 : 1) && (b ? b[i] != 0 : 1));
}
static __global__ void k_type_membership(const uint8_t* tmask, int n_types, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int c = 0;
    for (int t = 0; t < n_types; ++t) c += tmask[(size_t)t * n + i] != 0;
    if (c != 1) *bad = 1;
  }
}
static __global__ void k_graph_sizes(const int* node2graph, int n, int n_graphs, int* cnt, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int g = node2graph[i];
    if (g < 0 || g >= n_graphs) { *bad = 1; return; }
    if (i > 0 && node2graph[i - 1] > g) *bad = 2;   // nodes of a graph must be contiguous / ordered
    atomicAdd(cnt + g, 1);
  }
}
static __global__ void k_nodegraph_values(const int* node2graph, const int* cnt, int n, float* val) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) val[i] = (float)(1.0 * (1.0 / (double)cnt[node2graph[i]]));   // graph_class.py:136
}

// ---- tile-local view of the dst-CSR (fused forward iteration, rows_tma.cu) ------------------------------------------
// entry p of row j: lidx = src - 128 * (j / 128) if the source lies in j's 128-row tile, else -1
static __global__ void k_tile_lidx(const int* rowptr, const int* csr_src, int n, short* lidx, uint8_t* bnd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int t0 = j & ~(GNNFP_TILE_ROWS - 1);
  const int tend = t0 + GNNFP_TILE_ROWS < n ? t0 + GNNFP_TILE_ROWS : n;
  bool out = rowptr[tend] - rowptr[t0] > GNNFP_TILE_ARCS;     // too many arcs for the staged gather: the whole tile goes to the row-list pass
  for (int p = rowptr[j]; p < rowptr[j + 1]; ++p) {
    const int l = csr_src[p] - t0;
    const bool in = l >= 0 && l < GNNFP_TILE_ROWS;
    lidx[p] = in ? (short)l : (short)-1;
    out = out || !in;
  }
  bnd[j] = out ? 1 : 0;
}
static __global__ void k_tile_arc0(const int* rowptr, int n, int n_tiles, int* arc0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_tiles + 2) { const int r = i * GNNFP_TILE_ROWS; arc0[i] = rowptr[r < n ? r : n]; }
}
static __global__ void k_fill_i(int* p, int n, const int* value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = *value;
}

#define GRID(n) (((n) + 255) / 256), 256

// Stream-ordered allocations from the device's default memory pool (kept warm: release threshold
// = max), so building the structures of every batch costs no cudaMalloc after the first few steps.
struct DevAlloc {
  std::vector<void*>* list;
  size_t* bytes;
  cudaStream_t s;
  template <typename T>
  int get(T** p, size_t n) {
    void* q = nullptr;
    size_t b = (n ? n : 1) * sizeof(T);
    cudaError_t e = cudaMallocAsync(&q, b, s);
    if (e != cudaSuccess) {
      gnnfp_set_error("cudaMallocAsync(%zu) failed: %s", b, cudaGetErrorString(e));
      return GNNFP_E_CUDA;
    }
    list->push_back(q);
    *bytes += b;
    *p = (T*)q;
    return 0;
  }
};

// stable sort of arc ids by key -> CSR (rowptr, permutation)
static int build_csr(DevAlloc& A, const int* key, int n_arcs, int n_rows, int** rowptr, int** perm, void** tmp,
                     size_t* tmp_bytes, int* d_bad, cudaStream_t s, int rowptr_pad = 0) {
  int rc;
  if ((rc = A.get(rowptr, (size_t)n_rows + 1 + rowptr_pad))) return rc;
  if ((rc = A.get(perm, (size_t)n_arcs))) return rc;
  int *cnt = nullptr, *iota = nullptr, *key_out = nullptr;
  GNNFP_CHECK_CUDA(cudaMallocAsync(&cnt, sizeof(int) * ((size_t)n_rows + 1), s));
  GNNFP_CHECK_CUDA(cudaMallocAsync(&iota, sizeof(int) * (size_t)(n_arcs + 1), s));
  GNNFP_CHECK_CUDA(cudaMallocAsync(&key_out, sizeof(int) * (size_t)(n_arcs + 1), s));
  GNNFP_CHECK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)n_rows + 1), s));
  if (n_arcs > 0) {
    k_count<<<GRID(n_arcs), 0, s>>>(key, n_arcs, n_rows, cnt, d_bad);
    k_iota<<<GRID(n_arcs), 0, s>>>(iota, n_arcs);
  }
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, cnt, *rowptr, n_rows + 1, s);
  size_t need2 = 0;
  int bits = 1;
  while ((1ll << bits) < (long long)n_rows + 1 && bits < 31) ++bits;
  cub::DeviceRadixSort::SortPairs(nullptr, need2, key, key_out, iota, *perm, n_arcs, 0, bits, s);
  if (need2 > need) need = need2;
  if (need > *tmp_bytes) {
    if (*tmp) GNNFP_CHECK_CUDA(cudaFreeAsync(*tmp, s));
    GNNFP_CHECK_CUDA(cudaMallocAsync(tmp, need, s));
    *tmp_bytes = need;
  }
  size_t tb = *tmp_bytes;
  GNNFP_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(*tmp, tb, cnt, *rowptr, n_rows + 1, s));
  if (n_arcs > 0) {
    tb = *tmp_bytes;
    GNNFP_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(*tmp, tb, key, key_out, iota, *perm, n_arcs, 0, bits, s));
  }
  GNNFP_CHECK_CUDA(cudaFreeAsync(cnt, s));
  GNNFP_CHECK_CUDA(cudaFreeAsync(iota, s));
  GNNFP_CHECK_CUDA(cudaFreeAsync(key_out, s));
  return 0;
}

// compacted index list of set flags (ascending) -> *out (device), *count (host)
static int select_flagged(DevAlloc& A, const uint8_t* flags, int n, int** out, int* count, void** tmp,
                          size_t* tmp_bytes, cudaStream_t s) {
  int rc;
  if ((rc = A.get(out, (size_t)n))) return rc;
  int* d_num = nullptr;
  GNNFP_CHECK_CUDA(cudaMallocAsync(&d_num, sizeof(int), s));
  thrust::counting_iterator<int> it(0);
  size_t need = 0;
  cub::DeviceSelect::Flagged(nullptr, need, it, flags, *out, d_num, n, s);
  if (need > *tmp_bytes) {
    if (*tmp) GNNFP_CHECK_CUDA(cudaFreeAsync(*tmp, s));
    GNNFP_CHECK_CUDA(cudaMallocAsync(tmp, need, s));
    *tmp_bytes = need;
  }
  size_t tb = *tmp_bytes;
  GNNFP_CHECK_CUDA(cub::DeviceSelect::Flagged(*tmp, tb, it, flags, *out, d_num, n, s));
  GNNFP_CHECK_CUDA(cudaMemcpyAsync(count, d_num, sizeof(int), cudaMemcpyDeviceToHost, s));
  GNNFP_CHECK_CUDA(cudaStreamSynchronize(s));
  GNNFP_CHECK_CUDA(cudaFreeAsync(d_num, s));
  return 0;
}

extern "C" int gnnfp_graph_build(gnnfp_graph** out, const gnnfp_graph_desc* d, void* stream) {
  if (!out || !d) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: null argument");
  *out = nullptr;
  if (d->n_nodes <= 0 || d->n_arcs < 0) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: n_nodes=%d n_arcs=%d", d->n_nodes, d->n_arcs);
  if (d->n_arcs > 0 && (!d->src || !d->dst)) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: src/dst missing");
  if (d->n_types < 0 || d->n_types > GNNFP_MAX_TYPES) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: n_types=%d", d->n_types);
  if (d->n_types > 0 && !d->type_mask) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: type_mask missing");
  if (d->aggregation_mode < 0 || d->aggregation_mode > GNNFP_AGG_EXPLICIT) GNNFP_FAIL(GNNFP_E_INVALID, "ERROR: Unknown aggregation mode");
  if (d->aggregation_mode == GNNFP_AGG_EXPLICIT && !d->arc_values) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: explicit values missing");
  if (d->aggregation_mode == GNNFP_AGG_COMPOSITE_AVERAGE && d->n_types == 0) GNNFP_FAIL(GNNFP_E_INVALID, "composite_average needs type_mask");
  if (d->n_graphs > 0 && !d->node2graph) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: node2graph missing");
  if (d->mask_len != d->n_nodes && d->mask_len != d->n_arcs && d->mask_len != 0)
    GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: mask_len must be n_nodes or n_arcs");
  cudaStream_t s = (cudaStream_t)stream;
  gnnfp_graph* g = new gnnfp_graph();
  g->N = d->n_nodes; g->A = d->n_arcs; g->G = d->n_graphs; g->n_types = d->n_types;
  g->mode = d->aggregation_mode;
  g->mask_len = d->mask_len ? d->mask_len : d->n_nodes;
  DevAlloc A{&g->allocs, &g->device_bytes, s};
  g->stream = s;
  {
    static bool pool_set = false;
    if (!pool_set) {
      int dev = 0;
      cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
      pool_set = true;
    }
  }
  const int N = g->N, NA = g->A;
  int rc = 0;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int* d_bad = nullptr;
  int h_bad = 0;
#define BUILD_TRY(x) do { rc = (x); if (rc) { gnnfp_graph_free(g); return rc; } } while (0)
#define BUILD_CUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { gnnfp_set_error("CUDA error in graph_build: %s", cudaGetErrorString(_e)); gnnfp_graph_free(g); return GNNFP_E_CUDA; } } while (0)
  BUILD_TRY(A.get(&d_bad, (size_t)1));
  g->d_bad = d_bad;
  BUILD_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), s));
  // keep our own copy of src/dst (the caller's buffers may go away)
  BUILD_TRY(A.get(&g->src, (size_t)NA));
  BUILD_TRY(A.get(&g->dst, (size_t)NA));
  if (NA) {
    BUILD_CUDA(cudaMemcpyAsync(g->src, d->src, sizeof(int) * NA, cudaMemcpyDeviceToDevice, s));
    BUILD_CUDA(cudaMemcpyAsync(g->dst, d->dst, sizeof(int) * NA, cudaMemcpyDeviceToDevice, s));
  }
  // (the row pointer and the weights carry padding: the fused forward iteration stages fixed-size, 16-byte aligned slices)
  const int RP_PAD = 2 * GNNFP_TILE_ROWS + 8, ARC_PAD = GNNFP_TILE_ARCS + 32;
  BUILD_TRY(build_csr(A, g->dst, NA, N, &g->dst_rowptr, &g->dst_arc, &tmp, &tmp_bytes, d_bad, s, RP_PAD));
  k_fill_i<<<GRID(RP_PAD), 0, s>>>(g->dst_rowptr + N + 1, RP_PAD, g->dst_rowptr + N);
  BUILD_TRY(build_csr(A, g->src, NA, N, &g->src_rowptr, &g->src_arc, &tmp, &tmp_bytes, d_bad, s));
  BUILD_TRY(A.get(&g->dst_src, (size_t)NA));
  BUILD_TRY(A.get(&g->src_dst, (size_t)NA));
  BUILD_TRY(A.get(&g->arc_val, (size_t)NA));
  BUILD_TRY(A.get(&g->dst_w, (size_t)NA + ARC_PAD));
  BUILD_CUDA(cudaMemsetAsync(g->dst_w + NA, 0, sizeof(float) * ARC_PAD, s));
  BUILD_TRY(A.get(&g->src_w, (size_t)NA));
  if (NA) {
    k_gather_i<<<GRID(NA), 0, s>>>(g->src, g->dst_arc, g->dst_src, NA);
    k_gather_i<<<GRID(NA), 0, s>>>(g->dst, g->src_arc, g->src_dst, NA);
    // in-degree = rowptr difference; k_values needs counts per node: reuse a temp histogram
    int* indeg = nullptr;
    BUILD_CUDA(cudaMallocAsync(&indeg, sizeof(int) * (size_t)N, s));
    BUILD_CUDA(cudaMemsetAsync(indeg, 0, sizeof(int) * (size_t)N, s));
    k_count<<<GRID(NA), 0, s>>>(g->dst, NA, N, indeg, d_bad);
    const int base_mode = d->aggregation_mode == GNNFP_AGG_COMPOSITE_AVERAGE ? GNNFP_AGG_SUM : d->aggregation_mode;
    k_values<<<GRID(NA), 0, s>>>(base_mode, NA, g->dst, indeg, d->arc_values, g->arc_val);
    if (d->aggregation_mode == GNNFP_AGG_COMPOSITE_AVERAGE) {
      for (int t = 0; t < d->n_types; ++t) {
        BUILD_CUDA(cudaMemsetAsync(indeg, 0, sizeof(int) * (size_t)N, s));
        k_count_typed<<<GRID(NA), 0, s>>>(g->src, g->dst, d->type_mask + (size_t)t * N, NA, indeg);
        k_div_typed<<<GRID(NA), 0, s>>>(g->src, g->dst, d->type_mask + (size_t)t * N, NA, indeg, g->arc_val);
      }
    }
    BUILD_CUDA(cudaFreeAsync(indeg, s));
    k_gather_f<<<GRID(NA), 0, s>>>(g->arc_val, g->dst_arc, g->dst_w, NA);
    k_gather_f<<<GRID(NA), 0, s>>>(g->arc_val, g->src_arc, g->src_w, NA);
  }
  // tile-local CSR view + boundary row list (count stays on the device: no host synchronisation)
  {
    const int n_tiles = (N + GNNFP_TILE_ROWS - 1) / GNNFP_TILE_ROWS;
    uint8_t* bflag = nullptr;
    BUILD_TRY(A.get(&g->tile_lidx, (size_t)NA + ARC_PAD));
    BUILD_TRY(A.get(&g->tile_arc0, (size_t)n_tiles + 2));
    BUILD_TRY(A.get(&g->bnd_rows, (size_t)N));
    BUILD_TRY(A.get(&g->bnd_count, (size_t)1));
    BUILD_TRY(A.get(&bflag, (size_t)N));
    BUILD_CUDA(cudaMemsetAsync(g->tile_lidx + NA, 0xFF, sizeof(short) * ARC_PAD, s));
    k_tile_lidx<<<GRID(N), 0, s>>>(g->dst_rowptr, g->dst_src, N, g->tile_lidx, bflag);
    k_tile_arc0<<<GRID(n_tiles + 2), 0, s>>>(g->dst_rowptr, N, n_tiles, g->tile_arc0);
    thrust::counting_iterator<int> it(0);
    size_t need = 0;
    cub::DeviceSelect::Flagged(nullptr, need, it, bflag, g->bnd_rows, g->bnd_count, N, s);
    if (need > tmp_bytes) {
      if (tmp) BUILD_CUDA(cudaFreeAsync(tmp, s));
      BUILD_CUDA(cudaMallocAsync(&tmp, need, s));
      tmp_bytes = need;
    }
    size_t tb = tmp_bytes;
    BUILD_CUDA(cub::DeviceSelect::Flagged(tmp, tb, it, bflag, g->bnd_rows, g->bnd_count, N, s));
  }
  // masks -> index list
  if (!d->set_mask && !d->output_mask) {
    g->M = g->mask_len;             // all rows selected: no index list needed
  } else {
    uint8_t* m = nullptr;
    BUILD_TRY(A.get(&m, (size_t)g->mask_len));
    k_mask_and<<<GRID(g->mask_len), 0, s>>>(d->set_mask, d->output_mask, m, g->mask_len);
    g->mask = m;
    BUILD_TRY(select_flagged(A, m, g->mask_len, &g->mask_idx, &g->M, &tmp, &tmp_bytes, s));
  }
  // composite: type row lists + typed weights
  g->types_ok = 1;
  if (d->n_types > 0) {
    BUILD_TRY(A.get(&g->type_mask, (size_t)d->n_types * N));
    BUILD_CUDA(cudaMemcpyAsync(g->type_mask, d->type_mask, (size_t)d->n_types * N, cudaMemcpyDeviceToDevice, s));
    for (int t = 0; t < d->n_types; ++t) {
      BUILD_TRY(select_flagged(A, g->type_mask + (size_t)t * N, N, &g->type_rows[t], &g->type_count[t], &tmp, &tmp_bytes, s));
      BUILD_TRY(A.get(&g->typed_w[t], (size_t)NA));
      if (NA) k_typed_weights<<<GRID(NA), 0, s>>>(g->dst_src, g->dst_w, g->type_mask + (size_t)t * N, NA, g->typed_w[t]);
    }
    int* d_bad2 = nullptr;
    BUILD_CUDA(cudaMallocAsync(&d_bad2, sizeof(int), s));
    BUILD_CUDA(cudaMemsetAsync(d_bad2, 0, sizeof(int), s));
    k_type_membership<<<GRID(N), 0, s>>>(g->type_mask, d->n_types, N, d_bad2);
    int hb = 0;
    BUILD_CUDA(cudaMemcpyAsync(&hb, d_bad2, sizeof(int), cudaMemcpyDeviceToHost, s));
    BUILD_CUDA(cudaStreamSynchronize(s));
    BUILD_CUDA(cudaFreeAsync(d_bad2, s));
    g->types_ok = hb ? 0 : 1;
  }
  // NodeGraph
  if (d->n_graphs > 0) {
    BUILD_TRY(A.get(&g->node2graph, (size_t)N));
    BUILD_TRY(A.get(&g->ng_val, (size_t)N));
    BUILD_TRY(A.get(&g->graph_ptr, (size_t)d->n_graphs + 1));
    BUILD_CUDA(cudaMemcpyAsync(g->node2graph, d->node2graph, sizeof(int) * N, cudaMemcpyDeviceToDevice, s));
    int* cnt = nullptr;
    BUILD_CUDA(cudaMallocAsync(&cnt, sizeof(int) * ((size_t)d->n_graphs + 1), s));
    BUILD_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)d->n_graphs + 1), s));
    k_graph_sizes<<<GRID(N), 0, s>>>(g->node2graph, N, d->n_graphs, cnt, d_bad);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, cnt, g->graph_ptr, d->n_graphs + 1, s);
    if (need > tmp_bytes) {
      if (tmp) BUILD_CUDA(cudaFreeAsync(tmp, s));
      BUILD_CUDA(cudaMallocAsync(&tmp, need, s));
      tmp_bytes = need;
    }
    size_t tb = tmp_bytes;
    BUILD_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, g->graph_ptr, d->n_graphs + 1, s));
    if (d->nodegraph_values)
      BUILD_CUDA(cudaMemcpyAsync(g->ng_val, d->nodegraph_values, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
    else
      k_nodegraph_values<<<GRID(N), 0, s>>>(g->node2graph, cnt, N, g->ng_val);
    BUILD_CUDA(cudaFreeAsync(cnt, s));
  }
  if (tmp) BUILD_CUDA(cudaFreeAsync(tmp, s));
  BUILD_CUDA(cudaGetLastError());
  if (d->flags & GNNFP_GRAPH_DEFER_CHECK) { *out = g; return GNNFP_OK; }      // verdict read by gnnfp_graph_check
  BUILD_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s));
  BUILD_CUDA(cudaStreamSynchronize(s));
  if (h_bad == 1) { gnnfp_graph_free(g); GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: node / graph id out of range"); }
  if (h_bad == 2) { gnnfp_graph_free(g); GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "graph_build: nodes of a graph must be contiguous and in graph order (GraphObject.merge layout)"); }
  *out = g;
  return GNNFP_OK;
}

extern "C" int gnnfp_graph_check(const gnnfp_graph* g) {
  if (!g) GNNFP_FAIL(GNNFP_E_INVALID, "graph_check: null handle");
  int h_bad = 0;
  GNNFP_CHECK_CUDA(cudaMemcpyAsync(&h_bad, g->d_bad, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  GNNFP_CHECK_CUDA(cudaStreamSynchronize(g->stream));
  if (h_bad == 1) GNNFP_FAIL(GNNFP_E_INVALID, "graph_build: node / graph id out of range");
  if (h_bad == 2) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "graph_build: nodes of a graph must be contiguous and in graph order (GraphObject.merge layout)");
  return GNNFP_OK;
}

extern "C" void gnnfp_graph_free(gnnfp_graph* g) {
  if (!g) return;
  for (void* p : g->allocs) cudaFreeAsync(p, g->stream);
  delete g;
}

extern "C" int gnnfp_graph_get_info(const gnnfp_graph* g, gnnfp_graph_info* info) {
  if (!g || !info) GNNFP_FAIL(GNNFP_E_INVALID, "graph_get_info: null argument");
  memset(info, 0, sizeof(*info));
  info->n_nodes = g->N; info->n_arcs = g->A; info->n_graphs = g->G; info->n_types = g->n_types;
  info->n_masked = g->M;
  for (int t = 0; t < g->n_types; ++t) info->type_count[t] = g->type_count[t];
  info->types_disjoint_cover = g->types_ok;
  info->device_bytes = g->device_bytes;
  return GNNFP_OK;
}

extern "C" int gnnfp_graph_export(const gnnfp_graph* g, int which, void* host_dst, size_t bytes, void* stream) {
  if (!g || !host_dst) GNNFP_FAIL(GNNFP_E_INVALID, "graph_export: null argument");
  const void* p = nullptr;
  size_t n = 0;
  switch (which) {
    case GNNFP_X_DST_ROWPTR: p = g->dst_rowptr; n = sizeof(int) * ((size_t)g->N + 1); break;
    case GNNFP_X_DST_SRC: p = g->dst_src; n = sizeof(int) * (size_t)g->A; break;
    case GNNFP_X_DST_ARC: p = g->dst_arc; n = sizeof(int) * (size_t)g->A; break;
    case GNNFP_X_SRC_ROWPTR: p = g->src_rowptr; n = sizeof(int) * ((size_t)g->N + 1); break;
    case GNNFP_X_SRC_DST: p = g->src_dst; n = sizeof(int) * (size_t)g->A; break;
    case GNNFP_X_SRC_ARC: p = g->src_arc; n = sizeof(int) * (size_t)g->A; break;
    case GNNFP_X_ARC_VALUE: p = g->arc_val; n = sizeof(float) * (size_t)g->A; break;
    case GNNFP_X_MASK_INDEX:
      if (!g->mask_idx) {   // all rows selected
        if (sizeof(int) * (size_t)g->M > bytes) GNNFP_FAIL(GNNFP_E_INVALID, "graph_export: buffer too small");
        for (int i = 0; i < g->M; ++i) ((int*)host_dst)[i] = i;
        return GNNFP_OK;
      }
      p = g->mask_idx; n = sizeof(int) * (size_t)g->M; break;
    case GNNFP_X_GRAPH_PTR: p = g->graph_ptr; n = g->G ? sizeof(int) * ((size_t)g->G + 1) : 0; break;
    case GNNFP_X_NODEGRAPH_VALUE: p = g->ng_val; n = g->G ? sizeof(float) * (size_t)g->N : 0; break;
    case GNNFP_X_TYPE_ROWS: {
      size_t off = 0;
      for (int t = 0; t < g->n_types; ++t) {
        size_t b = sizeof(int) * (size_t)g->type_count[t];
        if (off + b > bytes) GNNFP_FAIL(GNNFP_E_INVALID, "graph_export: buffer too small");
        GNNFP_CHECK_CUDA(cudaMemcpyAsync((char*)host_dst + off, g->type_rows[t], b, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        off += b;
      }
      GNNFP_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
      return GNNFP_OK;
    }
    default: GNNFP_FAIL(GNNFP_E_INVALID, "graph_export: unknown array %d", which);
  }
  if (n > bytes) GNNFP_FAIL(GNNFP_E_INVALID, "graph_export: buffer too small (%zu > %zu)", n, bytes);
  if (n) GNNFP_CHECK_CUDA(cudaMemcpyAsync(host_dst, p, n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GNNFP_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return GNNFP_OK;
}
