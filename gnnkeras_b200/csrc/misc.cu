// misc.cu - small device ops around the loop: LGNN.update_graph (reference LGNN.py:175-214) and the
// fused train-step tail (Keras categorical_crossentropy + Adam).
#include "graph.h"

// Row-wise kernels: one warp per row (lane = column, 4 rows in flight) instead of one thread per element - a per-element
// 64-bit division by the runtime row width costs ~50 instructions on the conversion (XU) pipe and made these copies 2x
// slower than their traffic.
static __global__ void k_update_graph_fwd(int n, const float* state, int sw, int ow, const float* base, int bw, int ldb, float* dst) {
  const int W = sw + ow + bw;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = wid * 4; r0 < n; r0 += nw * 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
      if (r >= n) break;
      for (int c = lane; c < W; c += 32) {
        float v;
        if (c < sw) v = state[(size_t)r * sw + c];
        else if (c < sw + ow) v = 0.0f;                     // tf.scatter_nd zero fill (LGNN.py:203)
        else v = base[(size_t)r * ldb + (c - sw - ow)];
        dst[(size_t)r * W + c] = v;
      }
    }
  }
}
static __global__ void k_scatter_rows(int m, const int* idx, const float* rows, int ow, float* dst, int W, int col0) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    const size_t row = idx ? (size_t)idx[r] : (size_t)r;
    for (int c = 0; c < ow; ++c) dst[row * W + col0 + c] = rows[(size_t)r * ow + c];
  }
}
static __global__ void k_update_graph_bwd(int n, const float* d_dst, float* d_state, int sw, int ow, float* d_base, int bw, int accumulate) {
  const int W = sw + ow + bw;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = wid * 4; r0 < n; r0 += nw * 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
      if (r >= n) break;
      for (int c = lane; c < W; c += 32) {
        if (c >= sw && c < sw + ow) continue;
        if (c < sw && !d_state) continue;
        if (c >= sw + ow && !d_base) continue;
        const float v = d_dst[(size_t)r * W + c];
        if (c < sw) d_state[(size_t)r * sw + c] = v;
        else { float* d = d_base + (size_t)r * bw + (c - sw - ow); if (accumulate) *d += v; else *d = v; }
      }
    }
  }
}
static __global__ void k_gather_rows(int m, const int* idx, const float* src, int W, int col0, int ow, float* rows) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
    const size_t row = idx ? (size_t)idx[r] : (size_t)r;
    for (int c = 0; c < ow; ++c) rows[(size_t)r * ow + c] = src[row * W + col0 + c];
  }
}

static int blocks_for(size_t total) {
  size_t b = (total + 255) / 256;
  if (b > 4736) b = 4736;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int gnnfp_update_graph_forward(const gnnfp_graph* g, int32_t n_rows, const float* state, int32_t state_w,
                                          const float* out_rows, int32_t out_w, const float* base, int32_t base_w,
                                          int32_t ld_base, float* dst, void* stream) {
  if (!g || !dst || !base) GNNFP_FAIL(GNNFP_E_INVALID, "update_graph: null argument");
  if (n_rows != g->mask_len) GNNFP_FAIL(GNNFP_E_INVALID, "update_graph: n_rows must equal the mask length");
  if ((state_w > 0 && !state) || (out_w > 0 && !out_rows)) GNNFP_FAIL(GNNFP_E_INVALID, "update_graph: state / out rows missing");
  cudaStream_t s = (cudaStream_t)stream;
  const int W = state_w + out_w + base_w;
  k_update_graph_fwd<<<blocks_for((size_t)n_rows * 8), 256, 0, s>>>(n_rows, state, state_w, out_w, base, base_w, ld_base, dst);
  GNNFP_COUNT_LAUNCH();
  if (out_w > 0 && g->M > 0) {
    const int* idx = g->M == g->mask_len ? nullptr : g->mask_idx;
    k_scatter_rows<<<blocks_for((size_t)g->M), 256, 0, s>>>(g->M, idx, out_rows, out_w, dst, W, state_w);
    GNNFP_COUNT_LAUNCH();
  }
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

extern "C" int gnnfp_update_graph_backward(const gnnfp_graph* g, int32_t n_rows, const float* d_dst, float* d_state,
                                           int32_t state_w, float* d_out_rows, int32_t out_w, float* d_base,
                                           int32_t base_w, int32_t accumulate_base, void* stream) {
  if (!g || !d_dst) GNNFP_FAIL(GNNFP_E_INVALID, "update_graph_backward: null argument");
  if (n_rows != g->mask_len) GNNFP_FAIL(GNNFP_E_INVALID, "update_graph_backward: n_rows must equal the mask length");
  cudaStream_t s = (cudaStream_t)stream;
  const int W = state_w + out_w + base_w;
  k_update_graph_bwd<<<blocks_for((size_t)n_rows * 8), 256, 0, s>>>(n_rows, d_dst, d_state, state_w, out_w, d_base, base_w, accumulate_base);
  GNNFP_COUNT_LAUNCH();
  if (out_w > 0 && d_out_rows && g->M > 0) {
    const int* idx = g->M == g->mask_len ? nullptr : g->mask_idx;
    k_gather_rows<<<blocks_for((size_t)g->M), 256, 0, s>>>(g->M, idx, d_dst, W, state_w, out_w, d_out_rows);
    GNNFP_COUNT_LAUNCH();
  }
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// Keras categorical_crossentropy on probabilities: p/sum(p), clip to [1e-7, 1-1e-7],
// -sum y log p; reduction SUM_OVER_BATCH_SIZE with sample weights: sum_i w_i l_i / rows.
static __global__ void k_cce(const float* y, const float* p, const float* w, int rows, int cols, float scale,
                             float* loss, float* dp) {
  float local = 0.f;
  const float eps = 1e-7f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
    const float* pi = p + (size_t)i * cols;
    const float* yi = y + (size_t)i * cols;
    float s = 0.f;
    for (int c = 0; c < cols; ++c) s += pi[c];
    const float wi = (w ? w[i] : 1.0f) * scale / (float)rows;
    float li = 0.f, dot = 0.f;
    for (int c = 0; c < cols; ++c) {
      const float q = pi[c] / s;
      const float qc = fminf(fmaxf(q, eps), 1.0f - eps);
      li -= yi[c] * logf(qc);
      if (q >= eps && q <= 1.0f - eps) dot += (-yi[c] / qc) * q;
    }
    local += wi * li;
    if (dp) {
      for (int c = 0; c < cols; ++c) {
        const float q = pi[c] / s;
        const float qc = fminf(fmaxf(q, eps), 1.0f - eps);
        const float dq = (q >= eps && q <= 1.0f - eps) ? (-yi[c] / qc) : 0.0f;
        dp[(size_t)i * cols + c] = wi * (dq - dot) / s;
      }
    }
  }
  // block reduction, one atomic per block
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x + 31) / 32 ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, v);
  }
}

extern "C" int gnnfp_cce_loss(const float* y_true, const float* y_pred, const float* sample_weight, int32_t rows,
                              int32_t cols, float scale, float* loss_out, float* d_pred, void* stream) {
  if (!y_true || !y_pred || rows <= 0 || cols <= 0) GNNFP_FAIL(GNNFP_E_INVALID, "cce_loss: bad arguments");
  int blocks = (rows + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  k_cce<<<blocks, 256, 0, (cudaStream_t)stream>>>(y_true, y_pred, sample_weight, rows, cols, scale, loss_out, d_pred);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// Keras-2 Adam (non-amsgrad): alpha = lr*sqrt(1-b2^t)/(1-b1^t); m,v updates; p -= alpha*m/(sqrt(v)+eps)
static __global__ void k_adam(float* p, const float* g, float* m, float* v, size_t n, float alpha, float b1, float b2,
                              float eps, float gscale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.0f - b2);
    m[i] = mi;
    v[i] = vi;
    p[i] -= (mi * alpha) / (sqrtf(vi) + eps);
  }
}
// The same update with the step count on the device (int32 counter, incremented by gnnfp_adam_advance after the last span of
// an optimisation step): nothing of the step is baked into kernel arguments, so a whole train step can be captured into one
// CUDA graph and replayed.
static __global__ void k_adam_dev(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
                                  float eps, float gscale, const int* step_dev) {
  __shared__ float s_alpha;
  if (threadIdx.x == 0) {
    const double t = (double)(*step_dev + 1);
    s_alpha = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  __syncthreads();
  const float alpha = s_alpha;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.0f - b2);
    m[i] = mi;
    v[i] = vi;
    p[i] -= (mi * alpha) / (sqrtf(vi) + eps);
  }
}
static __global__ void k_adam_advance(int* step_dev) { *step_dev += 1; }
extern "C" int gnnfp_adam_step_dev(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                                   float beta2, float eps, const int32_t* step_dev, float grad_scale, void* stream) {
  if (!params || !grads || !m || !v || !step_dev) GNNFP_FAIL(GNNFP_E_INVALID, "adam_step_dev: bad arguments");
  if (n == 0) return GNNFP_OK;
  k_adam_dev<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, lr, beta1, beta2, eps, grad_scale, step_dev);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
extern "C" int gnnfp_adam_advance(int32_t* step_dev, void* stream) {
  if (!step_dev) GNNFP_FAIL(GNNFP_E_INVALID, "adam_advance: null counter");
  k_adam_advance<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

extern "C" int gnnfp_adam_step(float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                               float beta2, float eps, int32_t step, float grad_scale, void* stream) {
  if (!params || !grads || !m || !v || step < 1) GNNFP_FAIL(GNNFP_E_INVALID, "adam_step: bad arguments");
  if (n == 0) return GNNFP_OK;
  const double alpha = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  k_adam<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(params, grads, m, v, n, (float)alpha, beta1, beta2, eps, grad_scale);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// ---- device-side batcher (reference GraphObject.merge, graph_class.py:385-413, called for EVERY batch at
// ---- construction and at every epoch end by GraphSequencers.py:42-46, 123-127) -------------------------------------------
// The dataset lives flat on the device (gnnfp_store_desc: members back to back, arc ids LOCAL to their member); a batch is
// the list of member ids.  Two launches: the exclusive prefix sums of the selected members' sizes (one block), then ONE
// block per member that copies its rows to their place in the batch - node rows, arc rows with the node-id offset added
// to columns 0-1 (graph_class.py:391-394) and split off as int32 src / dst, targets, sample weights, masks, NodeGraph
// entries with the sub-graph offset (block-diagonal NodeGraph, :407), type masks transposed to [n_types, N].
static __global__ void __launch_bounds__(1024) k_batch_offsets(const gnnfp_store_desc st, const int64_t* ids, int n_ids, int64_t* off) {
  // off: [5][n_ids + 1] = nodes | arcs | targets | masks | sub-graphs
  __shared__ long long carry[5];
  __shared__ long long part[5][1024];
  if (threadIdx.x < 5) carry[threadIdx.x] = 0;
  __syncthreads();
  for (int base = 0; base < n_ids; base += 1024) {
    const int i = base + threadIdx.x;
    long long v[5] = {0, 0, 0, 0, 0};
    if (i < n_ids) {
      const int64_t m = ids[i];
      v[0] = st.node_ptr[m + 1] - st.node_ptr[m];
      v[1] = st.arc_ptr[m + 1] - st.arc_ptr[m];
      v[2] = st.tgt_ptr[m + 1] - st.tgt_ptr[m];
      v[3] = st.mask_ptr ? st.mask_ptr[m + 1] - st.mask_ptr[m] : 0;
      v[4] = st.n_sub ? st.n_sub[m] : 0;
    }
    for (int q = 0; q < 5; ++q) part[q][threadIdx.x] = v[q];
    __syncthreads();
    for (int s = 1; s < 1024; s <<= 1) {               // Hillis-Steele inclusive scan of the chunk
      long long t[5];
      for (int q = 0; q < 5; ++q) t[q] = threadIdx.x >= s ? part[q][threadIdx.x - s] : 0;
      __syncthreads();
      for (int q = 0; q < 5; ++q) part[q][threadIdx.x] += t[q];
      __syncthreads();
    }
    if (i < n_ids)
      for (int q = 0; q < 5; ++q) off[(size_t)q * (n_ids + 1) + i] = carry[q] + part[q][threadIdx.x] - v[q];
    __syncthreads();
    if (threadIdx.x == 1023)
      for (int q = 0; q < 5; ++q) carry[q] += part[q][1023];
    __syncthreads();
  }
  if (threadIdx.x < 5) off[(size_t)threadIdx.x * (n_ids + 1) + n_ids] = carry[threadIdx.x];
}

static __global__ void __launch_bounds__(256) k_batch_assemble(const gnnfp_store_desc st, const int64_t* ids, int n_ids,
                                                                const int64_t* off, const gnnfp_batch_out out, int64_t n_nodes_batch) {
  const int b = blockIdx.x;
  const int64_t m = ids[b];
  const int64_t* on = off, *oa = off + (n_ids + 1), *ot = off + 2 * (size_t)(n_ids + 1), *om = off + 3 * (size_t)(n_ids + 1),
               *os = off + 4 * (size_t)(n_ids + 1);
  const int64_t n0 = st.node_ptr[m], nn = st.node_ptr[m + 1] - n0, dn = on[b];
  const int64_t a0 = st.arc_ptr[m], na = st.arc_ptr[m + 1] - a0, da = oa[b];
  const int64_t t0 = st.tgt_ptr[m], nt = st.tgt_ptr[m + 1] - t0, dt = ot[b];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int NW = st.nodes_width, AW = st.arcs_width, TW = st.targets_width;
  for (int64_t e = tid; e < nn * NW; e += nth) out.nodes[dn * NW + e] = st.nodes[n0 * NW + e];
  const float foff = (float)dn;                         // node ids are stored as float32 in arcs (exact below 2^24)
  for (int64_t e = tid; e < na * AW; e += nth) {
    const int64_t r = e / AW;
    const int c = (int)(e - r * AW);
    float v = st.arcs[a0 * AW + e];
    if (c < 2) {
      v += foff;
      (c == 0 ? out.src : out.dst)[da + r] = (int32_t)v;
    }
    out.arcs[da * AW + e] = v;
  }
  for (int64_t e = tid; e < nt * TW; e += nth) out.targets[dt * TW + e] = st.targets[t0 * TW + e];
  for (int64_t e = tid; e < nt; e += nth) out.sample_weight[dt + e] = st.sample_weight[t0 + e];
  if (st.set_mask && out.set_mask) {
    const int64_t m0 = st.mask_ptr[m], nm = st.mask_ptr[m + 1] - m0, dm = om[b];
    for (int64_t e = tid; e < nm; e += nth) { out.set_mask[dm + e] = st.set_mask[m0 + e]; out.output_mask[dm + e] = st.output_mask[m0 + e]; }
  }
  if (st.node2graph && out.node2graph) {
    const int32_t so = (int32_t)os[b];
    for (int64_t e = tid; e < nn; e += nth) { out.node2graph[dn + e] = st.node2graph[n0 + e] + so; out.nodegraph_values[dn + e] = st.nodegraph_values[n0 + e]; }
  }
  if (st.type_mask && out.type_mask)
    for (int64_t e = tid; e < nn * st.n_types; e += nth) {
      const int64_t i = e / st.n_types;
      const int t = (int)(e - i * st.n_types);
      out.type_mask[(size_t)t * n_nodes_batch + dn + i] = st.type_mask[(n0 + i) * st.n_types + t];
    }
}

extern "C" int gnnfp_batch_assemble(const gnnfp_store_desc* st, const int64_t* ids_dev, int32_t n_ids, int64_t n_nodes_batch,
                                    int64_t* offsets_scratch, const gnnfp_batch_out* out, void* stream) {
  if (!st || !ids_dev || !out || !offsets_scratch || n_ids < 1) GNNFP_FAIL(GNNFP_E_INVALID, "batch_assemble: bad arguments");
  if (!st->nodes || !st->targets || !st->sample_weight || !st->node_ptr || !st->arc_ptr || !st->tgt_ptr)
    GNNFP_FAIL(GNNFP_E_INVALID, "batch_assemble: store arrays missing");
  if (!out->nodes || !out->targets || !out->sample_weight)      // (arc arrays may be empty: a batch of arc-less members)
    GNNFP_FAIL(GNNFP_E_INVALID, "batch_assemble: output arrays missing");
  cudaStream_t s = (cudaStream_t)stream;
  k_batch_offsets<<<1, 1024, 0, s>>>(*st, ids_dev, n_ids, offsets_scratch);
  GNNFP_COUNT_LAUNCH();
  k_batch_assemble<<<n_ids, 256, 0, s>>>(*st, ids_dev, n_ids, offsets_scratch, *out, n_nodes_batch);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}


// ---- measurement aid: sustained FP32 FMA rate of this GPU (the roofline denominator of the FP32-pipe tile kernels) -------------
// 8 independent FMA chains per thread, no memory traffic; bench.py times one launch with CUDA events.
static __global__ void __launch_bounds__(256) k_fma_peak(float* sink, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 12345.678f) sink[0] = r;                    // never true: keeps the chains alive
}
// launches blocks_per_sm x SMs blocks of 256 threads; *flops_out = FLOPs of the launch (2 per FMA)
extern "C" int gnnfp_debug_fma_peak(float* sink, int32_t iters, int32_t blocks_per_sm, double* flops_out, void* stream) {
  if (!sink || iters < 1 || blocks_per_sm < 1) GNNFP_FAIL(GNNFP_E_INVALID, "fma_peak: bad arguments");
  const int grid = gnnfp_num_sms() * blocks_per_sm;
  k_fma_peak<<<grid, 256, 0, (cudaStream_t)stream>>>(sink, iters, 0.999f, 0.001f);
  GNNFP_CHECK_CUDA(cudaGetLastError());
  if (flops_out) *flops_out = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)grid;
  return GNNFP_OK;
}
