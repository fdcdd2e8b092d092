// narrow.cu - net_output with a NARROW single Dense layer (H <= 4 output columns: the starters' Dense(2, softmax),
// reference starter.py:28, GNN.py:273) as three streaming kernels.  With 2 output columns the layer is a pair of dot products
// per row: a GEMM tile kernel spends its time on padded tiles (measured 53-78 us forward and 36-76 us per backward GEMM on the
// bench batch), while the rows themselves stream in ~13 us at the HBM rate.  Mapping: one warp per row (4 rows in flight),
// lane = input column (columns lane, lane + 32, ... of every input piece), warp shuffles for the row sums.
//   narrow_fwd   out[r] = act(x[r] . Wp + bias)            (Wp / bias: the BN-folded padded weights of fold_w_kernel)
//   narrow_dw    dW = x^T dz, db, BN sums P / Q            (per-block partial slot, same algebra as gemm_dw_kernel)
//   narrow_dx    dx[r] = colscale * (dz[r] . W^T) - BN-training correction, written (or added) per input piece
#include "gemm.h"
#include "tile.cuh"

#define NR_MAXSLOTS 6                    // column slots per lane: sum over pieces of ceil(width / 32); kernels are instantiated for 2, 3 and 6
#define NR_ROWS 4                        // rows in flight per warp
#define NR_THREADS 256

struct NrSlot { const float* base; int ld; int c; int kp; bool ok; };   // piece base + column, real input column, padded weight row

// slot q of this lane: column lane + 32 * i of piece p (pieces in order); `c` = real input column, `kp` = row of Wp.
// Called with a compile-time q from unrolled loops, so that the slot tables stay in registers.
__device__ __forceinline__ void nr_slot(const NarrowArgs& a, int lane, int q, NrSlot& s, int& piece, int& col) {
  s.ok = false; s.base = a.p[0].ptr; s.ld = 0; s.c = 0; s.kp = 0;
  piece = -1; col = 0;
  int q0 = 0, coff = 0;
#pragma unroll
  for (int p = 0; p < GEMM_MAXP; ++p) {
    if (p < a.n_pieces) {
      const int ns = (a.p[p].width + 31) >> 5;
      if (q >= q0 && q < q0 + ns) {
        col = lane + 32 * (q - q0);
        piece = p;
        s.ok = col < a.p[p].width;
        s.base = a.p[p].ptr + (s.ok ? col : 0);
        s.ld = a.p[p].ld;
        s.c = coff + (s.ok ? col : 0);
        s.kp = a.p[p].k8 + (s.ok ? col : 0);
      }
      q0 += ns; coff += a.p[p].width;
    }
  }
}
#define NR_SLOT_TABLE()                                           \
  NrSlot sl[NR_SLOTS];                                            \
  int sl_piece[NR_SLOTS], sl_col[NR_SLOTS];                       \
  _Pragma("unroll") for (int q = 0; q < NR_SLOTS; ++q) nr_slot(a, lane, q, sl[q], sl_piece[q], sl_col[q]);

template <int H, int NR_SLOTS>
__global__ void __launch_bounds__(NR_THREADS) narrow_fwd_kernel(const __grid_constant__ NarrowArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * NR_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * NR_THREADS) >> 5;
  NR_SLOT_TABLE();
  float w[NR_SLOTS][H], bias[H];
#pragma unroll
  for (int q = 0; q < NR_SLOTS; ++q)
#pragma unroll
    for (int j = 0; j < H; ++j) w[q][j] = sl[q].ok ? a.Wp[(size_t)sl[q].kp * a.ldw + j] : 0.f;
#pragma unroll
  for (int j = 0; j < H; ++j) bias[j] = a.bias ? a.bias[j] : 0.f;
  for (int r0 = wid * NR_ROWS; r0 < a.n_rows; r0 += nw * NR_ROWS) {
    float acc[NR_ROWS][H];
#pragma unroll
    for (int u = 0; u < NR_ROWS; ++u) {
      const int r = r0 + u;
      const int gr = r < a.n_rows ? (a.rowlist ? a.rowlist[r] : r) : -1;
      float x[NR_SLOTS];
#pragma unroll
      for (int q = 0; q < NR_SLOTS; ++q) x[q] = (gr >= 0 && sl[q].ok) ? sl[q].base[(size_t)gr * sl[q].ld] : 0.f;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < NR_SLOTS; ++q) t = fmaf(x[q], w[q][j], t);
        acc[u][j] = t;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < NR_ROWS; ++u)
#pragma unroll
        for (int j = 0; j < H; ++j) acc[u][j] += __shfl_xor_sync(0xffffffffu, acc[u][j], o);
    if (lane < NR_ROWS && r0 + lane < a.n_rows) {      // lane u finishes row r0 + u
      float v[H];
#pragma unroll
      for (int j = 0; j < H; ++j) {
        float t = acc[0][j];
#pragma unroll
        for (int u = 1; u < NR_ROWS; ++u) t = lane == u ? acc[u][j] : t;
        v[j] = act_fwd(a.act, t + bias[j]);
      }
      if (a.act == GNNFP_ACT_SOFTMAX) {                // Keras softmax over the row
        float mx = v[0];
#pragma unroll
        for (int j = 1; j < H; ++j) mx = fmaxf(mx, v[j]);
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) { v[j] = expf(v[j] - mx); se += v[j]; }
#pragma unroll
        for (int j = 0; j < H; ++j) v[j] = v[j] / se;
      }
      float* o = a.out + (size_t)(r0 + lane) * a.ld_out;
#pragma unroll
      for (int j = 0; j < H; ++j) o[j] = v[j];
    }
  }
}

template <int H, int NR_SLOTS>
__global__ void __launch_bounds__(NR_THREADS) narrow_dw_kernel(const __grid_constant__ NarrowArgs a) {
  __shared__ float s_acc[NR_THREADS / 32][NR_SLOTS * 32 * H + H];
  __shared__ float s_db[H], s_d[NR_SLOTS * 32 * H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wid = (blockIdx.x * NR_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * NR_THREADS) >> 5;
  NR_SLOT_TABLE();
  float g[NR_SLOTS][H], db[H];
#pragma unroll
  for (int q = 0; q < NR_SLOTS; ++q)
#pragma unroll
    for (int j = 0; j < H; ++j) g[q][j] = 0.f;
#pragma unroll
  for (int j = 0; j < H; ++j) db[j] = 0.f;
  for (int r0 = wid * NR_ROWS; r0 < a.n_rows; r0 += nw * NR_ROWS) {
    float x[NR_ROWS][NR_SLOTS], dz[NR_ROWS][H];
#pragma unroll
    for (int u = 0; u < NR_ROWS; ++u) {
      const int r = r0 + u;
      const int gr = r < a.n_rows ? (a.rowlist ? a.rowlist[r] : r) : -1;
#pragma unroll
      for (int q = 0; q < NR_SLOTS; ++q) x[u][q] = (gr >= 0 && sl[q].ok) ? sl[q].base[(size_t)gr * sl[q].ld] : 0.f;
#pragma unroll
      for (int j = 0; j < H; ++j) dz[u][j] = gr >= 0 ? a.dz[(size_t)r * H + j] : 0.f;      // compact dz: by position in the row set
    }
#pragma unroll
    for (int u = 0; u < NR_ROWS; ++u) {
#pragma unroll
      for (int j = 0; j < H; ++j) {
        db[j] += dz[u][j];
#pragma unroll
        for (int q = 0; q < NR_SLOTS; ++q) g[q][j] = fmaf(x[u][q], dz[u][j], g[q][j]);
      }
    }
  }
  // block reduction over the warps: s_d[(q * 32 + lane) * H + j], s_db[j]
#pragma unroll
  for (int q = 0; q < NR_SLOTS; ++q)
#pragma unroll
    for (int j = 0; j < H; ++j) s_acc[warp][(q * 32 + lane) * H + j] = g[q][j];
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < H; ++j) s_acc[warp][NR_SLOTS * 32 * H + j] = db[j];
  __syncthreads();
  for (int e = threadIdx.x; e < NR_SLOTS * 32 * H + H; e += NR_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < NR_THREADS / 32; ++w2) t += s_acc[w2][e];
    if (e < NR_SLOTS * 32 * H) s_d[e] = t; else s_db[e - NR_SLOTS * 32 * H] = t;
  }
  __syncthreads();
  // flush: this block's partial slot (+=, as gemm_dw_kernel), BN algebra, BN sums P / Q
  float* part = a.partial + (size_t)blockIdx.x * a.n_params;
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NR_SLOTS; ++q) {
      if (!sl[q].ok) continue;
      const int c = sl[q].c;
      float P = 0.f, Q = 0.f;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        float v = s_d[(q * 32 + lane) * H + j];
        if (a.bn_partial) { const float wv = a.W[(size_t)c * H + j]; Q = fmaf(wv, v, Q); P = fmaf(wv, s_db[j], P); }
        if (a.bnA) v = a.gamma[c] * fmaf(a.bnA[c], v, a.bnB[c] * s_db[j]) + a.beta[c] * s_db[j];
        part[(size_t)c * H + j] += v;
      }
      if (a.bn_partial) {
        float* bp = a.bn_partial + (size_t)blockIdx.x * 2 * a.K;
        bp[c] = P;
        bp[a.K + c] = fmaf(a.bnA[c], Q, a.bnB[c] * P);
      }
    }
    if (lane < H) part[a.bias_off + lane] += s_db[lane];
  }
}

template <int H, int NR_SLOTS>
__global__ void __launch_bounds__(NR_THREADS) narrow_dx_kernel(const __grid_constant__ NarrowArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * NR_THREADS + threadIdx.x) >> 5, nw = (gridDim.x * NR_THREADS) >> 5;
  NR_SLOT_TABLE();
  float w[NR_SLOTS][H], k0[NR_SLOTS], k1[NR_SLOTS], kA[NR_SLOTS], kB[NR_SLOTS];
  float* gout[NR_SLOTS];
  int gld[NR_SLOTS];
  bool gadd[NR_SLOTS];
#pragma unroll
  for (int q = 0; q < NR_SLOTS; ++q) {
    gout[q] = nullptr; gld[q] = 0; gadd[q] = false;
#pragma unroll
    for (int p = 0; p < GEMM_MAXP; ++p)
      if (sl_piece[q] == p && sl[q].ok && a.gout[p]) { gout[q] = a.gout[p] + sl_col[q]; gld[q] = a.gld[p]; gadd[q] = a.gadd[p] != 0; }
  }
#pragma unroll
  for (int q = 0; q < NR_SLOTS; ++q) {
    const int c = sl[q].c;
    const float cs = (a.colscale && gout[q]) ? a.colscale[c] : 1.f;
#pragma unroll
    for (int j = 0; j < H; ++j) w[q][j] = gout[q] ? cs * a.W[(size_t)c * H + j] : 0.f;
    const bool cr = a.corr != nullptr && gout[q] != nullptr;
    k0[q] = cr ? a.corr[c] : 0.f; k1[q] = cr ? a.corr[a.corr_in + c] : 0.f;
    kA[q] = cr ? a.corr[2 * a.corr_in + c] : 0.f; kB[q] = cr ? a.corr[3 * a.corr_in + c] : 0.f;
  }
  for (int r0 = wid * NR_ROWS; r0 < a.n_rows; r0 += nw * NR_ROWS) {
    float x[NR_ROWS][NR_SLOTS], dz[NR_ROWS][H], old[NR_ROWS][NR_SLOTS];
    int grs[NR_ROWS];
#pragma unroll
    for (int u = 0; u < NR_ROWS; ++u) {
      const int r = r0 + u;
      const int gr = r < a.n_rows ? (a.rowlist ? a.rowlist[r] : r) : -1;
      grs[u] = gr;
#pragma unroll
      for (int q = 0; q < NR_SLOTS; ++q) {
        x[u][q] = (gr >= 0 && gout[q] && a.corr) ? sl[q].base[(size_t)gr * sl[q].ld] : 0.f;
        old[u][q] = (gr >= 0 && gout[q] && gadd[q]) ? gout[q][(size_t)gr * gld[q]] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < H; ++j) dz[u][j] = gr >= 0 ? a.dz[(size_t)r * H + j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < NR_ROWS; ++u) {
      if (grs[u] < 0) continue;
#pragma unroll
      for (int q = 0; q < NR_SLOTS; ++q) {
        if (!gout[q]) continue;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) v = fmaf(dz[u][j], w[q][j], v);
        if (a.corr) v -= k0[q] + fmaf(x[u][q], kA[q], kB[q]) * k1[q];
        gout[q][(size_t)grs[u] * gld[q]] = old[u][q] + v;
      }
    }
  }
}

static int nr_slots_of(const NarrowArgs& a) {
  int slots = 0;
  for (int p = 0; p < a.n_pieces; ++p) slots += (a.p[p].width + 31) / 32;
  return slots;
}
int narrow_supported(const NarrowArgs& a) {
  if (a.H < 1 || a.H > 4 || a.n_pieces < 1 || a.n_pieces > GEMM_MAXP) return 0;
  return nr_slots_of(a) <= NR_MAXSLOTS;
}

static int nr_grid(int n_rows) {
  const int rows_per_block = (NR_THREADS / 32) * NR_ROWS;
  int g = (n_rows + rows_per_block - 1) / rows_per_block;
  const int cap = gnnfp_num_sms() * 4;                 // <= grid_cap partial slots (loop.cu)
  if (g > cap) g = cap;
  return g < 1 ? 1 : g;
}

#define NR_DISPATCH_H(kernel, grid, NS)                                                           \
  switch (a.H) {                                                                                  \
    case 1: kernel<1, NS><<<grid, NR_THREADS, 0, s>>>(a); break;                                  \
    case 2: kernel<2, NS><<<grid, NR_THREADS, 0, s>>>(a); break;                                  \
    case 3: kernel<3, NS><<<grid, NR_THREADS, 0, s>>>(a); break;                                  \
    default: kernel<4, NS><<<grid, NR_THREADS, 0, s>>>(a); break;                                 \
  }
#define NR_DISPATCH(kernel, grid)                                                                 \
  do {                                                                                            \
    const int ns_ = nr_slots_of(a);                                                               \
    if (ns_ <= 2) { NR_DISPATCH_H(kernel, grid, 2) }                                              \
    else if (ns_ <= 3) { NR_DISPATCH_H(kernel, grid, 3) }                                         \
    else { NR_DISPATCH_H(kernel, grid, 6) }                                                       \
  } while (0)

int launch_narrow_fwd(const NarrowArgs& a, cudaStream_t s, int prof_cat) {
  if (a.n_rows <= 0) return GNNFP_OK;
  ProfScope ps(prof_cat, s);
  NR_DISPATCH(narrow_fwd_kernel, nr_grid(a.n_rows));
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
int launch_narrow_dw(const NarrowArgs& a, cudaStream_t s, int prof_cat, int* grid_out) {
  if (a.n_rows <= 0) { if (grid_out) *grid_out = 0; return GNNFP_OK; }
  int grid = nr_grid(a.n_rows);
  if (grid > 2 * gnnfp_num_sms()) grid = 2 * gnnfp_num_sms();    // fewer partial slots for the final reduction to read
  if (grid_out) *grid_out = grid;
  ProfScope ps(prof_cat, s);
  NR_DISPATCH(narrow_dw_kernel, grid);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
int launch_narrow_dx(const NarrowArgs& a, cudaStream_t s, int prof_cat) {
  if (a.n_rows <= 0) return GNNFP_OK;
  ProfScope ps(prof_cat, s);
  NR_DISPATCH(narrow_dx_kernel, nr_grid(a.n_rows));
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
