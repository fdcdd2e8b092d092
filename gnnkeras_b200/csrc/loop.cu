// loop.cu - host orchestration of the fixed-point loop behind the C ABI (gnnfp.h).
//
// Forward  = reference GNN.py:245-274 Loop (+341-346 pooling, +317-330 arc focus) and
//            CompositeGNN.py:242-272: prologue (loop-invariant aggregates), max_iteration gated
//            iteration kernels (the while_loop of GNN.py:265 without any host round trip: kernel t
//            runs only if the device flag written by kernel t-1 is set), final state, net_output,
//            NodeGraph pooling.
// Backward = the BPTT that tf.GradientTape performs in train_step (GNN.py:284-295), hand written:
//            see loop_bwd.cu.
#include <stdarg.h>
#include <stdlib.h>

#include "loop.h"
#include "tile.cuh"
#include "rows_tma.h"

// ---- error / misc -------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
long long g_gnnfp_launches = 0;
void gnnfp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* gnnfp_last_error(void) { return g_err; }
extern "C" int gnnfp_abi_version(void) { return GNNFP_ABI_VERSION; }
// ---- optional kernel timing ---------------------------------------------------------------------------
#include <vector>
int g_gnnfp_prof = 0;
struct ProfRec { int cat; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
void gnnfp_prof_begin(int cat, cudaStream_t s) {
  ProfRec r;
  r.cat = cat;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  g_prof_recs.push_back(r);
}
void gnnfp_prof_end(cudaStream_t s) { cudaEventRecord(g_prof_recs.back().b, s); }
extern "C" int gnnfp_profile_enable(int on) {
  g_gnnfp_prof = on;
  return GNNFP_OK;
}
extern "C" int gnnfp_profile_collect(double* ms_by_cat, long long* count_by_cat, int ncat) {
  for (int i = 0; i < ncat; ++i) { ms_by_cat[i] = 0.0; count_by_cat[i] = 0; }
  for (auto& r : g_prof_recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.cat >= 0 && r.cat < ncat) { ms_by_cat[r.cat] += ms; count_by_cat[r.cat] += 1; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof_recs.clear();
  return GNNFP_OK;
}

extern "C" long long gnnfp_launch_count(int reset) {
  long long v = g_gnnfp_launches;
  if (reset) g_gnnfp_launches = 0;
  return v;
}
int gnnfp_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = GNNFP_NSM_FALLBACK;
  }
  return n;
}

int net_param_count(const gnnfp_net_desc& d) {
  int n = 0, in_l = d.in_dim;
  for (int l = 0; l < d.n_layers; ++l) {
    n += in_l * d.widths[l] + d.widths[l];
    in_l = d.widths[l];
  }
  return n;
}

// ---- small kernels ------------------------------------------------------------------------------
// condition() before the first iteration: state_old = ones (GNN.py:261), k = 0 < max_iteration
static __global__ void k_cond0(const float* s0, int ld, int n, int D, float thr, int max_iter, int* flag0) {
  int notconv = 0;
  const float normp = thr * sqrtf((float)D);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float sd = 0.f;
    for (int j = 0; j < D; ++j) {
      const float d = s0[(size_t)i * ld + j] - 1.0f;
      sd = fmaf(d, d, sd);
    }
    if (sqrtf(sd) > normp) notconv = 1;
  }
  const int any = __syncthreads_or(notconv);
  if (threadIdx.x == 0 && any && max_iter > 0) atomicOr(flag0, 1);
}

__device__ __forceinline__ int count_k(const int* flags, int max_iter) {
  int k = 0;
  for (int t = 0; t < max_iter; ++t) k += flags[t] != 0;   // flags[t] set => iteration t+1 executed
  return k;
}

// state_out = S_k ; k_out = k
// slot_mode: 0 = slot k-1 (training), 1 = slot k&1 (inference), 2 = slot k (interleaved layout, training)
static __global__ void k_finalize(const int* flags, int max_iter, const float* s0, int ld0, const float* slots,
                                  size_t slot_stride, int slot_mode, int ld_slot, int n, int D, float* state_out, int* k_out) {
  const int k = count_k(flags, max_iter);
  const float* src;
  int ld;
  if (k == 0) { src = s0; ld = ld0; }
  else { src = slots + (slot_mode == 0 ? (size_t)(k - 1) : (slot_mode == 1 ? (size_t)(k & 1) : (size_t)k)) * slot_stride; ld = ld_slot; }
  // one warp per row, 4 rows in flight (a per-element division by the runtime width ran on the conversion pipe)
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = wid * 4; r0 < n; r0 += nw * 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
      if (r >= n) break;
      for (int j = lane; j < D; j += 32) state_out[(size_t)r * D + j] = src[(size_t)r * ld + j];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && k_out) *k_out = k;
}

// out[g, :] = sum_{i in graph g} NodeGraph[i, g] * out_nodes[i, :]  (GNN.py:345), sequential in node order
static __global__ void k_pool(const float* out_nodes, const int* graph_ptr, const float* ng_val, int G, int T, float* out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= G * T) return;
  const int gi = e / T, j = e - gi * T;
  float acc = 0.f;
  for (int i = graph_ptr[gi]; i < graph_ptr[gi + 1]; ++i) acc = fmaf(ng_val[i], out_nodes[(size_t)i * T + j], acc);
  out[e] = acc;
}

// interleaved layout, one warp per row (coalesced): slot 0 columns [0, D) = the caller's initial state, columns
// [2D, 2D + LsM) of the first n_slots slots = the static block; the same pass evaluates condition() before the first
// iteration (state_old = ones, GNN.py:261 - otherwise k_cond0)
static __global__ void k_xlay_init(const float* s0, int ld0, const float* Xs, int ldXs, int LsM, float* slots, size_t stride,
                                   int n_slots, int ldX, int D, int n, float thr, int max_iter, int* flag0) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float normp = thr * sqrtf((float)D);
  int notconv = 0;
  // 4 rows in flight per warp (all loads first): one row at a time left the copy latency bound (74 us for 176 MB)
  const int wid = blockIdx.x * wpb + (threadIdx.x >> 5), nw = gridDim.x * wpb;
  for (int r0 = wid * 4; r0 < n; r0 += nw * 4) {
    float v[4][3], xs[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int j = lane + 32 * i;
        v[u][i] = (r < n && j < D) ? s0[(size_t)r * ld0 + j] : 1.0f;
      }
      xs[u] = (r < n && (lane & 7) < LsM && lane < 8 * n_slots) ? Xs[(size_t)r * ldXs + (lane & 7)] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
      if (r >= n) break;
      float* dst = slots + (size_t)r * ldX;
      float sd = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int j = lane + 32 * i;
        if (j < D) { dst[j] = v[u][i]; sd = fmaf(v[u][i] - 1.0f, v[u][i] - 1.0f, sd); }
      }
      for (int j = lane + 96; j < D; j += 32) {          // state widths above 96 (not on the TMA path today)
        const float w = s0[(size_t)r * ld0 + j];
        dst[j] = w;
        sd = fmaf(w - 1.0f, w - 1.0f, sd);
      }
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, of);
      if (sqrtf(sd) > normp) notconv = 1;
      // inline static blocks have LsM <= 8 columns: lane -> (slot q = lane / 8 (+ 4 per pass), column x = lane % 8)
      const int x = lane & 7;
      if (x < LsM)
        for (int q = lane >> 3; q < n_slots; q += 4) slots[(size_t)q * stride + (size_t)r * ldX + 2 * D + x] = xs[u];
    }
  }
  const int any = __syncthreads_or(notconv);
  if (threadIdx.x == 0 && any && max_iter > 0) atomicOr(flag0, 1);
}

// ---- piece builders -----------------------------------------------------------------------------
static Piece mk_piece(int kind, const float* ptr, int ld, int width, int col0) {
  Piece p;
  memset(&p, 0, sizeof(p));
  p.kind = kind; p.ptr = ptr; p.ld = ld; p.width = width; p.col0 = col0;
  p.magic = width > 0 ? (unsigned)((0x100000000ull + (unsigned)width - 1) / (unsigned)width) : 0u;
  return p;
}
Piece mk_direct(const float* ptr, int ld, int width, int col0) { return mk_piece(PK_DIRECT, ptr, ld, width, col0); }
Piece mk_gather(const float* ptr, int ld, int width, int col0, const int* rowptr, const int* idx, const float* wgt, int nnz) {
  Piece p = mk_piece(PK_GATHER, ptr, ld, width, col0);
  p.rowptr = rowptr; p.idx = idx; p.wgt = wgt; p.nnz = nnz;
  return p;
}
void add_piece(TileSrc& ts, const Piece& p) {
  if (p.width <= 0) return;
  ts.p[ts.n_pieces++] = p;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static void set_rows(const gnnfp_loop* L, int ty, TileSrc& ts) {
  if (L->composite) { ts.n_rows = L->g->type_count[ty]; ts.rowlist = L->g->type_rows[ty]; }
  else { ts.n_rows = L->Nact; ts.rowlist = nullptr; }
}

// input of net_state[ty] at iteration t (1-based): GNN.py:222-231 / CompositeGNN.py:224
void build_state_src(const Ctx& c, int ty, int t, TileSrc& ts, int agg_direct) {
  const gnnfp_loop* L = c.L;
  const gnnfp_graph* g = L->g;
  memset(&ts, 0, sizeof(ts));
  set_rows(L, ty, ts);
  ts.in_dim = L->snet[ty].in_dim;
  const int D = L->D;
  const float* Sp = c.S(t - 1);
  const int ld = c.ldS(t - 1);
  const float* wgt = g->mode == GNNFP_AGG_SUM ? nullptr : g->dst_w;
  const bool bn = L->bn_train_state;
  const int xw = c.stXw();
  if (!L->composite) {
    const int NLp = L->S > 0 ? L->NLw : 0;
    Piece p0 = mk_direct(Sp, ld, D, 0);
    p0.tag = TAG_STATE;
    if (bn) { p0.st_sum = c.stS(0, t - 1); p0.st_sq = p0.st_sum + D; }
    add_piece(ts, p0);
    if (NLp) {
      Piece p1 = mk_direct(c.Xs(), L->ldXs, NLp, D);
      p1.tag = TAG_STATIC;
      if (bn) { p1.st_sum = c.stX(0); p1.st_sq = c.stX(0) + xw; }
      add_piece(ts, p1);
    }
    Piece p2 = agg_direct ? mk_direct(c.AGG(t), c.ldA(), D, D + NLp) : mk_gather(Sp, ld, D, D + NLp, g->dst_rowptr, g->dst_src, wgt, g->A);
    p2.tag = TAG_AGG_STATE;
    if (bn) { p2.st_sum = c.stA(0, t - 1); p2.st_sq = p2.st_sum + D; }
    add_piece(ts, p2);
    Piece p3 = mk_direct(c.Xs() + NLp, L->ldXs, NLp + L->AL, 2 * D + NLp);
    p3.tag = TAG_STATIC;
    if (bn) { p3.st_sum = c.stX(0) + NLp; p3.st_sq = c.stX(0) + xw + NLp; }
    add_piece(ts, p3);
  } else {
    const int d = L->dt[ty];
    Piece p0 = mk_direct(c.io->nodes, c.io->ld_nodes, d, 0);
    p0.tag = TAG_NODES;
    if (bn) { p0.st_sum = c.stX(ty); p0.st_sq = c.stX(ty) + xw; }
    add_piece(ts, p0);
    Piece p1 = mk_direct(Sp, ld, D, d);
    p1.tag = TAG_STATE;
    if (bn) { p1.st_sum = c.stS(ty, t - 1); p1.st_sq = p1.st_sum + D; }
    add_piece(ts, p1);
    Piece p2 = agg_direct ? mk_direct(c.AGG(t), c.ldA(), D, d + D) : mk_gather(Sp, ld, D, d + D, g->dst_rowptr, g->dst_src, wgt, g->A);
    p2.tag = TAG_AGG_STATE;
    if (bn) { p2.st_sum = c.stA(ty, t - 1); p2.st_sq = p2.st_sum + D; }
    add_piece(ts, p2);
    Piece p3 = mk_direct(c.Xs(), L->ldXs, L->sum_dt + L->AL, d + 2 * D);
    p3.tag = TAG_STATIC;
    if (bn) { p3.st_sum = c.stX(ty) + d; p3.st_sq = c.stX(ty) + xw + d; }
    add_piece(ts, p3);
  }
}

// input of net_output: apply_filters (GNN.py:239-242, 317-330; CompositeGNN.py:237-239, 315-327)
void build_out_src(const Ctx& c, TileSrc& ts) {
  const gnnfp_loop* L = c.L;
  const gnnfp_graph* g = L->g;
  memset(&ts, 0, sizeof(ts));
  ts.n_rows = L->M;
  ts.rowlist = (L->M == g->mask_len) ? nullptr : g->mask_idx;
  ts.in_dim = L->out_in;
  const int D = L->D;
  const int NLp = (!L->composite && L->S > 0) ? L->NLw : 0;
  const bool bn = L->bn_train_out;
  double* st = c.stO();
  auto with_stats = [&](Piece p) {
    if (bn) { p.st_sum = st + p.col0; p.st_sq = st + L->out_in + p.col0; }
    return p;
  };
  if (L->cfg.kind == GNNFP_KIND_ARC) {
    const int w = D + NLp;
    Piece a = mk_direct(c.io->state_out, D, D, 0); a.map = g->src; a.tag = TAG_STATE; add_piece(ts, with_stats(a));
    if (NLp) { Piece b = mk_direct(c.io->nodes, c.io->ld_nodes, NLp, D); b.map = g->src; b.tag = TAG_NODES; add_piece(ts, with_stats(b)); }
    Piece d = mk_direct(c.io->state_out, D, D, w); d.map = g->dst; d.tag = TAG_STATE; add_piece(ts, with_stats(d));
    if (NLp) { Piece e = mk_direct(c.io->nodes, c.io->ld_nodes, NLp, w + D); e.map = g->dst; e.tag = TAG_NODES; add_piece(ts, with_stats(e)); }
    { Piece al = mk_direct(c.io->arc_labels, c.io->ld_arcs, L->AL, 2 * w); al.tag = TAG_ARC_LABELS; add_piece(ts, with_stats(al)); }
  } else {
    { Piece a = mk_direct(c.io->state_out, D, D, 0); a.tag = TAG_STATE; add_piece(ts, with_stats(a)); }
    if (NLp) { Piece b = mk_direct(c.io->nodes, c.io->ld_nodes, NLp, D); b.tag = TAG_NODES; add_piece(ts, with_stats(b)); }
  }
}

void fill_netdev(const gnnfp_net_desc& d, const gnnfp_net_params& p, int training, int n_rows, NetDev& nd) {
  memset(&nd, 0, sizeof(nd));
  nd.n_layers = d.n_layers;
  nd.in_dim = d.in_dim;
  for (int l = 0; l < d.n_layers; ++l) {
    nd.widths[l] = d.widths[l];
    nd.acts[l] = d.acts[l];
    nd.W[l] = p.W[l];
    nd.b[l] = p.b[l];
  }
  nd.bn_mode = d.has_bn ? (training ? 1 : 2) : 0;
  nd.bn_eps = d.bn_eps;
  nd.bn_momentum = d.bn_momentum;
  nd.gamma = p.bn_gamma; nd.beta = p.bn_beta; nd.mmean = p.bn_moving_mean; nd.mvar = p.bn_moving_var;
  nd.inv_n = n_rows > 0 ? 1.0 / (double)n_rows : 0.0;
}

int check_params(const gnnfp_net_desc& d, const gnnfp_net_params& p, const char* what) {
  for (int l = 0; l < d.n_layers; ++l)
    if (!p.W[l] || !p.b[l]) GNNFP_FAIL(GNNFP_E_INVALID, "%s: missing Dense parameters of layer %d", what, l);
  if (d.has_bn && (!p.bn_gamma || !p.bn_beta || !p.bn_moving_mean || !p.bn_moving_var))
    GNNFP_FAIL(GNNFP_E_INVALID, "%s: missing BatchNormalization parameters", what);
  return GNNFP_OK;
}

static int check_desc(const gnnfp_net_desc& d, const char* what, bool is_out) {
  if (d.n_layers < 1 || d.n_layers > GNNFP_MAX_LAYERS) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "%s: %d Dense layers (1..%d supported)", what, d.n_layers, GNNFP_MAX_LAYERS);
  for (int l = 0; l < d.n_layers; ++l) {
    if (d.widths[l] < 1) GNNFP_FAIL(GNNFP_E_INVALID, "%s: layer %d has width %d", what, l, d.widths[l]);
    if (d.acts[l] < 0 || d.acts[l] > GNNFP_ACT_SOFTMAX) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "%s: unknown activation %d", what, d.acts[l]);
  }
  (void)is_out;
  return GNNFP_OK;
}

// ---- plan ----------------------------------------------------------------------------------------
extern "C" int gnnfp_loop_create(gnnfp_loop** out, const gnnfp_graph* g, const gnnfp_loop_cfg* cfg,
                                 const gnnfp_net_desc* state_nets, const gnnfp_net_desc* out_net) {
  if (!out || !g || !cfg || !state_nets || !out_net) GNNFP_FAIL(GNNFP_E_INVALID, "loop_create: null argument");
  *out = nullptr;
  // the reference's constructor asserts (GNN.py:26-28, CompositeGNN.py:26-28)
  if (cfg->state_vect_dim < 0) GNNFP_FAIL(GNNFP_E_INVALID, "assert state_vect_dim >= 0");
  if (cfg->max_iteration < 0) GNNFP_FAIL(GNNFP_E_INVALID, "assert max_iteration >= 0");
  if (cfg->state_threshold < 0) GNNFP_FAIL(GNNFP_E_INVALID, "assert state_threshold >= 0");
  if (cfg->n_types != 0 && cfg->max_iteration == 0) GNNFP_FAIL(GNNFP_E_INVALID, "assert max_iteration > 0 (composite)");
  if (cfg->n_types != g->n_types) GNNFP_FAIL(GNNFP_E_INVALID, "loop_create: cfg.n_types=%d but the graph has %d node types", cfg->n_types, g->n_types);
  if (cfg->kind < 0 || cfg->kind > GNNFP_KIND_GRAPH) GNNFP_FAIL(GNNFP_E_INVALID, "loop_create: kind=%d", cfg->kind);
  gnnfp_loop* L = new gnnfp_loop();
  L->g = g; L->cfg = *cfg;
  L->composite = cfg->n_types > 0;
  L->nt = L->composite ? cfg->n_types : 1;
  L->N = g->N; L->A = g->A; L->M = g->M;
  L->S = cfg->state_vect_dim; L->NLw = cfg->nodes_width; L->AL = cfg->arc_label_width;
  L->D = L->S > 0 ? L->S : L->NLw;
  L->Nact = (cfg->n_active_rows > 0 && cfg->n_active_rows < g->N) ? cfg->n_active_rows : g->N;
  int rc = GNNFP_OK;
#define PLAN_FAIL(code, ...) do { gnnfp_set_error(__VA_ARGS__); delete L; return (code); } while (0)
  if (L->NLw <= 0) PLAN_FAIL(GNNFP_E_INVALID, "loop_create: nodes_width=%d", L->NLw);
  if (L->composite && !g->types_ok)
    PLAN_FAIL(GNNFP_E_UNSUPPORTED, "composite: every node must belong to exactly one type (one-hot type_mask)");
  for (int t = 0; t < L->nt; ++t) {
    L->snet[t] = state_nets[t];
    if ((rc = check_desc(L->snet[t], "net_state", false))) { delete L; return rc; }
  }
  L->onet = *out_net;
  if ((rc = check_desc(L->onet, "net_output", true))) { delete L; return rc; }
  L->T = L->onet.widths[L->onet.n_layers - 1];
  // widths (MLP.py:112-124 get_inout_dims)
  const int NLp = (!L->composite && L->S > 0) ? L->NLw : 0;
  if (!L->composite) {
    L->LsM = 2 * NLp + L->AL;
    const int din = 2 * L->D + L->LsM;
    if (L->snet[0].in_dim != din) PLAN_FAIL(GNNFP_E_INVALID, "net_state expects %d inputs but the loop feeds %d ([state|nodes?|agg_state|agg_nodes|agg_arcs])", L->snet[0].in_dim, din);
  } else {
    L->sum_dt = 0;
    for (int t = 0; t < L->nt; ++t) {
      int d = cfg->dim_node_label[t];
      if (d < 0) PLAN_FAIL(GNNFP_E_INVALID, "dim_node_label[%d]=%d", t, d);
      L->dt[t] = d > L->NLw ? L->NLw : d;   // python slicing nodes[:, :d] clamps
      L->sum_dt += L->dt[t];
    }
    L->LsM = L->sum_dt + L->AL;
    for (int t = 0; t < L->nt; ++t) {
      const int din = L->dt[t] + 2 * L->D + L->LsM;
      if (L->snet[t].in_dim != din) PLAN_FAIL(GNNFP_E_INVALID, "net_state[%d] expects %d inputs but the loop feeds %d", t, L->snet[t].in_dim, din);
    }
  }
  for (int t = 0; t < L->nt; ++t)
    if (L->snet[t].widths[L->snet[t].n_layers - 1] != L->D)
      PLAN_FAIL(GNNFP_E_INVALID, "net_state[%d] outputs %d columns but the state has %d", t, L->snet[t].widths[L->snet[t].n_layers - 1], L->D);
  if (cfg->kind == GNNFP_KIND_ARC) {
    if (g->mask_len != g->A) PLAN_FAIL(GNNFP_E_INVALID, "arc focus: masks must have one entry per arc");
    L->out_in = 2 * (L->D + NLp) + L->AL;
  } else {
    if (g->mask_len != g->N) PLAN_FAIL(GNNFP_E_INVALID, "node/graph focus: masks must have one entry per node");
    L->out_in = L->D + NLp;
  }
  if (L->onet.in_dim != L->out_in) PLAN_FAIL(GNNFP_E_INVALID, "net_output expects %d inputs but apply_filters yields %d", L->onet.in_dim, L->out_in);
  L->pool = cfg->pool < 0 ? (cfg->kind == GNNFP_KIND_GRAPH) : (cfg->pool != 0);
  if (L->pool) {
    if (g->G <= 0) PLAN_FAIL(GNNFP_E_INVALID, "graph focus needs a NodeGraph (n_graphs > 0)");
    if (L->M != L->N) PLAN_FAIL(GNNFP_E_INVALID, "graph focus: NodeGraph pooling needs every node unmasked (M=%d, N=%d)", L->M, L->N);
  }
  L->out_rows = L->pool ? g->G : L->M;
  L->bn_train_state = cfg->training;
  for (int t = 0; t < L->nt; ++t) L->bn_train_state = L->bn_train_state && L->snet[t].has_bn;
  for (int t = 0; t < L->nt; ++t)
    if (L->snet[t].has_bn != L->snet[0].has_bn) PLAN_FAIL(GNNFP_E_UNSUPPORTED, "composite: all state nets must agree on BatchNormalization");
  L->bn_train_out = cfg->training && L->onet.has_bn;
  if (L->Nact < L->N) {
    if (L->composite) PLAN_FAIL(GNNFP_E_UNSUPPORTED, "partitioned (n_active_rows) loops are homogeneous only");
    if (cfg->training && L->onet.has_bn) PLAN_FAIL(GNNFP_E_UNSUPPORTED, "partitioned (n_active_rows) training does not support BatchNormalization in net_output (global batch statistics)");
    if (cfg->training && cfg->want_input_grads) PLAN_FAIL(GNNFP_E_UNSUPPORTED, "partitioned (n_active_rows) training does not produce input gradients");
    if (cfg->training && L->snet[0].n_layers != 1) PLAN_FAIL(GNNFP_E_UNSUPPORTED, "partitioned (n_active_rows) training supports single-Dense-layer net_state only");
    if (L->snet[0].has_bn) PLAN_FAIL(GNNFP_E_UNSUPPORTED, "partitioned (n_active_rows) loops do not support BatchNormalization in net_state (global batch statistics)");
  }
  for (int t = 0; t < L->nt; ++t) L->nparam_s[t] = net_param_count(L->snet[t]);
  L->nparam_o = net_param_count(L->onet);
  { long long cpr = g->N > 0 ? (2ll * g->A + g->N - 1) / g->N : 4; L->cap_per_row = cpr < 4 ? 4 : (cpr > 16 ? 16 : (int)cpr); }
  for (int t = 0; t < L->nt; ++t) {
    const gnnfp_net_desc& d = L->snet[t];
    L->gemm_ok[t] = d.n_layers == 1 && d.acts[0] != GNNFP_ACT_SOFTMAX && d.widths[0] <= 80 && gemm_rows_supported(ceil_to(d.in_dim, 8) + 8 * 3, d.widths[0]) &&
                    getenv("GNNFP_NO_GEMM") == nullptr;
  }
  {
    const gnnfp_net_desc& d = L->onet;
    L->out_gemm_ok = cfg->kind != GNNFP_KIND_ARC && d.n_layers == 1 && d.widths[0] <= (d.acts[0] == GNNFP_ACT_SOFTMAX ? 16 : 80) &&
                     gemm_rows_supported(ceil_to(d.in_dim, 8) + 8 * 3, d.widths[0]) && getenv("GNNFP_NO_GEMM") == nullptr;
  }
  L->grid_cap = gnnfp_num_sms() * 4;   // backward tile kernels run at most 4 CTAs per SM (tile_cfg_bwd)
  // ---- TMA path (rows_tma.cu): homogeneous single-Dense-layer state net on an unpartitioned row set.  Chosen here by
  // ---- shape, never as a fallback after a failure; GNNFP_NO_RT=1 keeps the cp.async kernels (debug / comparison) ------
  L->ldG = L->D; L->ldXs = L->LsM; L->ldX = L->D;
  {
    static const int no_rt = getenv("GNNFP_NO_RT") ? 1 : 0;
    // (partitioned plans - n_active_rows < N, the rest are halo copies - take it too: no BatchNormalization there, and no
    //  fused aggregation, whose tile-local CSR view assumes that every row of a tile is computed by this launch)
    const bool part = L->Nact < L->N;
    const bool part_ok = !part || !(L->snet[0].has_bn || L->onet.has_bn);
    if (!no_rt && !L->composite && L->gemm_ok[0] && part_ok && cfg->max_iteration > 0 && L->D <= 96 && rows_tma_available()) {
      const int inl = L->LsM > 0 && L->LsM <= 8;
      const int w0 = 2 * L->D + (inl ? L->LsM : 0);
      const int nkc = (w0 + 31) / 32 + (inl ? 0 : (L->LsM + 31) / 32);
      RowsTmaArgs pf, pd;
      memset(&pf, 0, sizeof(pf));
      memset(&pd, 0, sizeof(pd));
      // fused aggregation (rows_tma.cu): measured a net loss when the aggregate's BN batch statistics have to be reduced in
      // the same kernel (column sums across thread = row partials), a gain otherwise; GNNFP_FUSE_AGG=0|1 overrides
      const bool bn_stats = cfg->training && L->snet[0].has_bn;
      const char* fenv = getenv("GNNFP_FUSE_AGG");
      const int no_fuse = part ? 1 : (fenv ? (atoi(fenv) == 0) : (bn_stats ? 1 : 0));
      pf.mode = RT_FWD; pf.n_kc = nkc; pf.n_oc = (L->D + 31) / 32; pf.BN = ceil_to(L->D, 16); pf.fuse_agg = !no_fuse;
      pd.mode = RT_DX; pd.n_kc = (L->D + 31) / 32; pd.n_oc = 2 * ((L->D + 31) / 32); pd.BN = 2 * ceil_to(L->D, 16);
      if (nkc <= RT_MAXKC && rows_tma_finish(pf) == GNNFP_OK && (!cfg->training || rows_tma_finish(pd) == GNNFP_OK)) {
        L->xlay = 1; L->xs_inline = inl; L->fuse_agg = !no_fuse;
        L->ldX = ceil_to(w0, 4); L->ldG = ceil_to(L->D, 4); L->ldXs = ceil_to(L->LsM, 4);
      }
    }
  }

  // ---- workspace layout ---------------------------------------------------------------------
  WsLayout& w = L->ws;
  const int MI = cfg->max_iteration;
  size_t off = 0;
  w.ctrl = off;
  w.flags = off; off = align_up(off + sizeof(int) * (MI + 4));
  const size_t st_per = (size_t)L->nt * (MI + 1) * 2 * L->D * sizeof(double);
  w.stS = off; off = align_up(off + st_per);
  w.stA = off; off = align_up(off + st_per);
  int xw = L->LsM;
  if (L->composite) { int m = 0; for (int t = 0; t < L->nt; ++t) m = L->dt[t] > m ? L->dt[t] : m; xw = m + L->sum_dt + L->AL; }
  w.stX = off; off = align_up(off + (size_t)L->nt * 2 * (xw > 0 ? xw : 1) * sizeof(double));
  w.stO = off; off = align_up(off + (size_t)2 * L->out_in * sizeof(double));
  w.ctrl_bytes = off - w.ctrl;
  w.Xs = off; off = align_up(off + (size_t)L->N * (L->ldXs > 0 ? L->ldXs : 1) * sizeof(float));
  const size_t slot_floats = ((size_t)L->N * (L->xlay ? L->ldX : L->D) + 31) / 32 * 32;
  if (L->xlay) L->slot_count = cfg->training ? MI + 1 : 2;        // X_0 .. X_MI (slot 0 = copy of the initial state)
  else L->slot_count = cfg->training ? MI : (MI > 0 ? 2 : 0);
  w.slots = off; off = align_up(off + (size_t)L->slot_count * slot_floats * sizeof(float));
  w.out_nodes = off; off = align_up(off + (size_t)L->M * L->T * sizeof(float));
  w.agg = off; off = align_up(off + (L->xlay ? (size_t)0 : (cfg->training ? (size_t)MI : (size_t)(MI > 0 ? 1 : 0))) * slot_floats * sizeof(float) + 4);
  {
    size_t wf = 0, wt = 0, bc = 0;
    for (int t = 0; t <= L->nt; ++t) {                 // slot nt: net_output
      const gnnfp_net_desc& dd = t < L->nt ? L->snet[t] : L->onet;
      const int in = dd.in_dim, H = dd.widths[0];
      const size_t kp = (size_t)gemm_rows_kpad(ceil_to(in, 8) + 8 * GNNFP_MAXP);
      const size_t f = kp * gemm_rows_ldw(H) + gemm_rows_ldw(H);
      const size_t tb = (size_t)ceil_to(H, 8) * ((size_t)ceil_to(in, 16) + 16 * GNNFP_MAXP);
      wf = f > wf ? f : wf; wt = tb > wt ? tb : wt; bc = (size_t)3 * in > bc ? (size_t)3 * in : bc;
    }
    w.wfold_stride = (wf + 63) / 64 * 64; w.wtb_stride = (wt + 63) / 64 * 64; w.bncoef_stride = (bc + 63) / 64 * 64;
    w.wfold = off; off = align_up(off + (L->nt + 1) * w.wfold_stride * sizeof(float));
    w.wtb = off; off = align_up(off + (L->nt + 1) * w.wtb_stride * sizeof(float));
    w.bncoef = off; off = align_up(off + (L->nt + 1) * w.bncoef_stride * sizeof(float));
    w.bncoef_t = off; off = align_up(off + (size_t)(L->xlay ? MI : 0) * w.bncoef_stride * sizeof(float));
  }
  if (cfg->training) {
    const size_t ND = (((size_t)L->N * L->ldG + 31) / 32 * 32) * sizeof(float);
    w.dSfin = off; off = align_up(off + ND);
    w.dOwn = off; off = align_up(off + 2 * ND);
    w.dAgg = off; off = align_up(off + 2 * ND);
    w.dz = off; off = align_up(off + ND);
    if (L->Nact < L->N) { w.pgather = off; off = align_up(off + ND); }
    w.dOutN = off; off = align_up(off + (size_t)L->M * L->T * sizeof(float));
    if (cfg->kind == GNNFP_KIND_ARC)
      { w.arc_tmp = off; off = align_up(off + (size_t)L->A * 2 * (L->D + ((!L->composite && L->S > 0) ? L->NLw : 0)) * sizeof(float) + 16); }
    size_t ps = 0, bg = 2 * (size_t)L->onet.in_dim;
    int din_max = L->onet.in_dim;
    for (int t = 0; t < L->nt; ++t) {
      ps += (size_t)L->nparam_s[t];
      bg += 2 * (size_t)L->snet[t].in_dim;
      din_max = L->snet[t].in_dim > din_max ? L->snet[t].in_dim : din_max;
    }
    w.bn_part = off; off = align_up(off + (size_t)L->grid_cap * 2 * din_max * sizeof(float));
    w.bn_const = off; off = align_up(off + (size_t)4 * din_max * sizeof(float));
    w.bn_const_t = off; off = align_up(off + (size_t)(MI + 2) * 4 * din_max * sizeof(float));
    w.bwd_zero = off;
    w.dXs = off; off = align_up(off + (size_t)L->N * (L->ldXs > 0 ? L->ldXs : 1) * sizeof(float) * (cfg->want_input_grads ? 1 : 0) + 4);
    w.part_state = off; off = align_up(off + (size_t)L->grid_cap * ps * sizeof(float));
    w.part_out = off; off = align_up(off + (size_t)L->grid_cap * L->nparam_o * sizeof(float));
    w.bn_grad = off; off = align_up(off + bg * sizeof(float));
    w.bn_static = off; off = align_up(off + (size_t)2 * din_max * sizeof(float));
    w.bwd_zero_bytes = off - w.bwd_zero;
  }
  w.total = off;
  *out = L;
  return GNNFP_OK;
}

extern "C" void gnnfp_loop_free(gnnfp_loop* L) { delete L; }
extern "C" size_t gnnfp_loop_workspace_bytes(const gnnfp_loop* L) { return L ? L->ws.total : 0; }
extern "C" int gnnfp_loop_out_rows(const gnnfp_loop* L) { return L ? L->out_rows : -1; }
extern "C" int gnnfp_loop_state_dim(const gnnfp_loop* L) { return L ? L->D : -1; }

int check_io(const gnnfp_loop* L, const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes) {
  if (!L || !io) GNNFP_FAIL(GNNFP_E_INVALID, "loop: null argument");
  if (!workspace || workspace_bytes < L->ws.total) GNNFP_FAIL(GNNFP_E_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, L->ws.total);
  if (((uintptr_t)workspace & 255) != 0) GNNFP_FAIL(GNNFP_E_INVALID, "workspace must be 256-byte aligned");
  if (!io->nodes || io->ld_nodes < L->NLw) GNNFP_FAIL(GNNFP_E_INVALID, "loop: nodes missing or ld_nodes < nodes_width");
  if (L->AL > 0 && L->A > 0 && (!io->arc_labels || io->ld_arcs < L->AL)) GNNFP_FAIL(GNNFP_E_INVALID, "loop: arc_labels missing or ld_arcs < AL");
  if (L->S > 0 && !io->state0) GNNFP_FAIL(GNNFP_E_INVALID, "loop: state0 must be passed explicitly when state_vect_dim > 0 (GNN.py:257 draws it unseeded)");
  if (!io->state_out || !io->out) GNNFP_FAIL(GNNFP_E_INVALID, "loop: state_out / out missing");
  return GNNFP_OK;
}

// ---- forward ---------------------------------------------------------------------------------------
static int fwd_begin(const Ctx& c, const gnnfp_net_params* sp, const gnnfp_net_params* op) {
  cudaStream_t s = c.s;
  gnnfp_loop* L = c.L;
  const gnnfp_loop_io* io = c.io;
  const gnnfp_graph* g = L->g;
  const int MI = L->cfg.max_iteration, D = L->D, N = L->N;
  const int training = L->cfg.training;
  const float* wgt = g->mode == GNNFP_AGG_SUM ? nullptr : g->dst_w;
  int rc = GNNFP_OK;
  (void)MI; (void)D; (void)N; (void)training; (void)wgt; (void)io; (void)op; (void)sp;
  GNNFP_CHECK_CUDA(cudaMemsetAsync(c.ws + L->ws.ctrl, 0, L->ws.ctrl_bytes, s));

  // ---- prologue: loop-invariant aggregates (GNN.py:254-258, CompositeGNN.py:251-253) -----------
  if (L->LsM > 0) {
    PassArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.src.n_rows = N; pa.src.in_dim = L->LsM;
    if (!L->composite) {
      const int NLp = L->S > 0 ? L->NLw : 0;
      if (NLp) {
        add_piece(pa.src, mk_direct(io->nodes, io->ld_nodes, NLp, 0));
        add_piece(pa.src, mk_gather(io->nodes, io->ld_nodes, NLp, NLp, g->dst_rowptr, g->dst_src, wgt, g->A));
      }
      add_piece(pa.src, mk_gather(io->arc_labels, io->ld_arcs, L->AL, 2 * NLp, g->dst_rowptr, g->dst_arc, wgt, g->A));
      if (L->bn_train_state) { pa.st_sum = c.stX(0); pa.st_sq = c.stX(0) + c.stXw(); }
    } else {
      int col = 0;
      for (int t = 0; t < L->nt; ++t) {
        add_piece(pa.src, mk_gather(io->nodes, io->ld_nodes, L->dt[t], col, g->dst_rowptr, g->dst_src, g->typed_w[t], g->A));
        col += L->dt[t];
      }
      add_piece(pa.src, mk_gather(io->arc_labels, io->ld_arcs, L->AL, col, g->dst_rowptr, g->dst_arc, wgt, g->A));
    }
    pa.out = c.Xs(); pa.ld_out = L->ldXs;
    pa.tc.cap_per_row = L->cap_per_row;
    if ((rc = tile_cfg_pass(L->LsM, N, &pa.tc))) return rc;
    if ((rc = launch_tile_pass(pa, s))) return rc;
  }
  if (L->composite && L->bn_train_state) {   // per-type statistics of the static columns
    for (int ty = 0; ty < L->nt; ++ty) {
      PassArgs pa;
      memset(&pa, 0, sizeof(pa));
      set_rows(L, ty, pa.src);
      pa.src.in_dim = L->dt[ty] + L->LsM;
      add_piece(pa.src, mk_direct(io->nodes, io->ld_nodes, L->dt[ty], 0));
      add_piece(pa.src, mk_direct(c.Xs(), L->ldXs, L->LsM, L->dt[ty]));
      // stX(ty) is laid out [d_t | LsM] with stride stXw: the pass writes sums to [0,in_dim) and squares
      // to [stXw, stXw+in_dim)
      pa.st_sum = c.stX(ty); pa.st_sq = c.stX(ty) + c.stXw();
      pa.tc.cap_per_row = L->cap_per_row;
      if ((rc = tile_cfg_pass(pa.src.in_dim, pa.src.n_rows, &pa.tc))) return rc;
      if ((rc = launch_tile_pass(pa, s))) return rc;
    }
  }
  if (L->xlay) {   // X_0[:, 0:D] = the caller's initial state; inline static columns into every slot
    const int ns = L->xs_inline ? L->slot_count : 0;
    int blocks = (N + 31) / 32;                          // 8 warps x 4 rows per block and trip
    if (blocks > 2368) blocks = 2368;
    k_xlay_init<<<blocks, 256, 0, s>>>(c.S0user(), c.ldS0user(), c.Xs(), L->ldXs, L->LsM, c.slots(), c.slot_stride(), ns, L->ldX, D, N,
                                       L->cfg.state_threshold, MI, c.flags());
    GNNFP_COUNT_LAUNCH();
  }
  // ---- condition before the first iteration ------------------------------------------------------
  if (!L->xlay) {
    const int blocks = (L->Nact + 255) / 256 < 1184 ? (L->Nact + 255) / 256 : 1184;
    k_cond0<<<blocks, 256, 0, s>>>(c.S(0), c.ldS(0), L->Nact, D, L->cfg.state_threshold, MI, c.flags());
    GNNFP_COUNT_LAUNCH();
  }
  if (L->bn_train_state && MI > 0) {   // statistics of S_0 per type
    for (int ty = 0; ty < L->nt; ++ty) {
      PassArgs pa;
      memset(&pa, 0, sizeof(pa));
      set_rows(L, ty, pa.src);
      pa.src.in_dim = D;
      add_piece(pa.src, mk_direct(c.S(0), c.ldS(0), D, 0));
      pa.st_sum = c.stS(ty, 0); pa.st_sq = pa.st_sum + D;
      pa.tc.cap_per_row = L->cap_per_row;
      if ((rc = tile_cfg_pass(D, pa.src.n_rows, &pa.tc))) return rc;
      if ((rc = launch_tile_pass(pa, s))) return rc;
    }
  }
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return rc;
}

// argument block of the TMA forward iteration t: operand chunks over X_{t-1} = [S_{t-1} | Adj^T S_{t-1} | static?] (+ the
// static block Xs when it is not inline), output chunks of S_t, the previous state as the side input of the convergence test
static int rt_build_fwd(const Ctx& c, int t, const gnnfp_net_params* sp, RowsTmaArgs& ra) {
  gnnfp_loop* L = c.L;
  const int MI = L->cfg.max_iteration, D = L->D, N = L->N, LsM = L->LsM;
  const int NLp = L->S > 0 ? L->NLw : 0;
  int rc;
  memset(&ra, 0, sizeof(ra));
  ra.mode = RT_FWD; ra.n_rows = L->Nact; ra.H = D;
  build_state_src(c, 0, t, ra.src, 1);
  fill_netdev(L->snet[0], sp[0], L->cfg.training, ra.src.n_rows, ra.net);
  ra.update_moving = L->cfg.training; ra.act = L->snet[0].acts[0];
  // training with batch statistics: CTA 0 leaves the backward's coefficients [rstd | -mean rstd | gamma rstd] of this iteration
  if (L->cfg.training && L->snet[0].has_bn && L->ws.bncoef_stride)
    ra.coef_out = (float*)(c.ws + L->ws.bncoef_t) + (size_t)(t - 1) * L->ws.bncoef_stride;
  const int w0 = 2 * D + (L->xs_inline ? LsM : 0);
  if ((rc = rows_tma_map(&ra.maps[0], c.S(t - 1), N, w0, L->ldX))) return rc;
  if (!L->xs_inline && LsM > 0 && (rc = rows_tma_map(&ra.maps[1], c.Xs(), N, LsM, L->ldXs))) return rc;
  if ((rc = rows_tma_map(&ra.maps[2], c.S(t - 1), N, D, L->ldX))) return rc;
  if ((rc = rows_tma_map(&ra.maps[3], c.S(t), L->Nact, D, L->ldX))) return rc;   // stores clip at the computed rows (halo rows below stay the exchange's)
  // input column (row of W) of X-slot column m / static column x: the net sees [S | nodes? | Adj^T S | agg_nodes | agg_arcs]
  auto stat_col = [&](int x) { return x < NLp ? D + x : 2 * D + x; };
  auto slot_col = [&](int m) { return m < D ? m : (m < 2 * D ? D + NLp + (m - D) : stat_col(m - 2 * D)); };
  for (int c0 = 0; c0 < w0; c0 += RT_CHUNK) {
    RtKChunk& k = ra.kc[ra.n_kc++];
    k.map = 0; k.col0 = c0; k.width = w0 - c0 < RT_CHUNK ? w0 - c0 : RT_CHUNK; k.k8 = (k.width + 7) / 8;
    for (int j = 0; j < RT_CHUNK; ++j) k.wrow[j] = (short)(j < k.width ? slot_col(c0 + j) : -1);
  }
  if (!L->xs_inline)
    for (int c0 = 0; c0 < LsM; c0 += RT_CHUNK) {
      RtKChunk& k = ra.kc[ra.n_kc++];
      k.map = 1; k.col0 = c0; k.width = LsM - c0 < RT_CHUNK ? LsM - c0 : RT_CHUNK; k.k8 = (k.width + 7) / 8;
      for (int j = 0; j < RT_CHUNK; ++j) k.wrow[j] = (short)(j < k.width ? stat_col(c0 + j) : -1);
    }
  for (int c0 = 0; c0 < D; c0 += RT_CHUNK) {
    RtOChunk& o = ra.oc[ra.n_oc];
    o.acc_col0 = c0; o.out_map = 3; o.out_col0 = c0;
    o.aux_map = t < MI ? 2 : -1; o.aux_col0 = c0; o.cidx0 = c0;
    o.width = D - c0 < RT_CHUNK ? D - c0 : RT_CHUNK; o.st_slot = ra.n_oc;
    ++ra.n_oc;
  }
  ra.BN = ceil_to(D, 16);
  if (t < MI) {
    ra.thr = L->cfg.state_threshold; ra.flag_next = c.flags() + t;
    if (L->bn_train_state) { ra.ost_sum = c.stS(0, t); ra.ost_sq = ra.ost_sum + D; }
  }
  ra.gate = c.flags() + (t - 1);
  if (L->fuse_agg && t < MI) {   // Adj^T S_t for iteration t + 1, gathered from the output stages (rows local to their tile)
    const gnnfp_graph* g = L->g;
    ra.fuse_agg = 1;
    ra.g_rowptr = g->dst_rowptr; ra.g_lidx = g->tile_lidx; ra.g_arc0 = g->tile_arc0;
    ra.g_w = g->mode == GNNFP_AGG_SUM ? nullptr : g->dst_w;
    ra.agg_out = c.AGG(t + 1); ra.ld_agg = c.ldA();
    if (L->bn_train_state) { ra.agg_sum = c.stA(0, t); ra.agg_sq = ra.agg_sum + D; }
  }
  return rows_tma_finish(ra);
}

// iteration t (1-based), gated on the device flag written by iteration t-1 (GNN.py:265)
static int fwd_iter(const Ctx& c, int t, const gnnfp_net_params* sp, const gnnfp_net_params* op) {
  cudaStream_t s = c.s;
  gnnfp_loop* L = c.L;
  const gnnfp_loop_io* io = c.io;
  const gnnfp_graph* g = L->g;
  const int MI = L->cfg.max_iteration, D = L->D, N = L->N;
  const int training = L->cfg.training;
  const float* wgt = g->mode == GNNFP_AGG_SUM ? nullptr : g->dst_w;
  int rc = GNNFP_OK;
  (void)MI; (void)D; (void)N; (void)training; (void)wgt; (void)io; (void)op; (void)sp;
  {
    const int* gate = c.flags() + (t - 1);
    for (int ty = 0; ty < L->nt; ++ty) {     // Adj^T.state of this iteration (+ BN batch statistics); saved in training
      if (!L->bn_train_state && !L->gemm_ok[ty]) continue;
      if (L->fuse_agg && t > 1) continue;    // produced by iteration t - 1 (fused gather + the row-list pass below)
      AggArgs aa;
      memset(&aa, 0, sizeof(aa));
      TileSrc rows;
      memset(&rows, 0, sizeof(rows));
      set_rows(L, ty, rows);
      aa.n_rows = rows.n_rows; aa.rowlist = rows.rowlist; aa.D = D;
      aa.S = c.S(t - 1); aa.ld = c.ldS(t - 1);
      aa.rowptr = g->dst_rowptr; aa.idx = g->dst_src; aa.wgt = wgt;
      aa.out = c.AGG(t); aa.ld_out = c.ldA();
      if (L->bn_train_state) { aa.st_sum = c.stA(ty, t - 1); aa.st_sq = aa.st_sum + D; }
      aa.gate = gate;
      if (L->xlay && rows.rowlist == nullptr && agg_tile_supported(aa.S, aa.ld, D)) {
        if ((rc = launch_agg_tile(aa.S, aa.ld, aa.n_rows, D, g->dst_rowptr, g->dst_src, wgt, g->tile_lidx, g->tile_arc0, aa.out, aa.ld_out,
                                  aa.st_sum, aa.st_sq, gate, s, PC_AGG))) return rc;
        continue;
      }
      if ((rc = launch_agg_stats(aa, s))) return rc;
    }
    if (L->xlay) {
      RowsTmaArgs ra;
      if ((rc = rt_build_fwd(c, t, sp, ra))) return rc;
      if ((rc = launch_rows_tma(ra, s, PC_FWD_ITER))) return rc;
      if (L->fuse_agg && t < MI) {
        // rows with an in-neighbour outside their 128-row tile (graphs straddling a tile boundary): the streaming kernel over
        // the graph's boundary row list (its length stays on the device), only if iteration t + 1 runs
        AggArgs aa;
        memset(&aa, 0, sizeof(aa));
        aa.n_rows = N; aa.n_rows_dev = g->bnd_count; aa.rowlist = g->bnd_rows; aa.D = D;
        aa.S = c.S(t); aa.ld = c.ldS(t);
        aa.rowptr = g->dst_rowptr; aa.idx = g->dst_src; aa.wgt = wgt;
        aa.out = c.AGG(t + 1); aa.ld_out = c.ldA();
        if (L->bn_train_state) { aa.st_sum = c.stA(0, t); aa.st_sq = aa.st_sum + D; }
        aa.gate = c.flags() + t;
        if ((rc = launch_agg_stats(aa, s))) return rc;
      }
    }
    for (int ty = 0; ty < L->nt; ++ty) {
      if (!L->gemm_ok[ty] || L->xlay) continue;
      // ---- pipelined GEMM path: fold BN into the padded weights, then one GEMM with the iteration's epilogue ----
      TileSrc ts;
      build_state_src(c, ty, t, ts, 1);
      FoldArgs fo;
      memset(&fo, 0, sizeof(fo));
      fo.src = ts;
      fill_netdev(L->snet[ty], sp[ty], training, ts.n_rows, fo.net);
      GemmRowsArgs ga;
      memset(&ga, 0, sizeof(ga));
      int k8 = 0;
      for (int p = 0; p < ts.n_pieces; ++p) {
        fo.k8[p] = k8;
        gemm_piece_set(ga.p[p], ts.p[p].ptr, ts.p[p].ld, ts.p[p].width, k8);
        k8 += ceil_to(ts.p[p].width, 2);
      }
      const int H = L->snet[ty].widths[0];
      float* wf = (float*)(c.ws + L->ws.wfold) + (size_t)ty * L->ws.wfold_stride;
      fo.Kpad = gemm_rows_kpad(k8); fo.ldw = gemm_rows_ldw(H); fo.Wp = wf; fo.biasp = wf + (size_t)fo.Kpad * fo.ldw;
      fo.update_moving = training; fo.gate = gate;
      if ((rc = launch_fold_w(fo, s))) return rc;
      ga.n_rows = ts.n_rows; ga.rowlist = ts.rowlist; ga.n_pieces = ts.n_pieces; ga.Kpad = fo.Kpad;
      ga.Wp = fo.Wp; ga.ldw = fo.ldw; ga.N = H; ga.bias = fo.biasp; ga.act = L->snet[ty].acts[0];
      ga.out = (float*)c.S(t); ga.ld_out = c.ldS(t); ga.fwd = 1;
      ga.vec2 = D % 2 == 0 && c.ldS(t - 1) % 2 == 0 && ((uintptr_t)c.S(t) & 7) == 0 && ((uintptr_t)c.S(t - 1) & 7) == 0;
      if (t < MI) {
        ga.prev = c.S(t - 1); ga.ld_prev = c.ldS(t - 1); ga.thr = L->cfg.state_threshold; ga.flag_next = c.flags() + t;
        if (L->bn_train_state) { ga.ost_sum = c.stS(ty, t); ga.ost_sq = ga.ost_sum + D; }
      }
      ga.gate = gate;
      if ((rc = launch_gemm_rows(ga, s, PC_FWD_ITER))) return rc;
    }
    for (int ty = 0; ty < L->nt; ++ty) {
      if (L->gemm_ok[ty]) continue;
      FwdArgs fa;
      memset(&fa, 0, sizeof(fa));
      build_state_src(c, ty, t, fa.src, L->bn_train_state ? 1 : 0);
      if (training && !L->bn_train_state) {          // no BN pass: the iteration kernel itself saves Adj^T.state
        fa.agg_out = c.AGG(t);
        fa.agg_col0 = (L->composite ? L->dt[ty] : 0) + D + ((!L->composite && L->S > 0) ? L->NLw : 0);
        fa.agg_w = D;
      }
      fill_netdev(L->snet[ty], sp[ty], training, fa.src.n_rows, fa.net);
      fa.tc.cap_per_row = L->cap_per_row;
      if ((rc = tile_cfg_fwd(fa.net, fa.src.n_rows, &fa.tc))) return rc;
      fa.prev_col0 = L->composite ? L->dt[ty] : 0;
      fa.out = (float*)c.S(t); fa.ld_out = c.ldS(t); fa.out_compact = 0;
      if (L->bn_train_state && t < MI) { fa.ost_sum = c.stS(ty, t); fa.ost_sq = fa.ost_sum + D; }
      if (t < MI) { fa.prev = c.S(t - 1); fa.ld_prev = c.ldS(t - 1); fa.thr = L->cfg.state_threshold; fa.flag_next = c.flags() + t; }
      fa.gate = gate;
      fa.update_moving = training;
      fa.prof_cat = PC_FWD_ITER;
      if ((rc = launch_tile_fwd(fa, s))) return rc;
    }
  }
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return rc;
}

static int fwd_end(const Ctx& c, const gnnfp_net_params* sp, const gnnfp_net_params* op) {
  cudaStream_t s = c.s;
  gnnfp_loop* L = c.L;
  const gnnfp_loop_io* io = c.io;
  const gnnfp_graph* g = L->g;
  const int MI = L->cfg.max_iteration, D = L->D, N = L->N;
  const int training = L->cfg.training;
  const float* wgt = g->mode == GNNFP_AGG_SUM ? nullptr : g->dst_w;
  int rc = GNNFP_OK;
  (void)MI; (void)D; (void)N; (void)training; (void)wgt; (void)io; (void)op; (void)sp;
  // ---- converged state, iteration count --------------------------------------------------------------
  {
    int blocks = (N + 31) / 32;                          // 8 warps x 4 rows per block
    if (blocks > 2368) blocks = 2368;
    if (blocks < 1) blocks = 1;
    k_finalize<<<blocks, 256, 0, s>>>(c.flags(), MI, c.S0user(), c.ldS0user(), c.slots(), c.slot_stride(),
                                      training ? (L->xlay ? 2 : 0) : 1, L->xlay ? L->ldX : D, N, D, io->state_out, io->k_out);
    GNNFP_COUNT_LAUNCH();
  }
  // ---- net_output (+ pooling) ----------------------------------------------------------------------------
  {
    FwdArgs fa;
    memset(&fa, 0, sizeof(fa));
    build_out_src(c, fa.src);
    if (L->bn_train_out) {
      PassArgs pa;
      memset(&pa, 0, sizeof(pa));
      pa.src = fa.src;
      for (int p = 0; p < pa.src.n_pieces; ++p) { pa.src.p[p].st_sum = nullptr; pa.src.p[p].st_sq = nullptr; }
      pa.st_sum = c.stO(); pa.st_sq = c.stO() + L->out_in;
      pa.tc.cap_per_row = L->cap_per_row;
      if ((rc = tile_cfg_pass(L->out_in, pa.src.n_rows, &pa.tc))) return rc;
      if ((rc = launch_tile_pass(pa, s))) return rc;
    }
    fill_netdev(L->onet, *op, training, fa.src.n_rows, fa.net);
    float* on = L->pool ? (float*)(c.ws + L->ws.out_nodes) : io->out;
    if (L->out_gemm_ok) {
      // single Dense layer over plain matrices: BN folded into padded weights, one pipelined GEMM (softmax in the epilogue)
      FoldArgs fo;
      memset(&fo, 0, sizeof(fo));
      fo.src = fa.src; fo.net = fa.net;
      GemmRowsArgs ga;
      memset(&ga, 0, sizeof(ga));
      int k2 = 0;
      for (int p = 0; p < fa.src.n_pieces; ++p) {
        fo.k8[p] = k2;
        gemm_piece_set(ga.p[p], fa.src.p[p].ptr, fa.src.p[p].ld, fa.src.p[p].width, k2);
        k2 += ceil_to(fa.src.p[p].width, 2);
      }
      const int H = L->onet.widths[0];
      float* wf = (float*)(c.ws + L->ws.wfold) + (size_t)L->nt * L->ws.wfold_stride;
      fo.Kpad = gemm_rows_kpad(k2); fo.ldw = gemm_rows_ldw(H); fo.Wp = wf; fo.biasp = wf + (size_t)fo.Kpad * fo.ldw;
      fo.update_moving = training;
      if (training && L->onet.has_bn && L->ws.bncoef_stride)      // the backward's coefficients of net_output, left here
        fo.coef_out = (float*)(c.ws + L->ws.bncoef) + (size_t)L->nt * L->ws.bncoef_stride;
      if ((rc = launch_fold_w(fo, s))) return rc;
      ga.n_rows = fa.src.n_rows; ga.rowlist = fa.src.rowlist; ga.n_pieces = fa.src.n_pieces; ga.Kpad = fo.Kpad;
      ga.Wp = fo.Wp; ga.ldw = fo.ldw; ga.N = H; ga.bias = fo.biasp; ga.act = L->onet.acts[0];
      ga.out = on; ga.ld_out = L->T; ga.out_compact = 1; ga.fwd = 1;
      ga.vec2 = L->T % 2 == 0 && ((uintptr_t)on & 7) == 0;
      // narrow Dense (H <= 4, the starters' Dense(2, softmax)): streaming kernel, warp per row (narrow.cu)
      static const int no_narrow = getenv("GNNFP_NO_NARROW") ? 1 : 0;
      NarrowArgs na;
      memset(&na, 0, sizeof(na));
      na.n_rows = ga.n_rows; na.rowlist = ga.rowlist; na.n_pieces = ga.n_pieces;
      for (int p = 0; p < ga.n_pieces; ++p) na.p[p] = ga.p[p];
      na.K = L->out_in; na.H = H; na.Wp = fo.Wp; na.ldw = fo.ldw; na.bias = fo.biasp; na.act = ga.act; na.out = on; na.ld_out = L->T;
      if (!no_narrow && narrow_supported(na)) {
        if ((rc = launch_narrow_fwd(na, s, PC_FWD_OUT))) return rc;
      } else if ((rc = launch_gemm_rows(ga, s, PC_FWD_OUT))) return rc;
    } else {
      fa.tc.cap_per_row = L->cap_per_row;
      if ((rc = tile_cfg_fwd(fa.net, fa.src.n_rows, &fa.tc))) return rc;
      fa.prev_col0 = -1;
      fa.out = on; fa.ld_out = L->T; fa.out_compact = 1;
      fa.update_moving = training;
      fa.prof_cat = PC_FWD_OUT;
      if ((rc = launch_tile_fwd(fa, s))) return rc;
    }
    if (L->pool) {
      const int tot = g->G * L->T;
      k_pool<<<(tot + 255) / 256, 256, 0, s>>>(on, g->graph_ptr, g->ng_val, g->G, L->T, io->out);
      GNNFP_COUNT_LAUNCH();
    }
    if (io->out_nodes && io->out_nodes != on)
      GNNFP_CHECK_CUDA(cudaMemcpyAsync(io->out_nodes, on, (size_t)L->M * L->T * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return rc;
}

static int fwd_check(gnnfp_loop* L, const gnnfp_net_params* sp, const gnnfp_net_params* op, const gnnfp_loop_io* io,
                     void* workspace, size_t workspace_bytes) {
  int rc;
  if ((rc = check_io(L, io, workspace, workspace_bytes))) return rc;
  if (!sp || !op) GNNFP_FAIL(GNNFP_E_INVALID, "loop_forward: parameters missing");
  for (int t = 0; t < L->nt; ++t) if ((rc = check_params(L->snet[t], sp[t], "net_state"))) return rc;
  if ((rc = check_params(L->onet, *op, "net_output"))) return rc;
  return GNNFP_OK;
}

extern "C" int gnnfp_loop_forward(gnnfp_loop* L, const gnnfp_net_params* sp, const gnnfp_net_params* op,
                                  const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = fwd_check(L, sp, op, io, workspace, workspace_bytes))) return rc;
  Ctx c{L, io, (char*)workspace, (cudaStream_t)stream};
  if ((rc = fwd_begin(c, sp, op))) return rc;
  for (int t = 1; t <= L->cfg.max_iteration; ++t)
    if ((rc = fwd_iter(c, t, sp, op))) return rc;
  return fwd_end(c, sp, op);
}

// ---- stepping API: the same forward, one phase per call, so that a multi-GPU driver can exchange halo rows
// ---- and all-reduce the convergence flag between iterations (partitioned single graph) ----------
extern "C" int gnnfp_loop_forward_begin(gnnfp_loop* L, const gnnfp_net_params* sp, const gnnfp_net_params* op,
                                        const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = fwd_check(L, sp, op, io, workspace, workspace_bytes))) return rc;
  Ctx c{L, io, (char*)workspace, (cudaStream_t)stream};
  return fwd_begin(c, sp, op);
}
extern "C" int gnnfp_loop_forward_iter(gnnfp_loop* L, int32_t t, const gnnfp_net_params* sp, const gnnfp_net_params* op,
                                       const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = fwd_check(L, sp, op, io, workspace, workspace_bytes))) return rc;
  if (t < 1 || t > L->cfg.max_iteration) GNNFP_FAIL(GNNFP_E_INVALID, "loop_forward_iter: t=%d outside 1..max_iteration", t);
  Ctx c{L, io, (char*)workspace, (cudaStream_t)stream};
  return fwd_iter(c, t, sp, op);
}
extern "C" int gnnfp_loop_forward_end(gnnfp_loop* L, const gnnfp_net_params* sp, const gnnfp_net_params* op,
                                      const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = fwd_check(L, sp, op, io, workspace, workspace_bytes))) return rc;
  Ctx c{L, io, (char*)workspace, (cudaStream_t)stream};
  return fwd_end(c, sp, op);
}
extern "C" int gnnfp_loop_ws_layout(const gnnfp_loop* L, int32_t* ld_state, int32_t* state1_slot, int32_t* ld_grad) {
  if (!L) GNNFP_FAIL(GNNFP_E_INVALID, "loop_ws_layout: null plan");
  if (ld_state) *ld_state = L->xlay ? L->ldX : L->D;
  if (state1_slot) *state1_slot = (L->xlay && L->cfg.training) ? 1 : 0;
  if (ld_grad) *ld_grad = L->ldG;
  return GNNFP_OK;
}
extern "C" int gnnfp_loop_ws_offsets(const gnnfp_loop* L, size_t* flags_off, size_t* slots_off, size_t* slot_stride_floats,
                                     int32_t* slot_count) {
  if (!L) GNNFP_FAIL(GNNFP_E_INVALID, "loop_ws_offsets: null plan");
  if (flags_off) *flags_off = L->ws.flags;
  if (slots_off) *slots_off = L->ws.slots;
  if (slot_stride_floats) *slot_stride_floats = ((size_t)L->N * (L->xlay ? L->ldX : L->D) + 31) / 32 * 32;
  if (slot_count) *slot_count = L->slot_count;
  return GNNFP_OK;
}
