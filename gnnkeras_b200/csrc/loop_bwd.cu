// loop_bwd.cu - hand-written BPTT over the k executed iterations (replaces tf.GradientTape in
// train_step, reference GNN.py:284-295 / LGNN.py:259-272 / CompositeGNN.py:282-293).
//
//   1. net_output backward (un-pooling through NodeGraph fused into the gradient staging)
//      -> dL/ds_final (+ d_nodes through the [state|nodes] concat, GNN.py:241)
//   2. for t = max_iteration .. 1, gated on the same device flags as the forward (no host sync):
//        G_t = (t is the last executed) ? dL/ds_final : dOwn_{t+1} + Adj . dAgg_{t+1}
//        tile_bwd_kernel -> dW/db partials, dOwn_t, dAgg_t (+ static-column gradients)
//        BN training: bn_reduce + tile_bnfix
//   3. input gradients (LGNN chaining): d_state0 / d_nodes / d_arc_labels
//   4. deterministic reduction of the per-CTA partials, optional /k (average_st_grads)
#include <stdlib.h>

#include "loop.h"
#include "tile.cuh"
#include "gemm.h"
#include "rows_tma.h"

// G0 and the loop-invariant aggregates' gradients -> d_state0 / d_nodes
struct InGradArgs {
  int N, D, S, NLp, NLw, AL, LsM, composite, nt;
  int dt[GNNFP_MAX_TYPES], doff[GNNFP_MAX_TYPES];
  const uint8_t* type_mask;          // [nt, N]
  const int* flags;
  const float* dSfin; const float* dOwn1; const float* dAgg1; const float* dXs;
  const int* src_rowptr; const int* src_dst; const float* src_w;
  float* d_nodes; float* d_state0;
  int want;
  // inline BN-training correction (homogeneous single-layer nets): constants of iteration 1 and static sums
  const float* cn; const float* csum; int in_dim;
  const float* s0; int ld0; const float* agg1; const float* Xs;
  int ldg, ld_agg1;                   // leading dimensions of dSfin / dOwn1 / dAgg1 and of agg1
};
static __global__ void k_input_grads_nodes(const __grid_constant__ InGradArgs a) {
  const int W = a.D > a.NLw ? a.D : a.NLw;
  const size_t total = (size_t)a.N * W;
  const bool ran = a.flags[0] != 0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / W), j = (int)(e - (size_t)i * W);
    const int a0 = a.src_rowptr[i], a1 = a.src_rowptr[i + 1];
    if (j < a.D) {   // G0 = dL/d state0
      float g0;
      if (ran) {
        g0 = a.dOwn1[(size_t)i * a.ldg + j];
        const float* cn = a.cn;
        const int in = a.in_dim, ac = a.D + a.NLp + j;
        if (cn) g0 -= cn[j] + fmaf(a.s0[(size_t)i * a.ld0 + j], cn[2 * in + j], cn[3 * in + j]) * cn[in + j];
        for (int p = a0; p < a1; ++p) {
          const size_t so = (size_t)a.src_dst[p] * a.ldg + j;
          float tv = a.dAgg1[so];
          if (cn) tv -= cn[ac] + fmaf(a.agg1[(size_t)a.src_dst[p] * a.ld_agg1 + j], cn[2 * in + ac], cn[3 * in + ac]) * cn[in + ac];
          g0 = fmaf(a.src_w ? a.src_w[p] : 1.0f, tv, g0);
        }
      } else {
        g0 = a.dSfin[(size_t)i * a.ldg + j];
      }
      if (a.S > 0) { if (a.d_state0) a.d_state0[(size_t)i * a.S + j] = g0; }
      else if (a.d_nodes) a.d_nodes[(size_t)i * a.NLw + j] += g0;           // state0 = nodes (GNN.py:259)
    }
    if (a.d_nodes && a.dXs && j < a.NLw) {
      float g = 0.f;
      if (!a.composite) {
        if (a.NLp) {   // own-label columns + Adj . d(agg_nodes)
          const bool fixs = a.cn != nullptr && a.csum != nullptr && ran;
          const int in = a.in_dim, ic0 = a.D + j, ic1 = 2 * a.D + a.NLp + j;      // input columns of Xs[:, j] and Xs[:, NLp + j]
          g = a.dXs[(size_t)i * a.LsM + j];
          if (fixs) g -= a.csum[ic0] + fmaf(a.Xs[(size_t)i * a.LsM + j], a.cn[2 * in + ic0], a.cn[3 * in + ic0]) * a.csum[in + ic0];
          for (int p = a0; p < a1; ++p) {
            const size_t so = (size_t)a.src_dst[p] * a.LsM + a.NLp + j;
            float tv = a.dXs[so];
            if (fixs) tv -= a.csum[ic1] + fmaf(a.Xs[so], a.cn[2 * in + ic1], a.cn[3 * in + ic1]) * a.csum[in + ic1];
            g = fmaf(a.src_w ? a.src_w[p] : 1.0f, tv, g);
          }
        }
      } else {
        for (int t = 0; t < a.nt; ++t) {
          if (j < a.dt[t] && a.type_mask[(size_t)t * a.N + i]) {   // CompositeAdjacencies[t] keeps arcs whose source is type t
            for (int p = a0; p < a1; ++p) g = fmaf(a.src_w ? a.src_w[p] : 1.0f, a.dXs[(size_t)a.src_dst[p] * a.LsM + a.doff[t] + j], g);
          }
        }
      }
      a.d_nodes[(size_t)i * a.NLw + j] += g;
    }
  }
}
// d_arc_labels[a] += v_a * d(agg_arcs)[dst_a]   (ArcNode . dInp[:, agg_arcs cols])
static __global__ void k_input_grads_arcs(int A, int AL, int LsM, int col0, const int* dst, const float* val,
                                          const float* dXs, float* d_arcs, const float* cn, const float* csum,
                                          int in_dim, int in_col0, const float* Xs, const int* flags) {
  const size_t total = (size_t)A * AL;
  const bool fixs = cn != nullptr && csum != nullptr && flags[0] != 0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int ar = (int)(e / AL), c = (int)(e - (size_t)ar * AL);
    const size_t so = (size_t)dst[ar] * LsM + col0 + c;
    float tv = dXs[so];
    if (fixs) {
      const int ic = in_col0 + c;
      tv -= csum[ic] + fmaf(Xs[so], cn[2 * in_dim + ic], cn[3 * in_dim + ic]) * csum[in_dim + ic];
    }
    d_arcs[e] += val[ar] * tv;
  }
}

// arc focus: d(state)[i] (+ d(nodes)[i]) += sum over the out-arcs of i of the source-side block + sum over the in-arcs of i of the
// destination-side block of the per-arc gradient tmp[A][2 aw]; arcs in CSR (= arc id) order, one thread per (node, column)
static __global__ void k_arc_grad_gather(int N, int D, int cols, const int* src_rowptr, const int* src_arc, const int* dst_rowptr,
                                         const int* dst_arc, const float* tmp, int aw, float* dS, int ldS, float* d_nodes, int ldn) {
  for (int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); i < N; i += gridDim.x * (blockDim.x / 32)) {
    for (int c = threadIdx.x & 31; c < cols; c += 32) {
      float acc = 0.f;
      for (int q = src_rowptr[i]; q < src_rowptr[i + 1]; ++q) acc += tmp[(size_t)src_arc[q] * 2 * aw + c];
      for (int q = dst_rowptr[i]; q < dst_rowptr[i + 1]; ++q) acc += tmp[(size_t)dst_arc[q] * 2 * aw + aw + c];
      if (c < D) dS[(size_t)i * ldS + c] += acc;
      else if (d_nodes) d_nodes[(size_t)i * ldn + (c - D)] += acc;
    }
  }
}

// dz of net_output's single Dense layer, compact [M, T]:  g = un-pooled d_out (+ d_out_nodes),  dz = act'(y) g
// (softmax: dz_j = y_j (g_j - sum_k g_k y_k), as TF's SoftmaxGrad)
struct OutDzArgs {
  int M, T, act;
  const int* rowlist;
  const float* saved;          // [M, T] outputs of the forward (compact)
  const float* d_out;          // pooled: row node2graph[gr] scaled by ng_val[gr]; else compact [M, T]
  const int* map; const float* rowscale;
  const float* d_out_nodes;    // compact [M, T] or NULL
  float* dz;
};
static __global__ void k_out_dz(const __grid_constant__ OutDzArgs a) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < a.M; r += gridDim.x * blockDim.x) {
    const int gr = a.rowlist ? a.rowlist[r] : r;
    const float* y = a.saved + (size_t)r * a.T;
    const float* g1 = nullptr;
    float sc = 1.0f;
    if (a.d_out) {
      g1 = a.d_out + (size_t)(a.map ? a.map[gr] : r) * a.T;
      if (a.rowscale) sc = a.rowscale[gr];
    }
    const float* g2 = a.d_out_nodes ? a.d_out_nodes + (size_t)r * a.T : nullptr;
    float* dz = a.dz + (size_t)r * a.T;
    if (a.act == GNNFP_ACT_SOFTMAX) {
      float dot = 0.f;
      for (int j = 0; j < a.T; ++j) {
        float g = g1 ? g1[j] * sc : 0.f;
        if (g2) g += g2[j];
        dot = fmaf(g, y[j], dot);
      }
      for (int j = 0; j < a.T; ++j) {
        float g = g1 ? g1[j] * sc : 0.f;
        if (g2) g += g2[j];
        dz[j] = y[j] * (g - dot);
      }
    } else {
      for (int j = 0; j < a.T; ++j) {
        float g = g1 ? g1[j] * sc : 0.f;
        if (g2) g += g2[j];
        dz[j] = act_bwd(a.act, y[j], g);
      }
    }
  }
}

// argument block of the TMA backward dX of iteration t:  [dOwn_t | dAgg_t] = dz_t (W^T . gamma rstd) - BN-training correction,
// both input blocks in one pass over dz (the side inputs x = S_{t-1} / Adj^T S_{t-1} come from the interleaved slot X_{t-1})
static int rt_build_dx(const Ctx& c, int t, const gnnfp_net_params* sp, const float* dz, float* dOwn, float* dAgg,
                       const float* coef, const float* cn, const int* gate, RowsTmaArgs& ra) {
  gnnfp_loop* L = c.L;
  const int D = L->D, N = L->N, in = L->snet[0].in_dim, H0 = L->snet[0].widths[0];
  const int NLp = L->S > 0 ? L->NLw : 0;
  const bool bn = L->snet[0].has_bn != 0;
  int rc;
  memset(&ra, 0, sizeof(ra));
  ra.mode = RT_DX; ra.n_rows = L->Nact; ra.H = H0;
  if ((rc = rows_tma_map(&ra.maps[0], dz, L->Nact, H0, L->ldG))) return rc;       // rows past the computed ones read as zeros
  if ((rc = rows_tma_map(&ra.maps[1], c.S(t - 1), N, D, L->ldX))) return rc;
  if ((rc = rows_tma_map(&ra.maps[2], c.S(t - 1), N, 2 * D, L->ldX))) return rc;
  if ((rc = rows_tma_map(&ra.maps[3], dOwn, L->Nact, D, L->ldG))) return rc;
  if ((rc = rows_tma_map(&ra.maps[4], dAgg, L->Nact, D, L->ldG))) return rc;
  for (int c0 = 0; c0 < H0; c0 += RT_CHUNK) {
    RtKChunk& k = ra.kc[ra.n_kc++];
    k.map = 0; k.col0 = c0; k.width = H0 - c0 < RT_CHUNK ? H0 - c0 : RT_CHUNK; k.k8 = (k.width + 7) / 8;
    for (int j = 0; j < RT_CHUNK; ++j) k.wrow[j] = (short)(j < k.width ? c0 + j : -1);
  }
  const int Dp = ceil_to(D, 16);
  ra.n_blk = 2;
  ra.blk_acc0[0] = 0; ra.blk_in0[0] = 0; ra.blk_w[0] = D;
  ra.blk_acc0[1] = Dp; ra.blk_in0[1] = D + NLp; ra.blk_w[1] = D;
  for (int b = 0; b < 2; ++b)
    for (int c0 = 0; c0 < D; c0 += RT_CHUNK) {
      RtOChunk& o = ra.oc[ra.n_oc++];
      o.acc_col0 = ra.blk_acc0[b] + c0; o.out_map = 3 + b; o.out_col0 = c0;
      o.aux_map = (bn && cn) ? 1 + b : -1; o.aux_col0 = b ? D + c0 : c0;
      o.cidx0 = ra.blk_in0[b] + c0; o.width = D - c0 < RT_CHUNK ? D - c0 : RT_CHUNK;
    }
  ra.BN = 2 * Dp;
  ra.W = sp[0].W[0];
  ra.colscale = bn ? coef + 2 * in : nullptr;
  ra.corr = bn ? cn : nullptr; ra.corr_in = in;       // cn == NULL: the consumer (dz_kernel / input-gradient kernels) applies the correction
  ra.gate = gate;
  return rows_tma_finish(ra);
}

// phases of the backward (the monolithic entry point runs BEGIN | ITERS | END; the stepping entry point runs one at a
// time so that a multi-GPU driver can reduce the halo rows of Adj . dAgg between iterations)
enum { PH_BEGIN = 1, PH_ITERS = 2, PH_END = 4, PH_GATHER = 8 };

static int backward_impl(gnnfp_loop* L, const gnnfp_net_params* sp, const gnnfp_net_params* op,
                         const gnnfp_loop_io* io, const gnnfp_loop_grads* gr, gnnfp_net_params* dsp,
                         gnnfp_net_params* dop, void* workspace, size_t workspace_bytes, void* stream, int phases, int t_only) {
  int rc;
  if ((rc = check_io(L, io, workspace, workspace_bytes))) return rc;
  if (!sp || !op || !gr || !dsp || !dop) GNNFP_FAIL(GNNFP_E_INVALID, "loop_backward: null argument");
  if (!L->cfg.training) GNNFP_FAIL(GNNFP_E_INVALID, "loop_backward needs a plan created with training=1 (states are not saved otherwise)");
  for (int t = 0; t < L->nt; ++t) {
    if ((rc = check_params(L->snet[t], sp[t], "net_state"))) return rc;
    gnnfp_net_desc nd = L->snet[t]; nd.has_bn = 0;
    if ((rc = check_params(nd, dsp[t], "d net_state"))) return rc;
    if (L->snet[t].has_bn && (!dsp[t].bn_gamma || !dsp[t].bn_beta)) GNNFP_FAIL(GNNFP_E_INVALID, "d net_state: BN gradient buffers missing");
  }
  if ((rc = check_params(L->onet, *op, "net_output"))) return rc;
  { gnnfp_net_desc nd = L->onet; nd.has_bn = 0; if ((rc = check_params(nd, *dop, "d net_output"))) return rc; }
  if (L->onet.has_bn && (!dop->bn_gamma || !dop->bn_beta)) GNNFP_FAIL(GNNFP_E_INVALID, "d net_output: BN gradient buffers missing");
  for (int l = 1; l < L->onet.n_layers; ++l)
    if (L->onet.acts[l - 1] == GNNFP_ACT_SOFTMAX) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "softmax is only supported as the last activation in the backward");
  for (int t = 0; t < L->nt; ++t)
    for (int l = 1; l < L->snet[t].n_layers; ++l)
      if (L->snet[t].acts[l - 1] == GNNFP_ACT_SOFTMAX) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "softmax is only supported as the last activation in the backward");
  const int want = L->cfg.want_input_grads;
  if ((want & 1) && !gr->d_nodes) GNNFP_FAIL(GNNFP_E_INVALID, "loop_backward: d_nodes missing");
  if ((want & 2) && L->AL > 0 && !gr->d_arc_labels) GNNFP_FAIL(GNNFP_E_INVALID, "loop_backward: d_arc_labels missing");
  if ((want & 4) && L->S > 0 && !gr->d_state0) GNNFP_FAIL(GNNFP_E_INVALID, "loop_backward: d_state0 missing");

  cudaStream_t s = (cudaStream_t)stream;
  Ctx c{L, io, (char*)workspace, s};
  const gnnfp_graph* g = L->g;
  const int MI = L->cfg.max_iteration, D = L->D, N = L->N;
  const int ldG = L->ldG;
  const size_t ND = (size_t)N * ldG;
  float* dSfin = (float*)(c.ws + L->ws.dSfin);
  const size_t NDa = (ND + 31) / 32 * 32;
  float* dOwn[2] = {(float*)(c.ws + L->ws.dOwn), (float*)(c.ws + L->ws.dOwn) + NDa};
  float* dAgg[2] = {(float*)(c.ws + L->ws.dAgg), (float*)(c.ws + L->ws.dAgg) + NDa};
  float* dXs = want ? (float*)(c.ws + L->ws.dXs) : nullptr;
  float* part_state = (float*)(c.ws + L->ws.part_state);
  float* part_out = (float*)(c.ws + L->ws.part_out);
  float* bn_part = (float*)(c.ws + L->ws.bn_part);
  float* bn_const = (float*)(c.ws + L->ws.bn_const);
  float* bn_grad = (float*)(c.ws + L->ws.bn_grad);
  float* bn_static = (float*)(c.ws + L->ws.bn_static);
  int din_max_ = L->onet.in_dim;
  for (int t = 0; t < L->nt; ++t) din_max_ = L->snet[t].in_dim > din_max_ ? L->snet[t].in_dim : din_max_;
  auto cn_t = [&](int t) { return (float*)(c.ws + L->ws.bn_const_t) + (size_t)t * 4 * din_max_; };
  const float* src_w = g->mode == GNNFP_AGG_SUM ? nullptr : g->src_w;

  if (phases & PH_BEGIN) {
    GNNFP_CHECK_CUDA(cudaMemsetAsync(c.ws + L->ws.bwd_zero, 0, L->ws.bwd_zero_bytes, s));
    if (gr->d_state) GNNFP_CHECK_CUDA(cudaMemcpy2DAsync(dSfin, (size_t)ldG * sizeof(float), gr->d_state, (size_t)D * sizeof(float),
                                                        (size_t)D * sizeof(float), (size_t)N, cudaMemcpyDeviceToDevice, s));
    else GNNFP_CHECK_CUDA(cudaMemsetAsync(dSfin, 0, ND * sizeof(float), s));
    if ((want & 1)) GNNFP_CHECK_CUDA(cudaMemsetAsync(gr->d_nodes, 0, (size_t)N * L->NLw * sizeof(float), s));
    if ((want & 2) && L->AL > 0) GNNFP_CHECK_CUDA(cudaMemsetAsync(gr->d_arc_labels, 0, (size_t)L->A * L->AL * sizeof(float), s));
    for (int t = 0; t < GNNFP_MAX_TYPES; ++t) L->bwd_grid_state[t] = 0;
    L->bwd_grid_out = 0;
  }

  // bn_grad layout: [state net 0 | state net 1 | ... | out net]
  size_t bg_off[GNNFP_MAX_TYPES + 1];
  {
    size_t o = 0;
    for (int t = 0; t < L->nt; ++t) { bg_off[t] = o; o += 2 * (size_t)L->snet[t].in_dim; }
    bg_off[L->nt] = o;
  }
  size_t ps_off[GNNFP_MAX_TYPES];
  {
    size_t o = 0;
    for (int t = 0; t < L->nt; ++t) { ps_off[t] = o; o += (size_t)L->grid_cap * L->nparam_s[t]; }
  }
  int* grid_state = L->bwd_grid_state;       // widest grids used so far: the partial slots the final reduction must read
  int& grid_out = L->bwd_grid_out;
  const int NLp = (!L->composite && L->S > 0) ? L->NLw : 0;

  // ---- 1. net_output backward -----------------------------------------------------------------
  NetDev ond;
  fill_netdev(L->onet, *op, 1, L->M, ond);
  if ((phases & PH_BEGIN) && (gr->d_out || gr->d_out_nodes)) {
    BwdArgs ba;
    memset(&ba, 0, sizeof(ba));
    build_out_src(c, ba.src);
    const bool arc = L->cfg.kind == GNNFP_KIND_ARC;
    // arc focus: net_output reads [state[src] | nodes[src]? | state[dst] | nodes[dst]? | arc labels] per arc (GNN.py:317-330).
    // Its input gradient is first STORED per arc (row = arc id, no atomics), then summed per node over the node's out-arcs
    // (source side) and in-arcs (destination side) in CSR order: deterministic, unlike a scatter with atomicAdd
    float* arc_tmp = arc ? (float*)(c.ws + L->ws.arc_tmp) : nullptr;
    const int aw = D + NLp;                           // columns per side
    if (arc) GNNFP_CHECK_CUDA(cudaMemsetAsync(arc_tmp, 0, (size_t)L->A * 2 * aw * sizeof(float), s));   // unmasked arcs contribute nothing
    int n_state_seen = 0, n_nodes_seen = 0;
    for (int p = 0; p < ba.src.n_pieces; ++p) {
      Piece& pc = ba.src.p[p];
      if (pc.tag == TAG_STATE) {
        if (arc) { pc.gptr = arc_tmp + (size_t)(n_state_seen++ ? aw : 0); pc.gld = 2 * aw; pc.gmode = GM_STORE; pc.gdirect = 1; }
        else { pc.gptr = dSfin; pc.gld = ldG; pc.gmode = GM_ADD; }
      } else if (pc.tag == TAG_NODES) {
        if (arc) { if (want & 1) { pc.gptr = arc_tmp + (size_t)(n_nodes_seen ? aw : 0) + D; pc.gld = 2 * aw; pc.gmode = GM_STORE; pc.gdirect = 1; } ++n_nodes_seen; }
        else if (want & 1) { pc.gptr = gr->d_nodes; pc.gld = L->NLw; pc.gmode = GM_ADD; }
      }
      else if (pc.tag == TAG_ARC_LABELS) { if ((want & 2) && L->AL > 0) { pc.gptr = gr->d_arc_labels; pc.gld = L->AL; pc.gmode = GM_ADD; } }
    }
    bool og = L->out_gemm_ok && !arc && getenv("GNNFP_NO_GEMM_BWD") == nullptr && ba.src.n_pieces <= GEMM_MAXP;
    int ok2 = 0;
    for (int p = 0; p < ba.src.n_pieces; ++p) {
      const Piece& pc = ba.src.p[p];
      if (pc.kind != PK_DIRECT || pc.map || pc.rowscale || pc.compact || pc.gate || pc.width > 80) og = false;
      ok2 += ceil_to(pc.width, 2);
    }
    og = og && gemm_dw_supported(ok2, L->T) && gemm_rows_supported(ceil_to(L->T, 8), 80);
    if (og) {
      // GEMM path: dz (compact) -> dW/db + BN sums -> BN constants -> dX per input block with the BN-training
      // correction folded into the epilogue (no separate fix-up pass over dSfin)
      const int in = L->onet.in_dim, H = L->T;
      const bool bn = L->onet.has_bn != 0;
      float* dzo = (float*)(c.ws + L->ws.dOutN);
      OutDzArgs oa;
      memset(&oa, 0, sizeof(oa));
      oa.M = L->M; oa.T = H; oa.act = L->onet.acts[0]; oa.rowlist = ba.src.rowlist;
      oa.saved = L->pool ? (const float*)(c.ws + L->ws.out_nodes) : io->out;
      oa.d_out = gr->d_out;
      if (gr->d_out && L->pool) { oa.map = g->node2graph; oa.rowscale = g->ng_val; }
      oa.d_out_nodes = gr->d_out_nodes;
      oa.dz = dzo;
      {
        int blocks = (L->M + 255) / 256;
        if (blocks > 2368) blocks = 2368;
        if (blocks < 1) blocks = 1;
        k_out_dz<<<blocks, 256, 0, s>>>(oa);
        GNNFP_COUNT_LAUNCH();
      }
      float* coef = (float*)(c.ws + L->ws.bncoef) + (size_t)L->nt * L->ws.bncoef_stride;
      if (bn && !L->out_gemm_ok) {                    // (the GEMM forward's fold_w_kernel left them: loop.cu fwd_end)
        BnCoefArgs bc;
        memset(&bc, 0, sizeof(bc));
        bc.src = ba.src; bc.net = ond; bc.coef = coef;
        if ((rc = launch_bn_coef(bc, s))) return rc;
      }
      GemmDwArgs dw;
      memset(&dw, 0, sizeof(dw));
      dw.n_rows = ba.src.n_rows; dw.rowlist = ba.src.rowlist; dw.n_pieces = ba.src.n_pieces;
      int k2 = 0;
      for (int p = 0; p < ba.src.n_pieces; ++p) {
        gemm_piece_set(dw.p[p], ba.src.p[p].ptr, ba.src.p[p].ld, ba.src.p[p].width, k2);
        k2 += ceil_to(ba.src.p[p].width, 2);
      }
      dw.Kp = k2; dw.dz = dzo; dw.ld_dz = H; dw.H = H; dw.dz_compact = 1;
      dw.partial = part_out; dw.n_params = L->nparam_o; dw.bias_off = in * H;
      dw.W = ond.W[0];
      if (bn) { dw.bnA = coef; dw.bnB = coef + in; dw.gamma = ond.gamma; dw.beta = ond.beta; dw.bn_partial = bn_part; }
      // narrow Dense (H <= 4): streaming kernels, warp per row (narrow.cu) - no transposed weight copies, no GEMM tiles
      static const int no_narrow = getenv("GNNFP_NO_NARROW") ? 1 : 0;
      NarrowArgs na;
      memset(&na, 0, sizeof(na));
      na.n_rows = dw.n_rows; na.rowlist = dw.rowlist; na.n_pieces = dw.n_pieces;
      for (int p = 0; p < dw.n_pieces; ++p) na.p[p] = dw.p[p];
      na.K = in; na.H = H; na.dz = dzo;
      na.partial = dw.partial; na.n_params = dw.n_params; na.bias_off = dw.bias_off;
      na.W = dw.W; na.bnA = dw.bnA; na.bnB = dw.bnB; na.gamma = dw.gamma; na.beta = dw.beta; na.bn_partial = dw.bn_partial;
      const bool narrow = !no_narrow && narrow_supported(na);
      if (narrow) { if ((rc = launch_narrow_dw(na, s, PC_BWD_OUT, &grid_out))) return rc; }
      else if ((rc = launch_gemm_dw(dw, s, PC_BWD_OUT, &grid_out))) return rc;
      if (bn) {
        ba.net = ond; ba.bn_partial = bn_part; ba.tc.grid = grid_out;
        if ((rc = launch_bn_tail(ba, bn_grad + bg_off[L->nt], bn_const, s, 1, nullptr))) return rc;
      }
      if (narrow) {
        bool any = false;
        for (int p = 0; p < ba.src.n_pieces; ++p) {
          const Piece& pc = ba.src.p[p];
          na.gout[p] = pc.gptr; na.gld[p] = pc.gld; na.gadd[p] = pc.gmode == GM_ADD;
          any = any || pc.gptr != nullptr;
        }
        na.colscale = bn ? coef + 2 * in : nullptr;
        na.corr = bn ? bn_const : nullptr; na.corr_in = in;
        if (any && (rc = launch_narrow_dx(na, s, PC_BWD_OUT))) return rc;
      }
      float* wt = (float*)(c.ws + L->ws.wtb) + (size_t)L->nt * L->ws.wtb_stride;
      const int KH = gemm_rows_kpad(H);
      for (int p = 0; !narrow && p < ba.src.n_pieces; ++p) {
        const Piece& pc = ba.src.p[p];
        if (!pc.gptr) continue;
        const int ldw = gemm_rows_ldw(pc.width);
        if ((rc = launch_transpose_block(ond.W[0], H, pc.col0, pc.width, KH, ldw, wt, s))) return rc;
        GemmRowsArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.n_rows = ba.src.n_rows; ga.rowlist = ba.src.rowlist; ga.n_pieces = 1; ga.a_compact = 1;
        gemm_piece_set(ga.p[0], dzo, H, H, 0);
        ga.fwd = 0; ga.Kpad = KH; ga.Wp = wt; ga.ldw = ldw; ga.N = pc.width;
        ga.colscale = bn ? coef + 2 * in + pc.col0 : nullptr;
        if (bn) { ga.corr = bn_const; ga.corr_in = in; ga.corr_col0 = pc.col0; ga.corr_x = pc.ptr; ga.corr_ld = pc.ld; }
        ga.out = pc.gptr; ga.ld_out = pc.gld; ga.out_add = pc.gmode == GM_ADD;
        ga.vec2 = pc.gld % 2 == 0 && ((uintptr_t)pc.gptr & 7) == 0;
        if ((rc = launch_gemm_rows(ga, s, PC_BWD_OUT))) return rc;
        wt += (size_t)KH * ldw;
      }
    } else {
    ba.gsrc.n_rows = ba.src.n_rows; ba.gsrc.rowlist = ba.src.rowlist; ba.gsrc.in_dim = L->T;
    if (gr->d_out) {
      Piece p = mk_direct(gr->d_out, L->T, L->T, 0);
      if (L->pool) { p.map = g->node2graph; p.rowscale = g->ng_val; }   // d out_nodes[i] = NodeGraph[i,g(i)] * d out[g(i)]
      else p.compact = 1;
      add_piece(ba.gsrc, p);
    }
    if (gr->d_out_nodes) {
      Piece p = mk_direct(gr->d_out_nodes, L->T, L->T, 0);
      p.compact = 1;
      p.accumulate = gr->d_out ? 1 : 0;
      add_piece(ba.gsrc, p);
    }
    ba.net = ond;
    ba.tc.cap_per_row = L->cap_per_row;
    if ((rc = tile_cfg_bwd(ba.net, ba.src.n_rows, L->T, &ba.tc))) return rc;
    grid_out = ba.tc.grid;
    ba.saved_out = L->pool ? (const float*)(c.ws + L->ws.out_nodes) : io->out;
    ba.ld_saved = L->T; ba.saved_compact = 1;
    ba.partial = part_out; ba.n_params = L->nparam_o;
    ba.bn_partial = bn_part;
    ba.prof_cat = PC_BWD_OUT;
    if ((rc = launch_tile_bwd(ba, s))) return rc;
    if (L->onet.has_bn && (rc = launch_bn_tail(ba, bn_grad + bg_off[L->nt], bn_const, s))) return rc;
    if (arc) {
      const int cols = D + ((want & 1) ? NLp : 0);
      const size_t total = (size_t)N * cols;
      int blocks = (int)((total + 255) / 256);
      if (blocks > 4736) blocks = 4736;
      if (blocks < 1) blocks = 1;
      k_arc_grad_gather<<<blocks, 256, 0, s>>>(N, D, cols, g->src_rowptr, g->src_arc, g->dst_rowptr, g->dst_arc, arc_tmp, aw,
                                               dSfin, ldG, (want & 1) ? gr->d_nodes : nullptr, L->NLw);
      GNNFP_COUNT_LAUNCH();
    }
    }
  }

  // ---- 2. iterations, newest first ---------------------------------------------------------------
  // Single-Dense-layer state nets (the reference's default MLP) take the "dz path": a streaming kernel forms
  // dz = act'(s_t) * G_t (with the Adj . dAgg gather), then the GEMM pair runs as 1..n column-split launches
  // whose inputs are all plain matrices.  Other nets use the fused kernel.
  float* dzbuf = (float*)(c.ws + L->ws.dz);
  struct Split { int p0, p1, c_off, width; };
  Split splits[GNNFP_MAX_TYPES][4];
  int nsplit[GNNFP_MAX_TYPES];
  bool dzpath[GNNFP_MAX_TYPES];
  for (int ty = 0; ty < L->nt; ++ty) {
    const gnnfp_net_desc& d = L->snet[ty];
    dzpath[ty] = d.n_layers == 1 && d.acts[0] != GNNFP_ACT_SOFTMAX && MI > 0;
    nsplit[ty] = 1;
    if (!dzpath[ty]) continue;
    TileSrc probe;
    build_state_src(c, ty, 1, probe, 1);
    NetDev nd;
    fill_netdev(d, sp[ty], 1, probe.n_rows, nd);
    TileCfg tcf;
    memset(&tcf, 0, sizeof(tcf));
    tcf.cap_per_row = L->cap_per_row;
    int want_splits = 1;
    // measured on B200 (C2): one launch over all input columns beats column splits whenever it fits shared memory
    if (tile_cfg_bwd(nd, probe.n_rows > 0 ? probe.n_rows : 1, D, &tcf) != GNNFP_OK)
      want_splits = probe.n_pieces >= 2 ? 2 : 1;
    if (d.in_dim > 384 && probe.n_pieces >= 3) want_splits = 3;
    if (const char* ev = getenv("GNNFP_BWD_SPLITS")) { const int v = atoi(ev); if (v >= 1 && v <= 3 && v <= probe.n_pieces) want_splits = v; }
    // cut at piece boundaries so that the widest split is as narrow as possible (<= 3 splits: brute force)
    int pref[GNNFP_MAXP + 1];
    pref[0] = 0;
    for (int p = 0; p < probe.n_pieces; ++p) pref[p + 1] = pref[p] + probe.p[p].width;
    const int np = probe.n_pieces;
    int best_a = np, best_b = np, best_w = d.in_dim + 1;
    if (want_splits == 1) { best_a = np; best_b = np; best_w = d.in_dim; }
    else {
      for (int ca = 1; ca < np; ++ca) {
        if (want_splits == 2) {
          const int w1 = pref[ca], w2 = pref[np] - pref[ca];
          const int mw = w1 > w2 ? w1 : w2;
          if (mw < best_w) { best_w = mw; best_a = ca; best_b = np; }
        } else {
          for (int cb = ca + 1; cb < np; ++cb) {
            int mw = pref[ca];
            if (pref[cb] - pref[ca] > mw) mw = pref[cb] - pref[ca];
            if (pref[np] - pref[cb] > mw) mw = pref[np] - pref[cb];
            if (mw < best_w) { best_w = mw; best_a = ca; best_b = cb; }
          }
        }
      }
    }
    int ns = 0;
    const int cuts[4] = {0, best_a, best_b, np};
    for (int k2 = 0; k2 < 3; ++k2) {
      if (cuts[k2] >= cuts[k2 + 1]) continue;
      splits[ty][ns++] = Split{cuts[k2], cuts[k2 + 1], pref[cuts[k2]], pref[cuts[k2 + 1]] - pref[cuts[k2]]};
    }
    nsplit[ty] = ns;
  }
  // GEMM path (gemm.cu) for homogeneous single-layer nets whose pieces are plain matrices of <= 80 columns: dW by
  // gemm_dw, dX per destination block by gemm_rows over the transposed weight block (built once per call here).
  bool gemm_bwd[GNNFP_MAX_TYPES];
  for (int ty = 0; ty < L->nt; ++ty) {
    gemm_bwd[ty] = false;
    if (!dzpath[ty] || !L->gemm_ok[ty] || L->composite || getenv("GNNFP_NO_GEMM_BWD")) continue;
    TileSrc probe;
    build_state_src(c, ty, 1, probe, 1);
    const int H0 = L->snet[ty].widths[0];
    bool ok = true;
    int k2 = 0;
    for (int p = 0; p < probe.n_pieces; ++p) {
      const Piece& pc = probe.p[p];
      if (pc.kind != PK_DIRECT || pc.map || pc.rowscale || pc.compact || pc.gate || pc.width > 80) ok = false;
      k2 += ceil_to(pc.width, 2);
    }
    if (!ok || probe.n_pieces > GEMM_MAXP || !gemm_dw_supported(k2, H0) || !gemm_rows_supported(ceil_to(H0, 8), 80)) continue;
    gemm_bwd[ty] = true;
    float* wt = (float*)(c.ws + L->ws.wtb) + (size_t)ty * L->ws.wtb_stride;
    const int KH = gemm_rows_kpad(H0);
    for (int p = 0; p < probe.n_pieces && (phases & PH_BEGIN); ++p) {
      const int ldw = gemm_rows_ldw(probe.p[p].width);
      // X-slot path: dOwn / dAgg come from rows_tma (it folds W itself), and a static block only has a consumer when input
      // gradients are wanted - no transposed copy for blocks nobody reads
      const bool via_rt = L->xlay && (probe.p[p].tag == TAG_STATE || probe.p[p].tag == TAG_AGG_STATE);
      const bool unused_static = probe.p[p].tag == TAG_STATIC && !want;
      if (!via_rt && !unused_static &&
          (rc = launch_transpose_block(sp[ty].W[0], H0, probe.p[p].col0, probe.p[p].width, KH, ldw, wt, s))) return rc;
      wt += (size_t)KH * ldw;
    }
  }
  float* pgather = L->Nact < L->N ? (float*)(c.ws + L->ws.pgather) : nullptr;
  if ((phases & PH_GATHER) && pgather && t_only >= 1 && t_only < MI) {
    // partitioned graph: this rank's share of Adj . dAgg_{t+1} for ALL local rows (owned + halo); the driver then sums
    // the halo rows into their owners' rows before iteration t consumes the buffer
    AggArgs aa;
    memset(&aa, 0, sizeof(aa));
    aa.n_rows = N; aa.rowlist = nullptr; aa.D = D;
    aa.S = dAgg[(t_only + 1) & 1]; aa.ld = ldG;
    aa.rowptr = g->src_rowptr; aa.idx = g->src_dst; aa.wgt = src_w;
    aa.out = pgather; aa.ld_out = ldG;
    aa.gate = c.flags() + t_only;               // only if iteration t+1 ran
    if ((rc = launch_agg_stats(aa, s))) return rc;
  }
  for (int t = MI; t >= 1 && (phases & PH_ITERS); --t) {
    if (t_only > 0 && t != t_only) continue;
    const int* gate = c.flags() + (t - 1);
    const int wb = t & 1, rb = (t + 1) & 1;
    for (int ty = 0; ty < L->nt; ++ty) {
      TileSrc full;
      build_state_src(c, ty, t, full, 1);
      for (int p = 0; p < full.n_pieces; ++p) {
        Piece& pc = full.p[p];
        if (pc.tag == TAG_AGG_STATE) { pc.gptr = dAgg[wb]; pc.gld = ldG; pc.gmode = GM_STORE; }
        else if (pc.tag == TAG_STATE) { pc.gptr = dOwn[wb]; pc.gld = ldG; pc.gmode = GM_STORE; }
        else if (want) {
          if (pc.tag == TAG_NODES) { if (want & 1) { pc.gptr = gr->d_nodes; pc.gld = L->NLw; pc.gmode = GM_ADD; } }   // composite nodes[:, :d_t]
          else if (pc.tag == TAG_STATIC) {   // static block columns; a block of arc-label aggregates only matters for d_arc_labels
            const bool arcs_only = !L->composite && NLp == 0;
            if (!arcs_only || (want & 2)) { pc.gptr = dXs + (pc.ptr - c.Xs()); pc.gld = L->ldXs; pc.gmode = GM_ADD; }
          }
        }
      }
      NetDev ndfull;
      fill_netdev(L->snet[ty], sp[ty], 1, full.n_rows, ndfull);
      BwdArgs last_ba;
      memset(&last_ba, 0, sizeof(last_ba));
      if (dzpath[ty]) {
        DzArgs da;
        memset(&da, 0, sizeof(da));
        da.n_rows = full.n_rows; da.rowlist = full.rowlist; da.D = D; da.act = L->snet[ty].acts[0];
        da.s_t = c.S(t); da.ld_s = c.ldS(t); da.ldg = ldG; da.ld_agg = c.ldA();
        da.dSfin = dSfin; da.dOwn = dOwn[rb]; da.dAgg = dAgg[rb];
        da.rowptr = g->src_rowptr; da.idx = g->src_dst; da.wgt = src_w;
        da.last_flag = t < MI ? c.flags() + t : nullptr; da.always_last = t == MI;
        da.dz = dzbuf; da.gate = gate;
        if (pgather && t < MI) da.pre = pgather;
        // BN-training correction of iteration t+1's raw gradients: applied here where they are consumed - unless the GEMM
        // path already folded it into the dX epilogue of iteration t+1 (then dOwn / dAgg are final)
        // (TMA path with a state width that is not a multiple of 4: the Adj^T S block of an X slot starts at an unaligned
        //  column, which a tensor-map box cannot address as the dX epilogue's side input - those plans correct here too)
        const bool inline_bn = L->snet[ty].has_bn && !L->composite && (!gemm_bwd[ty] || (L->xlay && D % 4 != 0));
        if (inline_bn && t < MI) {
          da.cn = cn_t(t + 1); da.agg_next = c.AGG(t + 1); da.in_dim = L->snet[ty].in_dim;
          da.own_col0 = 0; da.agg_col0 = D + NLp;
        }
        if ((rc = launch_dz(da, s))) return rc;
        if (L->snet[ty].has_bn && !gemm_bwd[ty])     // (the GEMM dW kernels ASSIGN their CTA's slot; bn_reduce reads the written slots only)
          GNNFP_CHECK_CUDA(cudaMemsetAsync(bn_part, 0, (size_t)L->grid_cap * 2 * L->snet[ty].in_dim * sizeof(float), s));
        const int H0 = L->snet[ty].widths[0];
        int grid_dw = 0;
        if (gemm_bwd[ty]) {
          const int in = L->snet[ty].in_dim;
          float* coef = (float*)(c.ws + L->ws.bncoef) + (size_t)ty * L->ws.bncoef_stride;
          const bool bn = L->snet[ty].has_bn != 0;
          const bool coef_saved = bn && L->xlay;          // left by the forward kernel of iteration t (rows_tma.cu, CTA 0)
          if (coef_saved) coef = (float*)(c.ws + L->ws.bncoef_t) + (size_t)(t - 1) * L->ws.bncoef_stride;
          if (bn && !coef_saved) {
            BnCoefArgs bc;
            memset(&bc, 0, sizeof(bc));
            bc.src = full; bc.net = ndfull; bc.coef = coef; bc.gate = gate;
            if ((rc = launch_bn_coef(bc, s))) return rc;
          }
          GemmDwArgs dw;
          memset(&dw, 0, sizeof(dw));
          dw.n_rows = full.n_rows; dw.rowlist = full.rowlist; dw.n_pieces = full.n_pieces;
          int k2 = 0;
          for (int p = 0; p < full.n_pieces; ++p) {
            gemm_piece_set(dw.p[p], full.p[p].ptr, full.p[p].ld, full.p[p].width, k2);
            k2 += ceil_to(full.p[p].width, 2);
          }
          dw.Kp = k2; dw.dz = dzbuf; dw.ld_dz = ldG; dw.H = H0;
          dw.partial = part_state + ps_off[ty]; dw.n_params = L->nparam_s[ty]; dw.bias_off = in * H0;
          dw.W = ndfull.W[0];
          if (bn) { dw.bnA = coef; dw.bnB = coef + in; dw.gamma = ndfull.gamma; dw.beta = ndfull.beta; dw.bn_partial = bn_part; }
          dw.gate = gate;
          static const int no_dw_tma = getenv("GNNFP_NO_DW_TMA") ? 1 : 0;
          bool dw_done = false;
          if (L->xlay && !no_dw_tma && full.rowlist == nullptr && H0 <= 128 && rows_tma_ok(dzbuf, ldG)) {
            // TMA path (dw_tma.cu): the X slot and dz are read as MN-major boxes, no transposing split
            DwTmaArgs da;
            memset(&da, 0, sizeof(da));
            const int LsM = L->LsM, NLq = L->S > 0 ? L->NLw : 0;
            const int w0 = 2 * D + (L->xs_inline ? LsM : 0);
            da.n_rows = full.n_rows; da.H = H0; da.K = in;
            da.n_zc = (H0 + 31) / 32;
            da.rows = dw_tma_rows((w0 + 31) / 32 + (L->xs_inline ? 0 : (LsM + 31) / 32), da.n_zc);
            bool ok = rows_tma_map(&da.xmap[0], c.S(t - 1), full.n_rows, w0, L->ldX, da.rows, 1) == GNNFP_OK &&   // rows past n_rows: zero-filled
                      rows_tma_map(&da.zmap, dzbuf, full.n_rows, H0, ldG, da.rows, 1) == GNNFP_OK;
            if (ok && !L->xs_inline && LsM > 0) ok = rows_tma_map(&da.xmap[1], c.Xs(), full.n_rows, LsM, L->ldXs, da.rows, 1) == GNNFP_OK;
            for (int c0_ = 0; ok && c0_ < w0; c0_ += 32) {
              if (da.n_xc >= DT_MAXXC) { ok = false; break; }
              da.xc_map[da.n_xc] = 0; da.xc_col0[da.n_xc] = c0_; ++da.n_xc;
            }
            const int sbase = L->xs_inline ? 2 * D : 32 * da.n_xc;          // accumulator column of static column 0
            if (ok && !L->xs_inline)
              for (int c0_ = 0; c0_ < LsM; c0_ += 32) {
                if (da.n_xc >= DT_MAXXC) { ok = false; break; }
                da.xc_map[da.n_xc] = 1; da.xc_col0[da.n_xc] = c0_; ++da.n_xc;
              }
            auto piece = [&](int in0, int w, int acc0) {
              if (w <= 0) return;
              da.p_in0[da.n_pieces] = in0; da.p_w[da.n_pieces] = w; da.p_acc0[da.n_pieces] = acc0; ++da.n_pieces;
            };
            // the net sees [S | nodes? | Adj^T S | agg_nodes | agg_arcs]; the slot holds [S | Adj^T S | static columns]
            piece(0, D, 0); piece(D + NLq, D, D); piece(D, NLq, sbase); piece(2 * D + NLq, LsM - NLq, sbase + NLq);
            da.partial = dw.partial; da.n_params = dw.n_params; da.bias_off = dw.bias_off;
            da.W = dw.W; da.bnA = dw.bnA; da.bnB = dw.bnB; da.gamma = dw.gamma; da.beta = dw.beta; da.bn_partial = dw.bn_partial;
            da.gate = gate;
            if (ok && in == 2 * D + LsM && dw_tma_finish(da) == GNNFP_OK) {
              if ((rc = launch_dw_tma(da, s, PC_BWD_ITER, &grid_dw))) return rc;
              dw_done = true;
            }
          }
          if (!dw_done && (rc = launch_gemm_dw(dw, s, PC_BWD_ITER, &grid_dw))) return rc;
          grid_state[ty] = grid_dw > grid_state[ty] ? grid_dw : grid_state[ty];
          if (bn) {   // batch sums are complete after dW: constants c0 | c1 | rstd | -mean*rstd of this iteration, BEFORE dX, so that
                      // the dX epilogue can apply dx = a dy - (c0 + x~ c1) while it writes (no correction pass, no gathers later)
            BwdArgs tb;
            memset(&tb, 0, sizeof(tb));
            tb.src = full; tb.net = ndfull; tb.tc.grid = grid_dw; tb.bn_partial = bn_part; tb.gate = gate;
            if ((rc = launch_bn_tail(tb, bn_grad + bg_off[ty], cn_t(t), s, 1, nullptr))) return rc;
          }
          const float* wt = (const float*)(c.ws + L->ws.wtb) + (size_t)ty * L->ws.wtb_stride;
          const int KH = gemm_rows_kpad(H0);
          // one dX launch per gradient destination; two destinations of the same width (dOwn_t and dAgg_t) share a
          // launch: dz is loaded and split once, two accumulators per row tile (gemm.h: nblk == 2)
          GemmRowsArgs dx[GNNFP_MAXP];
          int ndx = 0;
          if (L->xlay) {   // dOwn_t and dAgg_t: one TMA-fed launch (rows_tma.cu); other gradient destinations below
            RowsTmaArgs ra;
            if ((rc = rt_build_dx(c, t, sp, dzbuf, dOwn[wb], dAgg[wb], coef, D % 4 == 0 ? cn_t(t) : nullptr, gate, ra))) return rc;
            if ((rc = launch_rows_tma(ra, s, PC_BWD_DX))) return rc;
          }
          for (int p = 0; p < full.n_pieces; ++p) {
            const Piece& pc = full.p[p];
            const int ldw = gemm_rows_ldw(pc.width);
            if (pc.gptr && !(L->xlay && (pc.tag == TAG_STATE || pc.tag == TAG_AGG_STATE))) {
              GemmRowsArgs& ga = dx[ndx++];
              memset(&ga, 0, sizeof(ga));
              ga.n_rows = full.n_rows; ga.rowlist = full.rowlist; ga.n_pieces = 1;
              gemm_piece_set(ga.p[0], dzbuf, ldG, H0, 0);
              ga.fwd = 0; ga.Kpad = KH; ga.Wp = wt; ga.ldw = ldw; ga.N = pc.width;
              ga.colscale = bn ? coef + 2 * in + pc.col0 : nullptr;
              if (bn) { ga.corr = cn_t(t); ga.corr_in = in; ga.corr_col0 = pc.col0; ga.corr_x = pc.ptr; ga.corr_ld = pc.ld; }
              ga.out = pc.gptr; ga.ld_out = pc.gld; ga.out_add = pc.gmode == GM_ADD;
              ga.vec2 = pc.gld % 2 == 0 && ((uintptr_t)pc.gptr & 7) == 0;
              ga.gate = gate;
            }
            wt += (size_t)KH * ldw;
          }
          static const int no_pair = getenv("GNNFP_NO_DX_PAIR") ? 1 : 0;
          for (int i = 0; i < ndx; ++i) {
            if (dx[i].n_rows < 0) continue;              // already merged into an earlier launch
            for (int j = i + 1; j < ndx && !no_pair && dx[i].nblk == 0; ++j) {
              if (dx[j].n_rows < 0 || dx[j].N != dx[i].N || dx[j].ldw != dx[i].ldw) continue;
              dx[i].nblk = 2;
              dx[i].Wp2 = dx[j].Wp; dx[i].colscale2 = dx[j].colscale;
              dx[i].corr_col02 = dx[j].corr_col0; dx[i].corr_x2 = dx[j].corr_x; dx[i].corr_ld2 = dx[j].corr_ld;
              dx[i].out2 = dx[j].out; dx[i].ld_out2 = dx[j].ld_out; dx[i].out_add2 = dx[j].out_add;
              dx[j].n_rows = -1;
            }
            if ((rc = launch_gemm_rows(dx[i], s, PC_BWD_DX))) return rc;
          }
        }
        for (int si = 0; si < (gemm_bwd[ty] ? 0 : nsplit[ty]); ++si) {
          const Split& sp_ = splits[ty][si];
          BwdArgs ba;
          memset(&ba, 0, sizeof(ba));
          ba.src.n_rows = full.n_rows; ba.src.rowlist = full.rowlist; ba.src.in_dim = sp_.width;
          for (int p = sp_.p0; p < sp_.p1; ++p) { Piece pc = full.p[p]; pc.col0 -= sp_.c_off; ba.src.p[ba.src.n_pieces++] = pc; }
          ba.gsrc.n_rows = full.n_rows; ba.gsrc.rowlist = full.rowlist; ba.gsrc.in_dim = D;
          add_piece(ba.gsrc, mk_direct(dzbuf, ldG, D, 0));
          ba.net = ndfull;
          ba.net.in_dim = sp_.width;
          ba.net.W[0] = ndfull.W[0] + (size_t)sp_.c_off * H0;
          if (ndfull.gamma) { ba.net.gamma = ndfull.gamma + sp_.c_off; ba.net.beta = ndfull.beta + sp_.c_off;
                              ba.net.mmean = ndfull.mmean + sp_.c_off; ba.net.mvar = ndfull.mvar + sp_.c_off; }
          ba.tc.cap_per_row = L->cap_per_row;
          ba.tc.dz_ready = 1;
          // pieces that can be fetched by TMA bulk copies: plain contiguous 16-byte aligned matrices
          auto eligible = [&](const TileSrc& ts2, const Piece& pc) {
            return ts2.rowlist == nullptr && pc.kind == PK_DIRECT && !pc.map && !pc.rowscale && !pc.compact && !pc.gate &&
                   pc.ld == pc.width && (reinterpret_cast<uintptr_t>(pc.ptr) & 15) == 0;
          };
          for (int p = 0; p < ba.src.n_pieces; ++p)
            if (eligible(ba.src, ba.src.p[p])) { ba.tc.bulk_src |= 1u << p; ba.tc.raw_per_row += ba.src.p[p].width; }
          if (eligible(ba.gsrc, ba.gsrc.p[0])) { ba.tc.bulk_g = 1u; ba.tc.raw_per_row += D; }
          for (int p = 0; p < ba.src.n_pieces; ++p) {     // gradients that leave as plain dense matrices: bulk stores
            const Piece& pc = ba.src.p[p];
            if (ba.src.rowlist == nullptr && pc.gmode == GM_STORE && !pc.map && pc.gld == pc.width &&
                (reinterpret_cast<uintptr_t>(pc.gptr) & 15) == 0) { ba.tc.bulk_out |= 1u << p; ba.tc.out_per_row += pc.width; }
          }
          if (tile_cfg_bwd(ba.net, ba.src.n_rows, D, &ba.tc) != GNNFP_OK) {   // does not fit with both buffers
            ba.tc.bulk_out = 0u; ba.tc.out_per_row = 0;
            if (tile_cfg_bwd(ba.net, ba.src.n_rows, D, &ba.tc) != GNNFP_OK) { ba.tc.bulk_src = ba.tc.bulk_g = 0u; ba.tc.raw_per_row = 0; }
          }
          if ((rc = tile_cfg_bwd(ba.net, ba.src.n_rows, D, &ba.tc))) return rc;
          if (ba.tc.grid > L->grid_cap) ba.tc.grid = L->grid_cap;
          grid_state[ty] = ba.tc.grid > grid_state[ty] ? ba.tc.grid : grid_state[ty];
          ba.dz_ready = 1; ba.skip_bias = si > 0;
          ba.partial = part_state + ps_off[ty] + (size_t)sp_.c_off * H0; ba.n_params = L->nparam_s[ty];
          ba.bias_off = (L->snet[ty].in_dim - sp_.c_off) * H0;
          ba.bn_partial = bn_part; ba.bn_in_total = L->snet[ty].in_dim; ba.bn_c_off = sp_.c_off;
          ba.gate = gate; ba.prof_cat = PC_BWD_ITER;
          if ((rc = launch_tile_bwd(ba, s))) return rc;
        }
        // BN tail over the whole net (all splits wrote their columns of bn_partial)
        last_ba.src = full; last_ba.net = ndfull; last_ba.tc.grid = gemm_bwd[ty] ? grid_dw : L->grid_cap; last_ba.tc.cap_per_row = L->cap_per_row;
        last_ba.bn_partial = bn_part; last_ba.gate = gate;
      } else {
        if (pgather) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "partitioned (n_active_rows) backward supports single-Dense-layer net_state only");
        BwdArgs ba;
        memset(&ba, 0, sizeof(ba));
        ba.src = full;
        ba.gsrc.n_rows = ba.src.n_rows; ba.gsrc.rowlist = ba.src.rowlist; ba.gsrc.in_dim = D;
        {
          Piece pa = mk_direct(dSfin, ldG, D, 0);
          if (t < MI) { pa.gate = c.flags() + t; pa.gate_pol = 0; }   // enabled iff iteration t+1 did not run
          add_piece(ba.gsrc, pa);
          if (t < MI) {
            Piece pb = mk_direct(dOwn[rb], ldG, D, 0);
            pb.accumulate = 1; pb.gate = c.flags() + t; pb.gate_pol = 1;
            add_piece(ba.gsrc, pb);
            Piece pc2 = mk_gather(dAgg[rb], ldG, D, 0, g->src_rowptr, g->src_dst, src_w, g->A);
            pc2.accumulate = 1; pc2.gate = c.flags() + t; pc2.gate_pol = 1;
            add_piece(ba.gsrc, pc2);
          }
        }
        ba.net = ndfull;
        ba.tc.cap_per_row = L->cap_per_row;
        if ((rc = tile_cfg_bwd(ba.net, ba.src.n_rows, D, &ba.tc))) return rc;
        grid_state[ty] = ba.tc.grid > grid_state[ty] ? ba.tc.grid : grid_state[ty];
        ba.saved_out = c.S(t); ba.ld_saved = c.ldS(t); ba.saved_compact = 0;
        ba.partial = part_state + ps_off[ty]; ba.n_params = L->nparam_s[ty];
        ba.bn_partial = bn_part;
        ba.gate = gate;
        ba.prof_cat = PC_BWD_ITER;
        if ((rc = launch_tile_bwd(ba, s))) return rc;
        last_ba = ba;
      }
      if (L->snet[ty].has_bn && !gemm_bwd[ty]) {
        const bool inline_bn = dzpath[ty] && !L->composite;
        if (inline_bn) rc = launch_bn_tail(last_ba, bn_grad + bg_off[ty], cn_t(t), s, 1, want ? bn_static : nullptr);
        else rc = launch_bn_tail(last_ba, bn_grad + bg_off[ty], bn_const, s);
        if (rc) return rc;
      }
    }
  }

  if (!(phases & PH_END)) { GNNFP_CHECK_CUDA(cudaGetLastError()); return GNNFP_OK; }
  // ---- 3. input gradients --------------------------------------------------------------------------
  if (want) {
    InGradArgs ia;
    memset(&ia, 0, sizeof(ia));
    ia.N = N; ia.D = D; ia.S = L->S; ia.NLp = NLp; ia.NLw = L->NLw; ia.AL = L->AL; ia.LsM = L->ldXs;   // LsM: row pitch of Xs / dXs
    ia.ldg = ldG; ia.ld_agg1 = c.ldA();
    ia.composite = L->composite; ia.nt = L->nt;
    int o = 0;
    for (int t = 0; t < L->nt; ++t) { ia.dt[t] = L->dt[t]; ia.doff[t] = o; o += L->dt[t]; }
    ia.type_mask = g->type_mask; ia.flags = c.flags();
    ia.dSfin = dSfin; ia.dOwn1 = dOwn[1]; ia.dAgg1 = dAgg[1]; ia.dXs = L->LsM > 0 ? dXs : nullptr;
    ia.src_rowptr = g->src_rowptr; ia.src_dst = g->src_dst; ia.src_w = src_w;
    ia.d_nodes = (want & 1) ? gr->d_nodes : nullptr;
    ia.d_state0 = ((want & 4) && L->S > 0) ? gr->d_state0 : nullptr;
    ia.want = want;
    const bool rt_inline = L->xlay && gemm_bwd[0] && D % 4 != 0;   // own / aggregate corrections left to the consumers (static columns: dX epilogue)
    if (!L->composite && dzpath[0] && L->snet[0].has_bn && (!gemm_bwd[0] || rt_inline)) {
      ia.cn = cn_t(1); ia.csum = rt_inline ? nullptr : bn_static; ia.in_dim = L->snet[0].in_dim;
      ia.s0 = c.S(0); ia.ld0 = c.ldS(0); ia.agg1 = MI > 0 ? c.AGG(1) : nullptr; ia.Xs = c.Xs();
    }
    const size_t tot = (size_t)N * (D > L->NLw ? D : L->NLw);
    int blocks = (int)((tot + 255) / 256);
    if (blocks > 4736) blocks = 4736;
    if (!L->composite && L->S == 0 && MI > 0 && ia.d_nodes && L->NLw == D) {
      // state0 = nodes (GNN.py:259): d_nodes = G_0 = dOwn_1 + Adj . dAgg_1 - the same streaming kernel as the iterations'
      // dz with a linear activation (d_nodes was zeroed above and has no other contribution in this configuration)
      DzArgs da;
      memset(&da, 0, sizeof(da));
      da.n_rows = N; da.D = D; da.act = GNNFP_ACT_LINEAR;
      da.s_t = c.S(0); da.ld_s = c.ldS(0); da.ldg = ldG; da.ld_agg = c.ldA(); da.ld_dz = L->NLw;
      da.dSfin = dSfin; da.dOwn = dOwn[1]; da.dAgg = dAgg[1];
      da.rowptr = g->src_rowptr; da.idx = g->src_dst; da.wgt = src_w;
      da.last_flag = c.flags();
      da.dz = gr->d_nodes;
      if (ia.cn) { da.cn = ia.cn; da.agg_next = c.AGG(1); da.in_dim = ia.in_dim; da.own_col0 = 0; da.agg_col0 = D; }
      if ((rc = launch_dz(da, s))) return rc;
    } else if (ia.d_nodes || ia.d_state0) {
      k_input_grads_nodes<<<blocks, 256, 0, s>>>(ia);
      GNNFP_COUNT_LAUNCH();
    }
    if ((want & 2) && L->AL > 0 && L->A > 0) {
      const int col0 = L->composite ? L->sum_dt : 2 * NLp;
      const size_t ta = (size_t)L->A * L->AL;
      int b2 = (int)((ta + 255) / 256);
      if (b2 > 4736) b2 = 4736;
      k_input_grads_arcs<<<b2, 256, 0, s>>>(L->A, L->AL, L->ldXs, col0, g->dst, g->arc_val, dXs, gr->d_arc_labels,
                                            ia.cn, ia.csum, ia.in_dim, 2 * D + col0, c.Xs(), c.flags());
      GNNFP_COUNT_LAUNCH();
    }
  }

  // ---- 4. parameter gradients ------------------------------------------------------------------------
  for (int ty = 0; ty < L->nt; ++ty) {
    NetDev nd;
    fill_netdev(L->snet[ty], sp[ty], 1, 1, nd);
    if ((rc = launch_reduce_params(nd, part_state + ps_off[ty], grid_state[ty], L->nparam_s[ty], bn_grad + bg_off[ty],
                                   dsp[ty], c.flags(), MI, gr->average_st_grads, s))) return rc;
  }
  if ((rc = launch_reduce_params(ond, part_out, grid_out, L->nparam_o, bn_grad + bg_off[L->nt], *dop, c.flags(), MI, 0, s))) return rc;
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

extern "C" int gnnfp_loop_backward(gnnfp_loop* L, const gnnfp_net_params* sp, const gnnfp_net_params* op,
                                   const gnnfp_loop_io* io, const gnnfp_loop_grads* gr, gnnfp_net_params* dsp,
                                   gnnfp_net_params* dop, void* workspace, size_t workspace_bytes, void* stream) {
  if (L && L->Nact < L->N) GNNFP_FAIL(GNNFP_E_INVALID, "partitioned (n_active_rows) plans use gnnfp_loop_backward_step (halo gradients must be reduced between iterations)");
  return backward_impl(L, sp, op, io, gr, dsp, dop, workspace, workspace_bytes, stream, PH_BEGIN | PH_ITERS | PH_END, 0);
}

// stepping backward: phase 1 = begin (zeroing + net_output backward), 2 = iteration t, 4 = end (input gradients +
// parameter reduction), 8 = gather Adj . dAgg_{t+1} into the exposed buffer (partitioned graphs, before iteration t < max_iteration)
extern "C" int gnnfp_loop_backward_step(gnnfp_loop* L, int32_t phase, int32_t t, const gnnfp_net_params* sp,
                                        const gnnfp_net_params* op, const gnnfp_loop_io* io, const gnnfp_loop_grads* gr,
                                        gnnfp_net_params* dsp, gnnfp_net_params* dop, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  if (phase != PH_BEGIN && phase != PH_ITERS && phase != PH_END && phase != PH_GATHER)
    GNNFP_FAIL(GNNFP_E_INVALID, "loop_backward_step: phase %d (1 begin, 2 iteration, 4 end, 8 gather)", phase);
  if ((phase == PH_ITERS || phase == PH_GATHER) && (!L || t < 1 || t > L->cfg.max_iteration))
    GNNFP_FAIL(GNNFP_E_INVALID, "loop_backward_step: t=%d outside 1..max_iteration", t);
  return backward_impl(L, sp, op, io, gr, dsp, dop, workspace, workspace_bytes, stream, phase, t);
}

extern "C" int gnnfp_loop_bwd_offsets(const gnnfp_loop* L, size_t* gather_off) {
  if (!L) GNNFP_FAIL(GNNFP_E_INVALID, "loop_bwd_offsets: null plan");
  if (!(L->Nact < L->N) || !L->cfg.training) GNNFP_FAIL(GNNFP_E_INVALID, "loop_bwd_offsets: the gather buffer exists in partitioned training plans only");
  if (gather_off) *gather_off = L->ws.pgather;
  return GNNFP_OK;
}
