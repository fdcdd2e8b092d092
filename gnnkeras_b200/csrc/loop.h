// loop.h - the opaque loop plan of gnnfp.h and its workspace layout
#pragma once
#include "graph.h"
#include "gemm.h"

struct WsLayout {
  size_t ctrl = 0, ctrl_bytes = 0;          // int flags[max_iter+1], k, pad | double stats ... (zeroed each forward)
  size_t flags = 0;                         // int [max_iter + 2]
  size_t stS = 0, stA = 0, stX = 0, stO = 0;  // double blocks
  size_t Xs = 0;                            // float [N, LsM]
  size_t slots = 0;                         // float state slots
  size_t agg = 0;                           // float [max_iter][N, D]: Adj^T.state of every iteration (training)
  size_t out_nodes = 0;                     // float [M, T]
  // backward
  size_t dSfin = 0, dOwn = 0, dAgg = 0, dXs = 0, dOutN = 0, dz = 0, pgather = 0;
  size_t part_state = 0, part_out = 0, bn_part = 0, bn_const = 0, bn_grad = 0;
  size_t wfold = 0, wtb = 0, bncoef = 0;    // GEMM path: folded weights per type, W^T blocks, per-iteration BN coefficients
  size_t wfold_stride = 0, wtb_stride = 0, bncoef_stride = 0;   // floats per type
  size_t arc_tmp = 0;                       // arc focus: per-arc gradient of net_output's gathered input [A][2 (D + NL)]
  size_t bncoef_t = 0;                      // X-slot path: [max_iteration][bncoef_stride] backward BN coefficients left by the forward kernels
  size_t bn_const_t = 0, bn_static = 0;     // per-iteration BN constants, running static-column sums
  size_t bwd_zero = 0, bwd_zero_bytes = 0;   // region zeroed at the start of every backward
  size_t total = 0;
};

struct gnnfp_loop {
  const gnnfp_graph* g = nullptr;
  gnnfp_loop_cfg cfg{};
  gnnfp_net_desc snet[GNNFP_MAX_TYPES]{};
  gnnfp_net_desc onet{};
  int nt = 1;            // state nets
  int composite = 0;
  int N = 0, A = 0, M = 0, D = 0, S = 0, NLw = 0, AL = 0, T = 0;
  int Nact = 0;          // rows net_state runs on (N, or the owned block of an edge-cut partition)
  int dt[GNNFP_MAX_TYPES]{};   // clamped d_t (composite)
  int sum_dt = 0;
  int LsM = 0;           // materialised static block width
  int out_in = 0;        // net_output input width
  int out_rows = 0, pool = 0, n_out_rows_in = 0;   // n_out_rows_in: rows fed to net_output (M)
  int slot_count = 0;
  int bn_train_state = 0, bn_train_out = 0;
  int nparam_s[GNNFP_MAX_TYPES]{}, nparam_o = 0;
  int bwd_grid_state[GNNFP_MAX_TYPES]{};   // backward: widest partial-slot grids used since the last begin phase
  int bwd_grid_out = 0;
  int out_gemm_ok = 0;              // net_output runs on the GEMM kernels (single Dense layer, node / graph focus)
  int gemm_ok[GNNFP_MAX_TYPES]{};   // single Dense layer nets run the pipelined GEMM kernels (gemm.cu)
  // TMA path (rows_tma.cu): homogeneous single-Dense-layer state nets keep every iteration's state and aggregate
  // interleaved in ONE row-major slot per iteration,  X_t = [S_t (D) | Adj^T S_t (D) | static columns (inline, optional)],
  // with 16-byte aligned row pitches so that the tensor-map TMA unit can load / store every matrix of the loop
  int xlay = 0;          // interleaved layout + rows_tma kernels in use
  int ldX = 0;           // floats per row of an X slot
  int ldG = 0;           // leading dimension of dz / dOwn / dAgg / dSfin  (D unless xlay)
  int ldXs = 0;          // leading dimension of the static block Xs / dXs  (LsM unless xlay)
  int fuse_agg = 0;      // the forward iteration kernel also produces Adj^T S_t for the next iteration (in-tile gather)
  int xs_inline = 0;     // static columns are copied into every X slot (few columns: saves one operand chunk per tile)
  int cap_per_row = 4;   // CSR scratch capacity per tile row (from A/N)
  int grid_cap = 0;      // upper bound of any backward tile kernel grid (partials are sized by it)
  WsLayout ws;
};

// ---- per-call context (loop.cu / loop_bwd.cu) ---------------------------------------------------
struct Ctx {
  gnnfp_loop* L;
  const gnnfp_loop_io* io;
  char* ws;
  cudaStream_t s;
  int* flags() const { return (int*)(ws + L->ws.flags); }
  float* Xs() const { return (float*)(ws + L->ws.Xs); }
  float* slots() const { return (float*)(ws + L->ws.slots); }
  size_t slot_stride() const { return ((size_t)L->N * (L->xlay ? L->ldX : L->D) + 31) / 32 * 32; }   // 128-byte aligned slots
  const float* S0user() const { return L->S > 0 ? io->state0 : io->nodes; }   // the caller's initial state
  int ldS0user() const { return L->S > 0 ? L->S : io->ld_nodes; }
  int xslot(int t) const { return L->cfg.training ? t : (t & 1); }            // xlay: slot of iteration t (0 = the copy of the initial state)
  const float* S(int t) const {   // state after t iterations
    if (L->xlay) return slots() + (size_t)xslot(t) * slot_stride();
    if (t == 0) return S0user();
    return slots() + (L->cfg.training ? (size_t)(t - 1) : (size_t)(t & 1)) * slot_stride();
  }
  float* AGG(int t) const {       // Adj^T S_{t-1}, t = 1..max_iter (one slot in inference)
    if (L->xlay) return slots() + (size_t)xslot(t - 1) * slot_stride() + L->D;
    return (float*)(ws + L->ws.agg) + (L->cfg.training ? (size_t)(t - 1) : (size_t)0) * slot_stride();
  }
  int ldS(int t) const { return L->xlay ? L->ldX : (t == 0 ? ldS0user() : L->D); }
  int ldA() const { return L->xlay ? L->ldX : L->D; }
  int stXw() const {
    if (!L->composite) return L->LsM;
    int m = 0;
    for (int t = 0; t < L->nt; ++t) m = L->dt[t] > m ? L->dt[t] : m;
    return m + L->sum_dt + L->AL;
  }
  double* stS(int ty, int t) const { return (double*)(ws + L->ws.stS) + ((size_t)ty * (L->cfg.max_iteration + 1) + t) * 2 * L->D; }
  double* stA(int ty, int t) const { return (double*)(ws + L->ws.stA) + ((size_t)ty * (L->cfg.max_iteration + 1) + t) * 2 * L->D; }
  double* stX(int ty) const { return (double*)(ws + L->ws.stX) + (size_t)ty * 2 * stXw(); }
  double* stO() const { return (double*)(ws + L->ws.stO); }
};

Piece mk_direct(const float* ptr, int ld, int width, int col0);
Piece mk_gather(const float* ptr, int ld, int width, int col0, const int* rowptr, const int* idx, const float* wgt, int nnz = 0);
void add_piece(TileSrc& ts, const Piece& p);
void build_state_src(const Ctx& c, int ty, int t, TileSrc& ts, int agg_direct);
void build_out_src(const Ctx& c, TileSrc& ts);
void fill_netdev(const gnnfp_net_desc& d, const gnnfp_net_params& p, int training, int n_rows, NetDev& nd);
int check_io(const gnnfp_loop* L, const gnnfp_loop_io* io, void* workspace, size_t workspace_bytes);
int check_params(const gnnfp_net_desc& d, const gnnfp_net_params& p, const char* what);
int launch_bn_tail(const BwdArgs& a, float* bn_grad, float* bn_const, cudaStream_t s, int no_fix = 0, float* static_acc = nullptr);
int launch_reduce_params(const NetDev& net, const float* partial, int grid, int n_params, const float* bn_grad,
                         const gnnfp_net_params& d, const int* flags, int max_iter, int average, cudaStream_t s);
