// common.cuh - shared types of libgnnfp (sm_100a).  See DESIGN.md for the kernel inventory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "gnnfp.h"

#define GNNFP_MAXP 6        // max input pieces of one net application
#define GNNFP_JC 16         // output columns per thread chunk in the tile MLP
#define GNNFP_NSM_FALLBACK 148
#define GNNFP_TILE_ROWS 128       // row tile of the TMA kernels (rows_tma.cu) and of the tile-local CSR view (graph.cu)
#define GNNFP_TILE_ARCS 1024      // in-arcs of one tile the fused gather can stage; tiles above are handled by the row-list pass

// ---- error plumbing (never throw across the C ABI) -------------------------------------------
void gnnfp_set_error(const char* fmt, ...);
#define GNNFP_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      gnnfp_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                      cudaGetErrorString(_e));                                              \
      return GNNFP_E_CUDA;                                                                  \
    }                                                                                       \
  } while (0)
#define GNNFP_FAIL(code, ...)      \
  do {                             \
    gnnfp_set_error(__VA_ARGS__);  \
    return (code);                 \
  } while (0)

extern long long g_gnnfp_launches;
#define GNNFP_COUNT_LAUNCH() (++g_gnnfp_launches)

// optional per-category kernel timing with CUDA events on the launch stream (bench.py's roofline leg)
enum { PC_OTHER = 0, PC_FWD_ITER = 1, PC_BWD_ITER = 2, PC_PASS = 3, PC_FWD_OUT = 4, PC_BWD_OUT = 5, PC_BNFIX = 6,
       PC_DZ = 7, PC_BWD_DX = 8, PC_AGG = 9, PC_COUNT = 10 };
// optional per-phase cycle counters inside the tile kernels (debug; thread 0 of every CTA, clock64)
#ifdef GNNFP_PHASE_TIMING
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long _t = clock64(); atomicAdd((unsigned long long*)&g_phase_cycles[i], (unsigned long long)(_t - _pt)); _pt = _t; } } while (0)
#define PHASE_INIT() long long _pt = clock64()
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_INIT() do { } while (0)
#endif
extern int g_gnnfp_prof;
void gnnfp_prof_begin(int cat, cudaStream_t s);
void gnnfp_prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s; bool on;
  ProfScope(int cat, cudaStream_t st) : s(st), on(g_gnnfp_prof != 0) { if (on) gnnfp_prof_begin(cat, s); }
  ~ProfScope() { if (on) gnnfp_prof_end(s); }
};

// ---- a "piece": one column block of a net's input, read straight from where the data lives ---
// The reference materialises tf.concat([...]) every iteration (GNN.py:231); here the concat only
// ever exists as a shared-memory tile.
enum { PK_DIRECT = 0, PK_GATHER = 1 };
enum { GM_NONE = 0, GM_STORE = 1, GM_ADD = 2, GM_ATOMIC = 3 };
enum { TAG_NONE = 0, TAG_STATE = 1, TAG_NODES = 2, TAG_AGG_STATE = 3, TAG_STATIC = 4, TAG_ARC_LABELS = 5 };

// flat index / width with magic = ceil(2^32 / width); width 1 makes the 32-bit magic wrap to 0 -> identity
__device__ __forceinline__ int div_magic(unsigned e, unsigned magic) { return magic ? (int)__umulhi(e, magic) : (int)e; }

struct Piece {
  const float* ptr;        // source matrix (row-major)
  int ld;                  // its leading dimension
  int width;               // columns taken
  int col0;                // first column in the concatenated net input
  int kind;                // PK_DIRECT | PK_GATHER
  int accumulate;          // staging adds into the tile instead of overwriting (sum of pieces)
  int compact;             // DIRECT: source row = position in the row set (not the global row id)
  int tag;                 // TAG_*: what the columns are (host-side bookkeeping for the backward)
  unsigned magic;          // ceil(2^32/width) for the flat-index division
  const int* map;          // DIRECT: source row = map[gr] (arc focus: src/dst of the arc; un-pool: node2graph)
  const float* rowscale;   // DIRECT: value *= rowscale[gr]  (un-pooling by NodeGraph values)
  const int* rowptr;       // GATHER: CSR over the global row id gr
  const int* idx;          //         source row of each entry
  const float* wgt;        //         weight of each entry (NULL = 1)
  int nnz;                 //         entries in idx/wgt (bounds the next-tile prefetch)
  const double* st_sum;    // BN batch statistics of these columns (sum over rows) or NULL
  const double* st_sq;     //                                     (sum of squares)
  const int* gate;         // piece enabled iff gate==NULL || ((*gate != 0) == gate_pol)
  int gate_pol;
  // backward: where the gradient w.r.t. these columns goes
  float* gptr;
  int gld;
  int gmode;               // GM_*
  int gdirect;             // the gradient row is the row's own id even when the VALUES are read through `map`
};

struct TileSrc {
  int n_rows;              // rows of the row set
  const int* rowlist;      // global row id of row r (NULL = identity)
  int n_pieces;
  int in_dim;
  Piece p[GNNFP_MAXP];
};

struct NetDev {
  int n_layers;
  int in_dim;
  int widths[GNNFP_MAX_LAYERS];
  int acts[GNNFP_MAX_LAYERS];
  const float* W[GNNFP_MAX_LAYERS];
  const float* b[GNNFP_MAX_LAYERS];
  int bn_mode;             // 0 none, 1 batch statistics (training), 2 moving statistics
  float bn_eps, bn_momentum;
  const float* gamma;
  const float* beta;
  float* mmean;
  float* mvar;
  double inv_n;            // 1 / rows of the BN batch
};

// ---- tile geometry chosen on the host ------------------------------------------------------------
struct TileCfg {
  int RG;                  // row-group warps
  int CG;                  // column-group warps
  int R;                   // rows per tile = 64 * RG  (2 rows per lane)
  int XS0, XS1;            // row strides (% 8 == 4) of the two activation buffers
  int cap;                 // arcs of one tile the CSR scratch can hold (fast gather path)
  int cap_per_row;         // in: scratch capacity per tile row (0 = default 4), set by the caller from A/N
  int dz_ready;            // in (backward): the launch consumes a precomputed dz (no saved-output tile)
  int raw_per_row;         // in: floats per tile row of the TMA bulk-copy landing buffer (0 = synchronous staging only)
  unsigned bulk_src, bulk_g; // in: bit p set = piece p of src / gsrc is staged by cp.async.bulk (full tiles)
  unsigned bulk_out;       // in (backward): bit p set = gradient of src piece p leaves through a bulk store
  int out_per_row;         // in: floats per tile row of the bulk-store staging buffer
  int threads;
  size_t smem_bytes;
  int grid;
};

struct FwdArgs {
  TileSrc src;
  NetDev net;
  TileCfg tc;
  float* out;              // output rows
  int ld_out;
  int out_compact;         // 1: row r of the row set -> out[r]; 0: -> out[gr]
  double* ost_sum;         // column statistics of the output (next iteration's BN) or NULL
  double* ost_sq;
  const float* prev;       // convergence test against this matrix (rows gr), or NULL
  int ld_prev;
  int prev_col0;           // column of the raw previous state inside the staged input tile (-1: read `prev`)
  float thr;
  int* flag_next;          // set to 1 when any row is not converged
  const int* gate;         // whole kernel runs only if *gate != 0 (NULL = always)
  int update_moving;       // CTA 0 applies the Keras moving-average update
  float* agg_out;          // optional: save tile columns [agg_col0, agg_col0+agg_w) (Adj^T.state) for the backward
  int agg_col0, agg_w;
  int prof_cat;
};

struct PassArgs {           // tile pass without a net: materialise pieces and/or column statistics
  TileSrc src;
  TileCfg tc;
  float* out;              // [n_rows, in_dim] or NULL
  int ld_out;
  double* st_sum;          // [in_dim] or NULL
  double* st_sq;
  const int* gate;
  int prof_cat;
};

struct BwdArgs {
  TileSrc src;             // the net's input pieces (with gradient destinations)
  TileSrc gsrc;            // pieces that assemble dL/d(out) of this application (width = last layer)
  NetDev net;
  TileCfg tc;
  const float* saved_out;  // the forward output rows (rows gr or compact), for act' of the last layer
  int ld_saved;
  int saved_compact;
  float* partial;          // [grid, n_params] per-CTA partial sums, accumulated across launches
  int n_params;
  float* bn_partial;       // [grid, 2*bn_in_total] per-CTA sum(dy), sum(dy*x~) of THIS launch
  const int* gate;
  int prof_cat;
  // column-split mode (single Dense layer): `gsrc` already holds dz = G * act'(s_t) (dz_kernel), this launch
  // owns input columns [c_off, c_off + net.in_dim) of the full net
  int dz_ready;
  int skip_bias;           // only one split accumulates db
  int bias_off;            // offset of the bias block inside this launch's (shifted) partial slot
  int bn_in_total;         // in_dim of the full net (bn_partial row layout); 0 = net.in_dim
  int bn_c_off;
};

struct DzArgs {             // dz[i] = act'(s_t[i]) * G_t[i],  G_t = last ? dSfin : dOwn + Adj . dAgg
  int n_rows;
  const int* rowlist;
  int D, act;
  const float* s_t;        // saved output rows (by global row id)
  int ld_s;
  const float* dSfin;
  const float* dOwn;
  const float* dAgg;
  const int* rowptr;       // source-grouped CSR
  const int* idx;
  const float* wgt;
  const int* last_flag;    // G = dSfin iff last_flag == NULL || *last_flag == 0
  int always_last;
  float* dz;               // [N, D] by global row id
  int ldg;                 // leading dimension of dSfin / dOwn / dAgg / dz / pre (0 = D)
  int ld_agg;              // leading dimension of agg_next (0 = D)
  int ld_dz;               // leading dimension of dz (0 = ldg)
  const float* pre;        // optional [N, D]: Adj . dAgg already gathered (partitioned graphs) - replaces the CSR gather
  const int* gate;
  // BN training (homogeneous nets): constants [c0|c1|rstd|-mean*rstd] x in_dim of iteration t+1 and its saved Adj^T.s
  const float* cn;
  const float* agg_next;
  int in_dim, own_col0, agg_col0;
};
int launch_dz(const DzArgs& a, cudaStream_t s);

struct AggArgs {            // AGG = Adj^T . S (dst-CSR gather) as a streaming kernel + fp64 column statistics
  int n_rows;
  int pad4;                // set by the launcher: padded float4 walk (D % 4 == 2)
  const int* n_rows_dev;   // optional device scalar overriding n_rows (then n_rows is only the upper bound that sizes the grid)
  const int* rowlist;
  int D;
  const float* S; int ld;
  const int* rowptr; const int* idx; const float* wgt;   // rowptr NULL = direct rows (entry of row r is S[r])
  float* out;              // [N, D] by global row id (may be NULL)
  int ld_out;              // leading dimension of out (0 = D)
  double* st_sum; double* st_sq;   // [D] each (may be NULL)
  const int* gate;
};
int launch_agg_stats(const AggArgs& a, cudaStream_t s, int prof_cat = 0);

// launchers (kernels.cu)
int launch_tile_fwd(const FwdArgs& a, cudaStream_t s);
int launch_tile_pass(const PassArgs& a, cudaStream_t s);
int launch_tile_bwd(const BwdArgs& a, cudaStream_t s);
int tile_cfg_fwd(const NetDev& net, int n_rows, TileCfg* tc);
int tile_cfg_pass(int in_dim, int n_rows, TileCfg* tc);
int tile_cfg_bwd(const NetDev& net, int n_rows, int gwidth, TileCfg* tc);
int net_param_count(const gnnfp_net_desc& d);   // W,b of all layers (+ 2*in_dim for BN gamma/beta at the end)

int gnnfp_num_sms();
