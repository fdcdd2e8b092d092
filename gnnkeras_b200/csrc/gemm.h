// gemm.h - argument structs and launchers of the pipelined GEMM path (gemm.cu)
#pragma once
#include "common.cuh"

#define GEMM_BK 8
#define GEMM_STAGES 3
#define GEMM_MAXP 4

struct GemmPiece { const float* ptr; int ld; int width; int k8; int al8; };   // k8 = first padded K index (even); al8: 8-byte copies legal
void gemm_piece_set(GemmPiece& g, const float* ptr, int ld, int width, int k8);

struct GemmRowsArgs {
  int n_rows;
  const int* rowlist;        // global row id of row r (NULL = identity)
  int n_pieces;
  GemmPiece p[GEMM_MAXP];
  int fwd;                   // 1: forward epilogue (bias, act, convergence, statistics); 0: backward dX epilogue (colscale, out_add)
  int vec2;                  // out / prev rows may be accessed as aligned float2 (8-byte aligned bases, even leading dims)
  int Kpad;                  // padded K: pieces start at even offsets k8, total rounded up to a multiple of 8
  const float* Wp;           // [Kpad][ldw] padded weights (row k8+kk of piece p), zero rows in the padding
  int ldw;                   // == 16*ceil(N/16)
  int N;                     // real output columns
  const float* bias;         // [N] or NULL
  const float* colscale;     // [N] or NULL: epilogue multiply (backward: gamma*rstd)
  int act;
  float* out; int ld_out; int out_add;   // out[gr*ld_out + j] (= or +=)
  int out_compact;           // output row = position in the row set instead of the global row id
  int a_compact;             // input (piece) rows are addressed by position in the row set (a compact dz matrix)
  // backward epilogue, BN-training correction folded in: out -= c0 + (x*A + B)*c1 with corr = [c0|c1|A|B] x corr_in,
  // column corr_col0 + j, x = corr_x[gr*corr_ld + j]
  const float* corr; int corr_in; int corr_col0; const float* corr_x; int corr_ld;
  // second output block of the backward epilogue (nblk == 2): the same A operand (dz) against a second weight block of
  // the same shape, written to its own destination - dOwn and dAgg of one iteration in ONE launch.  launch_gemm_rows
  // splits it into two launches where the tensor-core kernel cannot take both.
  int nblk;
  const float* Wp2; const float* colscale2; int corr_col02; const float* corr_x2; int corr_ld2;
  float* out2; int ld_out2; int out_add2;
  const float* prev; int ld_prev; float thr; int* flag_next;   // convergence epilogue (forward) or NULL
  double* ost_sum; double* ost_sq;                            // output column statistics or NULL
  const int* gate;
};


struct GemmDwArgs {
  int n_rows;
  const int* rowlist;
  int n_pieces;
  GemmPiece p[GEMM_MAXP];
  int Kp;                    // padded K (even; pieces at even offsets k8)
  const float* dz; int ld_dz; int H; int dz_compact;   // dz_compact: dz rows by position in the row set
  float* partial;            // [grid][n_params]: dW at (real k index)*H + j, db at bias_off + j
  int n_params; int bias_off;
  // BN: per-CTA sums  P_c = sum_j W[c][j] db[j],  Q_c = rstd*(sum_j W[c][j] acc[c][j]) - mean*rstd*P_c
  const float* W; const float* bnA; const float* bnB; const float* gamma; const float* beta;   // bnA/bnB: [K] rstd, -mean*rstd
  float* bn_partial;         // [grid][2*K]
  const int* gate;
};


// narrow.cu: net_output with a single Dense layer of H <= 4 columns as streaming kernels (warp per row, lane = input column)
struct NarrowArgs {
  int n_rows; const int* rowlist;
  int n_pieces; GemmPiece p[GEMM_MAXP];    // input pieces (k8 = row offset inside the padded weights Wp)
  int K, H;                                // input columns, output columns
  // forward: out[r] = act(x . Wp + bias), rows compact
  const float* Wp; int ldw; const float* bias; int act; float* out; int ld_out;
  // backward: dz compact [n_rows, H]
  const float* dz;
  float* partial; int n_params, bias_off;  // dW: as GemmDwArgs
  const float* W; const float* bnA; const float* bnB; const float* gamma; const float* beta; float* bn_partial;
  const float* colscale; const float* corr; int corr_in;       // dX: per input column gamma * rstd, BN constants [c0|c1|A|B] x corr_in
  float* gout[GEMM_MAXP]; int gld[GEMM_MAXP]; int gadd[GEMM_MAXP];   // per piece: gradient destination (NULL = none), leading dimension, += or =
};
int narrow_supported(const NarrowArgs& a);
int launch_narrow_fwd(const NarrowArgs& a, cudaStream_t s, int prof_cat);
int launch_narrow_dw(const NarrowArgs& a, cudaStream_t s, int prof_cat, int* grid_out);
int launch_narrow_dx(const NarrowArgs& a, cudaStream_t s, int prof_cat);

struct FoldArgs {            // padded, BN-folded weights of a single Dense layer for gemm_rows (forward)
  TileSrc src;               // the net's input pieces (their st_sum/st_sq feed the BN batch statistics)
  NetDev net;
  int k8[GNNFP_MAXP];
  int Kpad, ldw;
  float* Wp;                 // [Kpad][ldw]
  float* biasp;              // [ldw]
  int update_moving;
  float* coef_out;           // block 0 (training, batch statistics): [3][in] = rstd | -mean*rstd | gamma*rstd for the backward, or NULL
  const int* gate;
};
struct BnCoefArgs { TileSrc src; NetDev net; float* coef; const int* gate; };
int launch_fold_w(const FoldArgs& a, cudaStream_t s);
int launch_bn_coef(const BnCoefArgs& a, cudaStream_t s);
int launch_gemm_rows(const GemmRowsArgs& a, cudaStream_t s, int prof_cat);
int gemm_rows_tc_supported(const GemmRowsArgs& a);                      // gemm_tc.cu: tcgen05 (3xTF32) version of the same GEMM
int launch_gemm_rows_tc(const GemmRowsArgs& a, cudaStream_t s, int prof_cat);
int gemm_rows_ldw(int N);
int gemm_rows_kpad(int k8);
int gemm_rows_supported(int k8, int N);
int launch_gemm_dw(const GemmDwArgs& a, cudaStream_t s, int prof_cat, int* grid_out);
int gemm_dw_tc_supported(const GemmDwArgs& a);                          // gemm_tc.cu: tcgen05 (3xTF32) version
int launch_gemm_dw_tc(const GemmDwArgs& a, cudaStream_t s, int prof_cat, int* grid_out);
int gemm_dw_supported(int Kp, int H);
int gemm_dw_grid(int n_rows);
int launch_transpose_block(const float* W, int H, int col0, int width, int Kpad, int ldw, float* out, cudaStream_t s);
