// rows_tma.h - argument block of the TMA-fed row GEMM (rows_tma.cu): the fused forward iteration and the backward dX
#pragma once
#include <cuda.h>            // CUtensorMap (type only: the encoder is resolved through cudaGetDriverEntryPoint, libcuda is not linked)

#include "common.cuh"

#define RT_ROWS 128          // rows per tile = MMA M
#define RT_CHUNK 32          // fp32 columns per stage = one 128-byte swizzle row
#define RT_STAGE_BYTES (RT_ROWS * 128)
#define RT_MAXKC 12          // K chunks (A operand stages per tile)
#define RT_MAXOC 8           // output chunks per tile
#define RT_MAPS 8
#define RT_MAXIN 384         // input columns of the net (RT_MAXKC * 32)
#define RT_MAXBN 224         // accumulator columns per tile (two buffers + one chunk of slack inside 512 TMEM columns)

enum { RT_FWD = 0, RT_DX = 1 };

struct RtKChunk {            // A operand chunk: columns [col0, col0 + 32) of matrix maps[map]
  int map, col0;
  int k8;                    // 8-wide K steps that carry data (1..4)
  int width;                 // valid columns (<= 32)
  short wrow[RT_CHUNK];      // column k of the chunk multiplies row wrow[k] of W (FWD: input column; DX: dz column), -1 = none
};
struct RtOChunk {            // output chunk: accumulator columns [acc_col0, acc_col0 + 32) -> columns [out_col0, ..) of maps[out_map]
  int acc_col0, out_map, out_col0;
  int aux_map, aux_col0;     // side input preloaded into the output stage (-1 = none): FWD previous state, DX the BN input x
  int cidx0;                 // FWD: first output column (bias index);  DX: input column index of the first column (constants)
  int width;                 // valid columns (<= 32)
  int st_slot;               // FWD: index of this chunk among the output's chunks (statistics registers)
};

struct RowsTmaArgs {
  CUtensorMap maps[RT_MAPS];
  int n_rows;
  int n_kc; RtKChunk kc[RT_MAXKC];
  int n_oc; RtOChunk oc[RT_MAXOC];
  int mode;                  // RT_FWD | RT_DX
  int BN;                    // MMA N: accumulator columns per tile (multiple of 16)
  unsigned bn_magic;         // ceil(2^32 / BN): division by BN as a multiplication (set by rows_tma_finish)
  int tmem_cols;             // power of two >= 2 * BN + 32
  int n_stages;              // operand ring depth (hi + lo tile per stage)
  int n_ostages;             // output stage ring depth
  int H;                     // Dense width (columns of W)
  // ---- RT_FWD: s_t = act(BN(x) W + b); BN folded into the shared-memory weights by every CTA ----------------------
  TileSrc src;               // the net's input pieces (BN batch statistics per piece, column layout of W's rows)
  NetDev net;
  int update_moving;         // CTA 0 applies the Keras moving-average update
  float* coef_out;           // CTA 0: [3][in] = rstd | -mean*rstd | gamma*rstd for the backward of this iteration (or NULL)
  int act;
  float thr;                 // convergence test against the aux (previous state) chunk; flag_next NULL = no test
  int* flag_next;
  double* ost_sum; double* ost_sq;    // column statistics of the output (next iteration's BN) or NULL
  // fused aggregation (reference GNN.py:228 of the NEXT iteration): Adj^T S_t of every row whose in-neighbours all lie in
  // its 128-row tile is gathered from the finished output stage in shared memory (the graph handle's tile-local CSR view);
  // the remaining rows (graph.cu: bnd_rows) are left to the row-list pass of agg_stats_kernel
  int fuse_agg;
  const int* g_rowptr; const short* g_lidx; const float* g_w; const int* g_arc0;
  float* agg_out; int ld_agg;         // Adj^T S_t [row][c] = agg_out[row * ld_agg + c]
  double* agg_sum; double* agg_sq;    // its column statistics or NULL
  // ---- RT_DX: dX = dz (W^T * colscale) - BN-training correction ------------------------------------------------------
  const float* W;            // [in][H] Dense kernel
  const float* colscale;     // [in] gamma * rstd (NULL = 1)
  const float* corr;         // [4][corr_in] = c0 | c1 | A | B of this iteration (NULL = none): out -= c0 + (x*A + B)*c1
  int corr_in;
  int n_blk; int blk_acc0[2], blk_in0[2], blk_w[2];   // output blocks: accumulator column / input column / width
  const int* gate;
};

// dw_tma.cu: dW = X^T dz with both operands read as MN-major TMA boxes (no transposing pass)
#define DT_MAXXC 8
#define DT_MAXP 4
struct DwTmaArgs {
  CUtensorMap xmap[2];       // X sources: [n_rows x cols] row-major, box 32 columns x 32 rows
  CUtensorMap zmap;          // dz [n_rows x H]
  int n_rows, H, K;          // K = input columns of the net (rows of W)
  int n_xc; int xc_map[DT_MAXXC], xc_col0[DT_MAXXC];   // X chunk i = columns [col0, col0 + 32) of xmap[map] -> accumulator columns [32 i, 32 i + 32)
  int n_zc;                  // dz chunks = ceil(H / 32)
  int n_pieces; int p_in0[DT_MAXP], p_w[DT_MAXP], p_acc0[DT_MAXP];   // input columns [in0, in0 + w) sit in accumulator columns [acc0, ..)
  int n_stages, n_lo;        // hi ring depth, lo slots
  int rows;                  // rows per stage / TMA box (32, 64, 128): dw_tma_rows()
  unsigned h_magic;          // ceil(2^32 / H) (set by dw_tma_finish)
  float* partial; int n_params, bias_off;               // as GemmDwArgs
  const float* W; const float* bnA; const float* bnB; const float* gamma; const float* beta;
  float* bn_partial;
  const int* gate;
};
int dw_tma_rows(int n_xc, int n_zc);
int dw_tma_finish(DwTmaArgs& a);                                 // ring depth from the shared-memory budget; error if the shape does not fit
int launch_dw_tma(const DwTmaArgs& a, cudaStream_t s, int prof_cat, int* grid_out);

int rows_tma_available();                                        // the driver exports cuTensorMapEncodeTiled
int rows_tma_map(CUtensorMap* m, const float* ptr, int rows, int cols, int ld, int box_rows = RT_ROWS, int atom32 = 0);   // atom32: SWIZZLE_128B_ATOM_32B   // [rows, cols] fp32 row-major, box 32 x 128, SWIZZLE_128B
int rows_tma_ok(const float* ptr, int ld);                       // 16-byte aligned base and row pitch
size_t rows_tma_smem(const RowsTmaArgs& a);
int rows_tma_finish(RowsTmaArgs& a);                             // derive tmem_cols / ring depths from the shared-memory budget; error if it does not fit
int launch_rows_tma(const RowsTmaArgs& a, cudaStream_t s, int prof_cat);
// agg_tile.cu: Adj^T . S as a TMA-pipelined tile kernel over the graph's tile-local CSR view
int agg_tile_supported(const float* S, int ld_s, int D);
int launch_agg_tile(const float* S, int ld_s, int n_rows, int D, const int* rowptr, const int* src, const float* wgt,
                    const short* lidx, const int* arc0, float* out, int ld_out, double* st_sum, double* st_sq,
                    const int* gate, cudaStream_t s, int prof_cat);
