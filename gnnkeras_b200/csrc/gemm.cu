// gemm.cu - software-pipelined FP32 GEMM kernels for single-Dense-layer nets (the reference's default MLP):
//
//   gemm_rows_kernel<TM,TN> : OUT[rows, N] = epilogue( [piece_0 | piece_1 | ...][rows, K] . Wp[K, N] )
//        forward  : one fixed-point iteration  s_t = act([s | nodes? | Adj^T s | static] . (a (.) W) + b')  with the
//                   convergence test (GNN.py:200-214) and the next BN's column statistics in the epilogue;
//        backward : dX = dz . W^T per destination block (dOwn, dAgg, static columns), scaled by gamma*rstd.
//   gemm_dw_kernel<TC,TJ>   : dW[K, H] (+ db) = X^T . dz over this CTA's rows -> per-CTA partial slot.
//
// Both stream 8-wide K chunks (resp. 8-row chunks) through a 3-stage cp.async ring (LDGSTS, zero-filling tails),
// keep the output tile in registers (TM x TN per thread, 256 threads = 16 x 16 thread grid) and run as persistent
// CTAs over consecutive tiles, so global latency is hidden behind the FMAs instead of being exposed per phase.
// All operands are plain row-major matrices: the sparse aggregation is materialised by agg_stats_kernel first.
#include <stdlib.h>

#include "tile.cuh"

#include "gemm.h"

__device__ __forceinline__ void cp_async4(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// gemm_rows: the whole padded weight matrix stays resident in shared memory for the life of the (persistent) CTA; only
// the activation rows stream, GR_BK columns per stage.  Thread (ty, tx) owns rows ty + 16*i (adjacent rows of a warp sit
// 4 banks apart: conflict-free LDS.128 broadcasts) and, for N > 48, the column quad 4*tx..4*tx+3 (one LDS.128 of W per
// k-step) plus the single columns 64 + 16*m + tx.
#define GR_BK 32
#define GR_AS 36        // floats per A row in shared memory (32 + 4 pad: 144 B rows, 16-byte aligned)
#define GR_STAGES 3

template <int TN>
__device__ __forceinline__ int gr_col(int m, int tx) {
  if (TN >= 4) return m < 4 ? 4 * tx + m : 64 + 16 * (m - 4) + tx;
  return tx + 16 * m;
}

__device__ __forceinline__ float selu_fwd(float z) {   // same arithmetic as act_fwd(SELU), without the branch
  const float e = expf(fminf(z, 0.0f));
  return z < 0.0f ? (SELU_SCALE_F * SELU_ALPHA_F) * (e - 1.0f) : SELU_SCALE_F * z;
}

// FWD = true : out = act(x.Wp + bias), convergence test against `prev`, output column statistics
// FWD = false: out (+)= (x.Wp) * colscale                                   (backward dX blocks)
template <int TM, int TN, bool FWD>
__global__ void __launch_bounds__(256, 2) gemm_rows_kernel(const __grid_constant__ GemmRowsArgs a) {
  if (a.gate && *a.gate == 0) return;
  constexpr int BM = 16 * TM, BN = 16 * TN;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                                   // [Kpad][BN]
  float* As = smem + a.Kpad * BN;                     // [GR_STAGES][BM][GR_AS]
  __shared__ double colacc[2][BN];
  __shared__ float sbias[BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n = a.n_rows;
  const int n_tiles = (n + BM - 1) / BM;
  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int tile1 = min(n_tiles, tile0 + tiles_per_cta);
  const int my_tiles = max(0, tile1 - tile0);
  const int NS = (a.Kpad + GR_BK - 1) / GR_BK;        // stages per tile (the last one may be partial)

  // ---- producer: stage `it` = (tile, 32-column group); this thread copies column pair pq of rows ty + 16*i -------
  const int pq = tid & 15;
  const int* arl = a.a_compact ? nullptr : a.rowlist;  // row list of the input pieces
  int i_tq = 0, i_sg = 0, i_st = 0;                  // next stage to issue: tile, column group, ring slot
  auto issue = [&]() {
    if (i_tq < my_tiles) {
      const int tq = i_tq, sg = i_sg;
      float* dst = As + i_st * BM * GR_AS + ty * GR_AS + 2 * pq;
      const int kcol = sg * GR_BK + 2 * pq;           // padded K index of the pair
      if (kcol < a.Kpad) {
        int p = 0;
        while (p + 1 < a.n_pieces && kcol >= a.p[p + 1].k8) ++p;
        const float* pptr = a.p[p].ptr;
        const int pld = a.p[p].ld;
        const int kk = kcol - a.p[p].k8;              // column inside the piece
        const int nv = a.p[p].width - kk;             // valid floats from kk on (<= 0: padding)
        const bool al8 = a.p[p].al8 != 0;
        const int row0 = (tile0 + tq) * BM + ty;
        if (nv <= 0) {
#pragma unroll
          for (int i = 0; i < TM; ++i) cp_async8(dst + i * 16 * GR_AS, pptr, 0);
        } else if (al8 && arl == nullptr && (tile0 + tq + 1) * BM <= n) {
          const float* src = pptr + (size_t)row0 * pld + kk;
          const size_t step = (size_t)16 * pld;
          const int bytes = nv > 1 ? 8 : 4;
#pragma unroll
          for (int i = 0; i < TM; ++i) { cp_async8(dst + i * 16 * GR_AS, src, bytes); src += step; }
        } else {
#pragma unroll 2
          for (int i = 0; i < TM; ++i) {
            const int grow = row0 + 16 * i;
            const bool valid = grow < n;
            const float* src = pptr;
            if (valid) {
              const int gr = arl ? arl[grow] : grow;
              src = pptr + (size_t)gr * pld + kk;
            }
            if (al8) {
              cp_async8(dst + i * 16 * GR_AS, src, valid ? (nv > 1 ? 8 : 4) : 0);
            } else {
              cp_async4(dst + i * 16 * GR_AS, src, valid ? 4 : 0);
              cp_async4(dst + i * 16 * GR_AS + 1, (valid && nv > 1) ? src + 1 : pptr, (valid && nv > 1) ? 4 : 0);
            }
          }
        }
      }
      if (++i_sg == NS) { i_sg = 0; ++i_tq; }
      if (++i_st == GR_STAGES) i_st = 0;
    }
    cp_async_commit();
  };

  for (int s = 0; s < GR_STAGES - 1; ++s) issue();
  {  // resident weights + bias (plain loads; overlapped with the first stages in flight)
    const float4* w4 = reinterpret_cast<const float4*>(a.Wp);
    float4* s4 = reinterpret_cast<float4*>(Ws);
    for (int e = tid; e < a.Kpad * (BN / 4); e += 256) s4[e] = w4[e];
    for (int j = tid; j < BN; j += 256) sbias[j] = (a.bias && j < a.N) ? a.bias[j] : 0.f;
    for (int j = tid; j < 2 * BN; j += 256) (&colacc[0][0])[j] = 0.0;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int m = 0; m < TN; ++m) acc[i][m] = 0.f;
  int notconv = 0;

  int c_st = 0;
  for (int tq = 0; tq < my_tiles; ++tq)
  for (int sg = 0; sg < NS; ++sg) {
    cp_async_wait<GR_STAGES - 2>();
    __syncthreads();
    issue();
    const float* Ad = As + c_st * BM * GR_AS + ty * GR_AS;
    if (++c_st == GR_STAGES) c_st = 0;
    const float* Wd = Ws + sg * GR_BK * BN;
    const int nkq = min(GR_BK, a.Kpad - sg * GR_BK) >> 2;      // even (Kpad is a multiple of 8)
#pragma unroll 2
    for (int kq = 0; kq < nkq; ++kq) {
      float av[TM][4];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(Ad + i * 16 * GR_AS + 4 * kq);
        av[i][0] = t.x; av[i][1] = t.y; av[i][2] = t.z; av[i][3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* wr = Wd + (4 * kq + k) * BN;
        float bv[TN];
        if (TN >= 4) {
          const float4 t = *reinterpret_cast<const float4*>(wr + 4 * tx);
          bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
#pragma unroll
          for (int m = 4; m < TN; ++m) bv[m] = wr[64 + 16 * (m - 4) + tx];
        } else {
#pragma unroll
          for (int m = 0; m < TN; ++m) bv[m] = wr[tx + 16 * m];
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int m = 0; m < TN; ++m) acc[i][m] = fmaf(av[i][k], bv[m], acc[i][m]);
      }
    }
    if (sg == NS - 1) {
      // ---- epilogue of this tile -------------------------------------------------------------------------------
      const int row0 = (tile0 + tq) * BM + ty;
      int col[TN];
      bool cval[TN];
      float cadd[TN];                                  // FWD: folded bias; else: column scale
#pragma unroll
      for (int m = 0; m < TN; ++m) {
        col[m] = gr_col<TN>(m, tx);
        cval[m] = col[m] < a.N;
        if (FWD) cadd[m] = sbias[col[m]];
        else cadd[m] = (a.colscale && cval[m]) ? a.colscale[col[m]] : 1.0f;
      }
      const bool quad = TN >= 4 && a.vec2 && 4 * tx + 3 < a.N;      // this thread's column quad as two 8-byte accesses
      const bool selu = a.act == GNNFP_ACT_SELU;
      float cs[TN], cq[TN];
#pragma unroll
      for (int m = 0; m < TN; ++m) { cs[m] = 0.f; cq[m] = 0.f; }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int grow = row0 + 16 * i;
        const bool valid = grow < n;
        float sd = 0.f, sp = 0.f;
        float v[TN];
        if (FWD) {
#pragma unroll
          for (int m = 0; m < TN; ++m) {
            const float z = acc[i][m] + cadd[m];
            v[m] = selu ? selu_fwd(z) : act_fwd(a.act, z);
          }
          if (a.act == GNNFP_ACT_SOFTMAX) {   // Keras softmax over the row: its columns live in the 16 lanes sharing ty
            float mx = -INFINITY;
#pragma unroll
            for (int m = 0; m < TN; ++m) if (cval[m]) mx = fmaxf(mx, v[m]);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float se = 0.f;
#pragma unroll
            for (int m = 0; m < TN; ++m) { v[m] = cval[m] ? expf(v[m] - mx) : 0.f; se += v[m]; }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
#pragma unroll
            for (int m = 0; m < TN; ++m) v[m] = v[m] / se;
          }
#pragma unroll
          for (int m = 0; m < TN; ++m) if (!cval[m]) v[m] = 0.f;
        }
        if (valid) {
          const int gr = a.rowlist ? a.rowlist[grow] : grow;
          float* orow = a.out + (size_t)(a.out_compact ? grow : gr) * a.ld_out;
          float pv[TN];
          if (FWD) {
            if (a.prev) {
              const float* prow = a.prev + (size_t)gr * a.ld_prev;
              if (quad) {
                const float2 p0 = *reinterpret_cast<const float2*>(prow + 4 * tx);
                const float2 p1 = *reinterpret_cast<const float2*>(prow + 4 * tx + 2);
                pv[0] = p0.x; pv[1] = p0.y; pv[2] = p1.x; pv[3] = p1.y;
#pragma unroll
                for (int m = 4; m < TN; ++m) pv[m] = cval[m] ? prow[col[m]] : 0.f;
              } else {
#pragma unroll
                for (int m = 0; m < TN; ++m) pv[m] = cval[m] ? prow[col[m]] : 0.f;
              }
            }
            if (a.prev) {
#pragma unroll
              for (int m = 0; m < TN; ++m) {
                const float dd = v[m] - pv[m];
                sd = fmaf(dd, dd, sd);
                sp = fmaf(pv[m], pv[m], sp);
              }
            }
#pragma unroll
            for (int m = 0; m < TN; ++m) { cs[m] += v[m]; cq[m] = fmaf(v[m], v[m], cq[m]); }
          } else {
#pragma unroll
            for (int m = 0; m < TN; ++m) v[m] = acc[i][m] * cadd[m];
            if (a.corr) {
              const float* xr = a.corr_x + (size_t)gr * a.corr_ld;
              const float* k = a.corr + a.corr_col0;
#pragma unroll
              for (int m = 0; m < TN; ++m)
                if (cval[m]) {
                  const int j = col[m];
                  v[m] -= k[j] + fmaf(xr[j], k[2 * a.corr_in + j], k[3 * a.corr_in + j]) * k[a.corr_in + j];
                }
            }
            if (a.out_add) {
              if (quad) {
                const float2 p0 = *reinterpret_cast<const float2*>(orow + 4 * tx);
                const float2 p1 = *reinterpret_cast<const float2*>(orow + 4 * tx + 2);
                v[0] += p0.x; v[1] += p0.y; v[2] += p1.x; v[3] += p1.y;
#pragma unroll
                for (int m = 4; m < TN; ++m) if (cval[m]) v[m] += orow[col[m]];
              } else {
#pragma unroll
                for (int m = 0; m < TN; ++m) if (cval[m]) v[m] += orow[col[m]];
              }
            }
          }
          if (quad) {
            *reinterpret_cast<float2*>(orow + 4 * tx) = make_float2(v[0], v[1]);
            *reinterpret_cast<float2*>(orow + 4 * tx + 2) = make_float2(v[2], v[3]);
#pragma unroll
            for (int m = 4; m < TN; ++m) if (cval[m]) orow[col[m]] = v[m];
          } else {
#pragma unroll
            for (int m = 0; m < TN; ++m) if (cval[m]) orow[col[m]] = v[m];
          }
        }
#pragma unroll
        for (int m = 0; m < TN; ++m) acc[i][m] = 0.f;
        if (FWD && a.prev) {   // row sums live in the 16 lanes that share ty: reduce over tx
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            sd += __shfl_xor_sync(0xffffffffu, sd, o);
            sp += __shfl_xor_sync(0xffffffffu, sp, o);
          }
          if (valid && sqrtf(sd) > a.thr * sqrtf(sp)) notconv = 1;
        }
      }
      if (FWD && a.ost_sum) {
#pragma unroll
        for (int m = 0; m < TN; ++m) {   // fold the warp's two row groups, then one shared atomic per (warp, column)
          const float s1 = cs[m] + __shfl_xor_sync(0xffffffffu, cs[m], 16);
          const float q1 = cq[m] + __shfl_xor_sync(0xffffffffu, cq[m], 16);
          if ((tid & 16) == 0 && cval[m]) {
            atomicAdd(&colacc[0][col[m]], (double)s1);
            atomicAdd(&colacc[1][col[m]], (double)q1);
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  if (FWD && a.flag_next) {
    const int any = __syncthreads_or(notconv);
    if (tid == 0 && any) atomicOr(a.flag_next, 1);
  }
  if (FWD && a.ost_sum) {
    __syncthreads();
    for (int j = tid; j < a.N; j += 256) {
      atomicAdd(a.ost_sum + j, colacc[0][j]);
      atomicAdd(a.ost_sq + j, colacc[1][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// dW = X^T dz (+ db) over this CTA's rows.  32-row stages of X (all pieces side by side at their even k8 offsets) and dz
// stream through the cp.async ring; every thread reads the SAME row of the stage at a time (pure broadcasts).  Thread
// (ty, tx) owns the K quads 64*q + 4*ty .. +3 (q < NQ) and the dz columns gr_col<TJ>(m, tx): 4*NQ x TJ accumulators.
#define DW_ROWS 32
template <int NQ, int TJ>
__global__ void __launch_bounds__(256, 2) gemm_dw_kernel(const __grid_constant__ GemmDwArgs a) {
  if (a.gate && *a.gate == 0) return;
  constexpr int BK = 64 * NQ, BJ = 16 * TJ, TC = 4 * NQ;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                                   // [STAGES][DW_ROWS][BK]
  float* Zs = smem + GEMM_STAGES * DW_ROWS * BK;      // [STAGES][DW_ROWS][BJ]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_rows;
  const int n_chunks_all = (n + DW_ROWS - 1) / DW_ROWS;
  const int per_cta = (n_chunks_all + gridDim.x - 1) / gridDim.x;
  const int c0 = blockIdx.x * per_cta;
  const int c1 = min(n_chunks_all, c0 + per_cta);
  const int total = max(0, c1 - c0);
  const int xpairs = a.Kp >> 1, zpairs = (a.H + 1) >> 1;
  const bool zal8 = ((reinterpret_cast<uintptr_t>(a.dz) & 7) == 0) && (a.ld_dz % 2 == 0);

  // warp w copies rows w, w+8, w+16, w+24 of the stage; its lanes cover the column pairs
  auto issue = [&](int it) {
    if (it < total) {
      const int st = it % GEMM_STAGES;
      float* Xd = Xs + st * DW_ROWS * BK;
      float* Zd = Zs + st * DW_ROWS * BJ;
      const int row0 = (c0 + it) * DW_ROWS + warp;
      const bool fullrows = a.rowlist == nullptr && (c0 + it + 1) * DW_ROWS <= n;
      for (int pp = lane; pp < xpairs; pp += 32) {
        const int kcol = 2 * pp;
        int p = 0;
        while (p + 1 < a.n_pieces && kcol >= a.p[p + 1].k8) ++p;
        const float* pptr = a.p[p].ptr;
        const int pld = a.p[p].ld;
        const int kk = kcol - a.p[p].k8;
        const int nv = a.p[p].width - kk;
        const bool al8 = a.p[p].al8 != 0;
        float* dst = Xd + warp * BK + kcol;
        if (nv <= 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) cp_async8(dst + i * 8 * BK, pptr, 0);
        } else if (al8 && fullrows) {
          const float* src = pptr + (size_t)row0 * pld + kk;
          const int bytes = nv > 1 ? 8 : 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) { cp_async8(dst + i * 8 * BK, src, bytes); src += (size_t)8 * pld; }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int grow = row0 + 8 * i;
            const bool valid = grow < n;
            const float* src = pptr;
            if (valid) src = pptr + (size_t)(a.rowlist ? a.rowlist[grow] : grow) * pld + kk;
            if (al8) cp_async8(dst + i * 8 * BK, src, valid ? (nv > 1 ? 8 : 4) : 0);
            else {
              cp_async4(dst + i * 8 * BK, src, valid ? 4 : 0);
              cp_async4(dst + i * 8 * BK + 1, (valid && nv > 1) ? src + 1 : pptr, (valid && nv > 1) ? 4 : 0);
            }
          }
        }
      }
      for (int pp = lane; pp < zpairs; pp += 32) {
        const int j0 = 2 * pp;
        const int nv = a.H - j0;
        float* dst = Zd + warp * BJ + j0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int grow = row0 + 8 * i;
          const bool valid = grow < n;
          const float* src = a.dz;
          if (valid) src = a.dz + (size_t)((a.rowlist && !a.dz_compact) ? a.rowlist[grow] : grow) * a.ld_dz + j0;
          if (zal8) cp_async8(dst + i * 8 * BJ, src, valid ? (nv > 1 ? 8 : 4) : 0);
          else {
            cp_async4(dst + i * 8 * BJ, src, valid ? 4 : 0);
            cp_async4(dst + i * 8 * BJ + 1, (valid && nv > 1) ? src + 1 : a.dz, (valid && nv > 1) ? 4 : 0);
          }
        }
      }
    }
    cp_async_commit();
  };

  // zero the shared ring once: columns outside the pieces (K padding, H padding) are never written by cp.async
  for (int e = tid; e < GEMM_STAGES * DW_ROWS * (BK + BJ); e += 256) smem[e] = 0.f;
  __syncthreads();

  float acc[TC][TJ];
#pragma unroll
  for (int i = 0; i < TC; ++i)
#pragma unroll
    for (int m = 0; m < TJ; ++m) acc[i][m] = 0.f;
  float dbacc[TJ];
#pragma unroll
  for (int m = 0; m < TJ; ++m) dbacc[m] = 0.f;

  for (int s = 0; s < GEMM_STAGES - 1; ++s) issue(s);
  for (int it = 0; it < total; ++it) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    issue(it + GEMM_STAGES - 1);
    const int st = it % GEMM_STAGES;
    const float* Xd = Xs + st * DW_ROWS * BK + 4 * ty;
    const float* Zd = Zs + st * DW_ROWS * BJ;
#pragma unroll 4
    for (int r = 0; r < DW_ROWS; ++r) {
      float xv[TC], zv[TJ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(Xd + r * BK + 64 * q);
        xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w;
      }
      if (TJ >= 4) {
        const float4 t = *reinterpret_cast<const float4*>(Zd + r * BJ + 4 * tx);
        zv[0] = t.x; zv[1] = t.y; zv[2] = t.z; zv[3] = t.w;
#pragma unroll
        for (int m = 4; m < TJ; ++m) zv[m] = Zd[r * BJ + 64 + 16 * (m - 4) + tx];
      } else {
#pragma unroll
        for (int m = 0; m < TJ; ++m) zv[m] = Zd[r * BJ + tx + 16 * m];
      }
#pragma unroll
      for (int i = 0; i < TC; ++i)
#pragma unroll
        for (int m = 0; m < TJ; ++m) acc[i][m] = fmaf(xv[i], zv[m], acc[i][m]);
      if (ty == 0) {
#pragma unroll
        for (int m = 0; m < TJ; ++m) dbacc[m] += zv[m];
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- flush: this CTA's partial slot (private, accumulated across launches with +=) -------------------------
  float* part = a.partial + (size_t)blockIdx.x * a.n_params;
  float* sdb = smem;                     // [BJ] db of this CTA (for the BN correction of every row of dW)
  float* sQ = smem + BJ;                 // [BK] sum_j W[c][j]*acc[c][j]
  for (int e = tid; e < BJ + BK; e += 256) smem[e] = 0.f;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int m = 0; m < TJ; ++m) sdb[gr_col<TJ>(m, tx)] = dbacc[m];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TC; ++i) {
    const int kp = 64 * (i >> 2) + 4 * ty + (i & 3);
    int creal = -1, coff = 0;            // padded K index -> real input column (pieces sit at their k8 offsets)
    for (int p = 0; p < a.n_pieces; ++p) {
      if (kp >= a.p[p].k8 && kp < a.p[p].k8 + a.p[p].width) creal = coff + (kp - a.p[p].k8);
      coff += a.p[p].width;
    }
    float q = 0.f;
#pragma unroll
    for (int m = 0; m < TJ; ++m) {
      const int j = gr_col<TJ>(m, tx);
      if (j < a.H && creal >= 0) {
        float v = acc[i][m];
        if (a.bn_partial) q = fmaf(a.W[(size_t)creal * a.H + j], v, q);
        if (a.bnA) v = a.gamma[creal] * fmaf(a.bnA[creal], v, a.bnB[creal] * sdb[j]) + a.beta[creal] * sdb[j];
        part[(size_t)creal * a.H + j] += v;
      }
    }
    if (a.bn_partial) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      if (tx == 0 && creal >= 0) sQ[kp] = q;
    }
  }
  if (ty == 0) {
#pragma unroll
    for (int m = 0; m < TJ; ++m) {
      const int j = gr_col<TJ>(m, tx);
      if (j < a.H) part[a.bias_off + j] += dbacc[m];
    }
  }
  if (a.bn_partial) {
    __syncthreads();
    int K = 0;
    for (int p = 0; p < a.n_pieces; ++p) K += a.p[p].width;
    float* bp = a.bn_partial + (size_t)blockIdx.x * 2 * K;
    for (int c = tid; c < K; c += 256) {
      int kp = 0, coff = 0;
      for (int p = 0; p < a.n_pieces; ++p) {
        if (c >= coff && c < coff + a.p[p].width) kp = a.p[p].k8 + (c - coff);
        coff += a.p[p].width;
      }
      float P = 0.f;
      for (int j = 0; j < a.H; ++j) P = fmaf(a.W[(size_t)c * a.H + j], sdb[j], P);
      bp[c] = P;
      bp[K + c] = fmaf(a.bnA[c], sQ[kp], a.bnB[c] * P);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host launchers
static int grid_for(int n_tiles, int per_sm) {
  const int cap = gnnfp_num_sms() * per_sm;
  int g = n_tiles < cap ? n_tiles : cap;
  return g < 1 ? 1 : g;
}

template <int TM, int TN, bool FWD>
static int launch_rows_t(const GemmRowsArgs& a, cudaStream_t s, int prof_cat) {
  constexpr int BM = 16 * TM, BN = 16 * TN;
  const size_t smem = ((size_t)a.Kpad * BN + (size_t)GR_STAGES * BM * GR_AS) * sizeof(float);
  if (a.ldw != BN || a.Kpad % 8 != 0 || smem > 226 * 1024)
    GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows: padded weights %d x %d (ldw %d) do not fit the resident-weight kernel", a.Kpad, BN, a.ldw);
  static size_t attr = 0;
  if (smem > attr) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(gemm_rows_kernel<TM, TN, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_tiles = (a.n_rows + BM - 1) / BM;
  const int per_sm = 2 * (smem + 2048) <= 227 * 1024 ? 2 : 1;
  ProfScope ps(prof_cat, s);
  gemm_rows_kernel<TM, TN, FWD><<<grid_for(n_tiles, per_sm), 256, smem, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// N output columns -> (TM, TN): wider outputs take fewer rows per thread to bound registers
int launch_gemm_rows(const GemmRowsArgs& a, cudaStream_t s, int prof_cat) {
  if (a.n_rows <= 0) return GNNFP_OK;
  // tensor-core (tcgen05, 3xTF32) version for every eligible shape; GNNFP_TC=0 selects the FP32-pipe kernels below,
  // GNNFP_TC=<n> restricts the tensor-core path to N >= n output columns
  static const int tc_min_n = [] { const char* e = getenv("GNNFP_TC"); return e ? atoi(e) : 1; }();
  if (tc_min_n > 0 && a.N >= tc_min_n && gemm_rows_tc_supported(a)) return launch_gemm_rows_tc(a, s, prof_cat);
  if (a.nblk == 2) {                                   // two output blocks that do not fit one tensor-core launch: one launch each
    GemmRowsArgs b0 = a, b1 = a;
    b0.nblk = 0;
    b1.nblk = 0;
    b1.Wp = a.Wp2; b1.colscale = a.colscale2; b1.corr_col0 = a.corr_col02; b1.corr_x = a.corr_x2; b1.corr_ld = a.corr_ld2;
    b1.out = a.out2; b1.ld_out = a.ld_out2; b1.out_add = a.out_add2;
    b1.vec2 = a.ld_out2 % 2 == 0 && ((uintptr_t)a.out2 & 7) == 0;
    const int rc = launch_gemm_rows(b0, s, prof_cat);
    return rc ? rc : launch_gemm_rows(b1, s, prof_cat);
  }
  const int tn = (a.N + 15) / 16;
  const bool fwd = a.fwd != 0;
  switch (tn) {
    case 1: return fwd ? launch_rows_t<8, 1, true>(a, s, prof_cat) : launch_rows_t<8, 1, false>(a, s, prof_cat);
    case 2: return fwd ? launch_rows_t<8, 2, true>(a, s, prof_cat) : launch_rows_t<8, 2, false>(a, s, prof_cat);
    case 3: return fwd ? launch_rows_t<8, 3, true>(a, s, prof_cat) : launch_rows_t<8, 3, false>(a, s, prof_cat);
    case 4: return fwd ? launch_rows_t<8, 4, true>(a, s, prof_cat) : launch_rows_t<8, 4, false>(a, s, prof_cat);
    case 5: return fwd ? launch_rows_t<8, 5, true>(a, s, prof_cat) : launch_rows_t<8, 5, false>(a, s, prof_cat);
    default: GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows: %d output columns (max 80 per launch)", a.N);
  }
}
int gemm_rows_ldw(int N) { return 16 * ((N + 15) / 16); }
int gemm_rows_kpad(int k2) { return (k2 + 7) / 8 * 8; }
int gemm_rows_supported(int k8, int N) {
  return N <= 80 && ((size_t)gemm_rows_kpad(k8) * gemm_rows_ldw(N) + (size_t)GR_STAGES * 128 * GR_AS) * sizeof(float) <= 226 * 1024;
}
void gemm_piece_set(GemmPiece& g, const float* ptr, int ld, int width, int k8) {
  g.ptr = ptr; g.ld = ld; g.width = width; g.k8 = k8;
  g.al8 = ((reinterpret_cast<uintptr_t>(ptr) & 7) == 0 && ld % 2 == 0) ? 1 : 0;
}

template <int NQ, int TJ>
static int launch_dw_t(const GemmDwArgs& a, cudaStream_t s, int prof_cat, int* grid_out) {
  constexpr int BK = 64 * NQ, BJ = 16 * TJ;
  const size_t smem = (size_t)GEMM_STAGES * DW_ROWS * (BK + BJ) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(gemm_dw_kernel<NQ, TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int grid = gemm_dw_grid(a.n_rows);
  if (grid_out) *grid_out = grid;
  ProfScope ps(prof_cat, s);
  gemm_dw_kernel<NQ, TJ><<<grid, 256, smem, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

int gemm_dw_supported(int Kp, int H) { return Kp <= 192 && H <= 80; }
int gemm_dw_grid(int n_rows) { return grid_for((((n_rows + DW_ROWS - 1) / DW_ROWS) + 7) / 8, 2); }   // >= 256 rows per CTA

int launch_gemm_dw(const GemmDwArgs& a, cudaStream_t s, int prof_cat, int* grid_out) {
  if (a.n_rows <= 0) { if (grid_out) *grid_out = 0; return GNNFP_OK; }
  // tensor-core (tcgen05, 3xTF32) version for every eligible shape; GNNFP_TC_DW=0 selects the FP32-pipe kernel below
  static const int tc_dw = [] { const char* e = getenv("GNNFP_TC_DW"); return e ? atoi(e) : 1; }();
  if (tc_dw > 0 && gemm_dw_tc_supported(a)) return launch_gemm_dw_tc(a, s, prof_cat, grid_out);
  if (a.Kp % 2 != 0 || !gemm_dw_supported(a.Kp, a.H)) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_dw: K=%d, H=%d outside the supported tile shapes", a.Kp, a.H);
  const int nq = (a.Kp + 63) / 64, tj = (a.H + 15) / 16;
#define DW_CASE(Q, J) if (nq == Q && tj == J) return launch_dw_t<Q, J>(a, s, prof_cat, grid_out)
  DW_CASE(1, 1); DW_CASE(1, 2); DW_CASE(1, 3); DW_CASE(1, 4); DW_CASE(1, 5);
  DW_CASE(2, 1); DW_CASE(2, 2); DW_CASE(2, 3); DW_CASE(2, 4); DW_CASE(2, 5);
  DW_CASE(3, 1); DW_CASE(3, 2); DW_CASE(3, 3); DW_CASE(3, 4); DW_CASE(3, 5);
#undef DW_CASE
  GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_dw: K=%d, H=%d outside the supported tile shapes", a.Kp, a.H);
}

// ------------------------------------------------------------------------------------------------------------
// small helper kernels around the GEMMs
__global__ void __launch_bounds__(256) fold_w_kernel(const __grid_constant__ FoldArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ float sm[];
  const NetDev& net = a.net;
  const int in = net.in_dim, H = net.widths[0];
  float* A = sm;
  float* B = sm + in;
  const int tid = threadIdx.x;
  if (net.bn_mode) {
    float* mean = (a.update_moving && net.bn_mode == 1 && blockIdx.x == 0) ? sm + 2 * in : nullptr;
    float* var = mean ? sm + 3 * in : nullptr;
    bn_coefficients(a.src, net, 1, A, B, mean, var);
    if (a.coef_out && blockIdx.x == 0 && net.bn_mode == 1) {   // what bn_coef_kernel would compute before the backward
      bn_coefficients(a.src, net, 0, a.coef_out, a.coef_out + in, nullptr, nullptr);
      for (int c = tid; c < in; c += blockDim.x) a.coef_out[2 * in + c] = net.gamma[c] * a.coef_out[c];   // own element
    }
    __syncthreads();
    if (mean) {   // Keras BatchNormalization._assign_moving_average: var -= (var - value) * (1 - momentum)
      const float decay = (float)(1.0 - (double)net.bn_momentum);
      for (int c = tid; c < in; c += blockDim.x) {
        net.mmean[c] -= (net.mmean[c] - mean[c]) * decay;
        net.mvar[c] -= (net.mvar[c] - var[c]) * decay;
      }
    }
  }
  const int total = a.Kpad * a.ldw;
  for (int e = blockIdx.x * blockDim.x + tid; e < total; e += gridDim.x * blockDim.x) {
    const int kp = e / a.ldw, j = e - kp * a.ldw;
    int creal = -1;
    for (int p = 0; p < a.src.n_pieces; ++p)
      if (kp >= a.k8[p] && kp < a.k8[p] + a.src.p[p].width) creal = a.src.p[p].col0 + (kp - a.k8[p]);
    float w = 0.f;
    if (creal >= 0 && j < H) {
      w = net.W[0][(size_t)creal * H + j];
      if (net.bn_mode) w *= A[creal];
    }
    a.Wp[e] = w;
  }
  // folded bias b_j + sum_c B_c W[c][j]: one warp per output column (lanes split the input columns, fixed-order
  // shuffle reduction), spread over all blocks - a serial loop over the input columns per thread cost ~8 us per launch
  const int lane = tid & 31, wpb = blockDim.x >> 5;
  for (int j = blockIdx.x * wpb + (tid >> 5); j < a.ldw; j += gridDim.x * wpb) {
    float part = 0.f;
    if (j < H && net.bn_mode)
      for (int c = lane; c < in; c += 32) part = fmaf(B[c], net.W[0][(size_t)c * H + j], part);
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) part += __shfl_xor_sync(0xffffffffu, part, of);
    if (lane == 0) a.biasp[j] = j < H ? net.b[0][j] + part : 0.f;
  }
}

int launch_fold_w(const FoldArgs& a, cudaStream_t s) {
  const int total = a.Kpad * a.ldw;
  int blocks = (total + 255) / 256;
  if (blocks > 32) blocks = 32;
  fold_w_kernel<<<blocks, 256, 4 * a.net.in_dim * sizeof(float), s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// backward: coef = [rstd | -mean*rstd | gamma*rstd] of this iteration's BN (per input column)
__global__ void __launch_bounds__(256) bn_coef_kernel(const __grid_constant__ BnCoefArgs a) {
  if (a.gate && *a.gate == 0) return;
  const int in = a.net.in_dim;
  bn_coefficients(a.src, a.net, 0, a.coef, a.coef + in, nullptr, nullptr);
  __syncthreads();
  for (int c = threadIdx.x; c < in; c += blockDim.x) a.coef[2 * in + c] = a.net.gamma[c] * a.coef[c];
}
int launch_bn_coef(const BnCoefArgs& a, cudaStream_t s) {
  bn_coef_kernel<<<1, 256, 0, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// backward: WT block of one destination piece: out[kp][n] = W[col0 + n][kp]  (kp < H, n < width), zero padded
__global__ void transpose_block_kernel(const float* W, int H, int col0, int width, int Kpad, int ldw, float* out) {
  const int total = Kpad * ldw;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kp = e / ldw, nn = e - kp * ldw;
    out[e] = (kp < H && nn < width) ? W[(size_t)(col0 + nn) * H + kp] : 0.f;
  }
}
int launch_transpose_block(const float* W, int H, int col0, int width, int Kpad, int ldw, float* out, cudaStream_t s) {
  const int total = Kpad * ldw;
  int blocks = (total + 255) / 256;
  if (blocks > 64) blocks = 64;
  transpose_block_kernel<<<blocks, 256, 0, s>>>(W, H, col0, width, Kpad, ldw, out);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
