// kernels_fwd.cu - forward tile kernels of libgnnfp (sm_100a, FP32 CUDA cores).
//
//   tile_fwd_kernel  : one application of a net (Keras Sequential [BN] + Dense*) to a row set whose
//                      input is a concatenation of "pieces" gathered straight from HBM into a
//                      shared-memory tile.  For the state net this is one whole fixed-point iteration
//                      (reference GNN.py:217-236 convergence() + GNN.py:196-214 condition()):
//                      sparse aggregation Adj^T.state over the device CSR, concat, BN, Dense(s),
//                      activation, new state, per-row convergence test -> device flag, and the
//                      column statistics the next iteration's BN needs.
//   tile_pass_kernel : materialise pieces and/or their column statistics (loop-invariant
//                      aggregates of GNN.py:254-258, BN batch statistics).
//
// Persistent grids (a multiple of the SM count), weights staged once per CTA into shared memory,
// activations kept row-major with an odd stride so that lane==row accesses are conflict-free and
// weight reads are 128-bit broadcasts.
#include <stdlib.h>
#include "tile.cuh"

// shared-memory layout as integer offsets (floats from the start of dynamic shared memory).  Kept in
// shared memory itself so that every access stays in the shared address space (LDS/STS, 32-bit addressing).
struct FwdLayout {
  int oW[GNNFP_MAX_LAYERS], ob[GNNFP_MAX_LAYERS];
  int obnA, obnB, oScr, obuf0, obuf1, oOst;
  int total;
};
__host__ __device__ inline void fwd_layout(const NetDev& net, int R, int XS0, int XS1, int cap, FwdLayout& y) {
  int o = 0;
  y.oOst = o; o += 4 * ceil_to(net.widths[net.n_layers - 1], 4);   // 2*H doubles
  int in_l = net.in_dim;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Hpad = ceil_to(net.widths[l], GNNFP_JC);
    y.oW[l] = o; o += ceil_to(in_l, 4) * Hpad;
    y.ob[l] = o; o += Hpad;
    in_l = net.widths[l];
  }
  y.obnA = o; o += ceil_to(net.in_dim, 4);
  y.obnB = o; o += ceil_to(net.in_dim, 4);
  y.oScr = o; o += (int)scratch_floats(R, cap);
  y.obuf0 = o; o += R * XS0;
  y.obuf1 = o; o += R * XS1;
  y.total = o;
}

__global__ void __launch_bounds__(512) tile_fwd_kernel(const __grid_constant__ FwdArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ __align__(16) float smem[];
  const NetDev& net = a.net;
  const TileCfg& tc = a.tc;
  __shared__ FwdLayout y;
  if (threadIdx.x == 0) fwd_layout(net, tc.R, tc.XS0, tc.XS1, tc.cap, y);
  __syncthreads();
  const int tid = threadIdx.x, T = blockDim.x;
  float* const bnA_ = smem + y.obnA;
  float* const bnB_ = smem + y.obnB;
  double* const ost = reinterpret_cast<double*>(smem + y.oOst);
  float* const buf0 = smem + y.obuf0;
  float* const buf1 = smem + y.obuf1;
  const StageScratch sc = carve_scratch(smem + y.oScr, tc.R, tc.cap);
  const int lane = tid & 31, warp = tid >> 5;
  const int rg = warp % tc.RG, cg = warp / tc.RG;
  const int L = net.n_layers;
  const int H = net.widths[L - 1];

  // ---- BN coefficients (x_hat = x*a + b), moving-average update -------------------------------------
  if (net.bn_mode) {
    bn_coefficients(a.src, net, 1, bnA_, bnB_, nullptr, nullptr);
    if (a.update_moving && net.bn_mode == 1 && blockIdx.x == 0) {
      // Keras BatchNormalization._assign_moving_average: var -= (var - value) * (1 - momentum)
      const float decay = (float)(1.0 - (double)net.bn_momentum);
      for (int cc = tid; cc < net.in_dim; cc += T) {
        float mean = 0.f, var = 0.f;
        for (int p = 0; p < a.src.n_pieces; ++p) {
          const Piece& pc = a.src.p[p];
          if (cc >= pc.col0 && cc < pc.col0 + pc.width && pc.st_sum && !pc.accumulate) {
            const double m = pc.st_sum[cc - pc.col0] * net.inv_n;
            double v = pc.st_sq[cc - pc.col0] * net.inv_n - m * m;
            if (v < 0.0) v = 0.0;
            mean = (float)m;
            var = (float)v;
          }
        }
        net.mmean[cc] -= (net.mmean[cc] - mean) * decay;
        net.mvar[cc] -= (net.mvar[cc] - var) * decay;
      }
    }
    __syncthreads();
  }
  // ---- weights: zero padded to [ceil4(in)][ceil16(H)]; BN affine folded into layer 0 -------------------
  //      x_hat . W + b = x . (a (.) W) + (b + bnB . W)
  {
    int in_l = net.in_dim;
    for (int l = 0; l < L; ++l) {
      const int Hl = net.widths[l], Hpad = ceil_to(Hl, GNNFP_JC), inp = ceil_to(in_l, 4);
      const bool fold = (l == 0 && net.bn_mode != 0);
      for (int e = tid; e < inp * Hpad; e += T) {
        const int c = e / Hpad, j = e - c * Hpad;
        float w = (j < Hl && c < in_l) ? net.W[l][(size_t)c * Hl + j] : 0.0f;
        if (fold && c < in_l) w *= bnA_[c];
        smem[y.oW[l] + e] = w;
      }
      for (int j = tid; j < Hpad; j += T) {
        float bj = j < Hl ? net.b[l][j] : 0.0f;
        if (fold && j < Hl)
          for (int c = 0; c < in_l; ++c) bj = fmaf(bnB_[c], net.W[l][(size_t)c * Hl + j], bj);
        smem[y.ob[l] + j] = bj;
      }
      in_l = Hl;
    }
  }
  for (int j = tid; j < 2 * H; j += T) ost[j] = 0.0;
  for (int e = tid; e < tc.R * (tc.XS0 + tc.XS1); e += T) buf0[e] = 0.0f;   // buf0 and buf1 are adjacent
  __syncthreads();

  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  int notconv = 0;
  const unsigned magicH = (unsigned)((0x100000000ull + (unsigned)H - 1) / (unsigned)H);

  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_end = min(n_tiles, ((int)blockIdx.x + 1) * tiles_per_cta);
  for (int tile = blockIdx.x * tiles_per_cta; tile < tile_end; ++tile) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    stage_tile(a.src, row0, nr, tc.R, buf0, tc.XS0, sc);
    __syncthreads();
    if (a.agg_out) {
      const unsigned magicA = (unsigned)((0x100000000ull + (unsigned)a.agg_w - 1) / (unsigned)a.agg_w);
      for (int e = tid; e < nr * a.agg_w; e += T) {
        const int r = div_magic((unsigned)e, magicA);
        const int j = e - r * a.agg_w;
        const int orow = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
        a.agg_out[(size_t)orow * a.agg_w + j] = buf0[r * tc.XS0 + a.agg_col0 + j];
      }
    }
    // ---- the MLP --------------------------------------------------------------------------
    float* cur = buf0;
    float* nxt = buf1;
    int XSc = tc.XS0, XSn = tc.XS1;
    int in_l = net.in_dim;
    for (int l = 0; l < L; ++l) {
      const int Hl = net.widths[l], Hpad = ceil_to(Hl, GNNFP_JC);
      if (cg < Hpad / GNNFP_JC)
        dense_tile(cur, XSc, nxt, XSn, smem + y.oW[l], smem + y.ob[l], (in_l + 3) / 4, Hpad, net.acts[l], rg, cg, tc.CG, lane);
      __syncthreads();
      if (net.acts[l] == GNNFP_ACT_SOFTMAX) {
        softmax_rows(nxt, XSn, Hl, tc.R);
        __syncthreads();
      }
      float* t = cur; cur = nxt; nxt = t;
      const int ti = XSc; XSc = XSn; XSn = ti;
      in_l = Hl;
    }
    // ---- epilogue: `cur` holds the output tile [R][XSc], first H columns ------------------------
#pragma unroll 4
    for (int e = tid; e < nr * H; e += T) {
      const int r = div_magic((unsigned)e, magicH);
      const int j = e - r * H;
      const int orow = a.out_compact ? (row0 + r) : (a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r);
      a.out[(size_t)orow * a.ld_out + j] = cur[r * XSc + j];
    }
    if (a.prev) {   // GNN.py:200-209: sqrt(sum (s-s_old)^2) > thr * sqrt(sum s_old^2), strict
      const bool from_tile = (L == 1 && a.prev_col0 >= 0);   // the raw previous state still sits in buf0
      for (int r = tid; r < nr; r += T) {
        float sd = 0.f, sp = 0.f;
        if (from_tile) {
          const float* pv = buf0 + r * tc.XS0 + a.prev_col0;
          for (int j = 0; j < H; ++j) {
            const float p = pv[j];
            const float d = cur[r * XSc + j] - p;
            sd = fmaf(d, d, sd);
            sp = fmaf(p, p, sp);
          }
        } else {
          const int gr = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
          const float* pv = a.prev + (size_t)gr * a.ld_prev;
          for (int j = 0; j < H; ++j) {
            const float p = pv[j];
            const float d = cur[r * XSc + j] - p;
            sd = fmaf(d, d, sd);
            sp = fmaf(p, p, sp);
          }
        }
        if (sqrtf(sd) > a.thr * sqrtf(sp)) notconv = 1;
      }
    }
    if (a.ost_sum) tile_col_stats(cur, XSc, H, nr, ost);
    __syncthreads();
  }
  if (a.flag_next) {
    const int any = __syncthreads_or(notconv);
    if (tid == 0 && any) atomicOr(a.flag_next, 1);
  }
  if (a.ost_sum) {
    __syncthreads();
    for (int j = tid; j < H; j += T) {
      atomicAdd(a.ost_sum + j, ost[j]);
      atomicAdd(a.ost_sq + j, ost[H + j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_pass_kernel(const __grid_constant__ PassArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ __align__(16) float smem[];
  const TileCfg& tc = a.tc;
  const int W = a.src.in_dim;
  double* acc = reinterpret_cast<double*>(smem);          // [2*W]
  float* p = smem + 4 * ceil_to(W, 4);
  StageScratch sc = carve_scratch(p, tc.R, tc.cap);
  float* X = p + scratch_floats(tc.R, tc.cap);
  const int tid = threadIdx.x, T = blockDim.x;
  for (int j = tid; j < 2 * W; j += T) acc[j] = 0.0;
  __syncthreads();
  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  const unsigned magicW = (unsigned)((0x100000000ull + (unsigned)W - 1) / (unsigned)W);
  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_end = min(n_tiles, ((int)blockIdx.x + 1) * tiles_per_cta);
  for (int tile = blockIdx.x * tiles_per_cta; tile < tile_end; ++tile) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    stage_tile(a.src, row0, nr, tc.R, X, tc.XS0, sc);
    __syncthreads();
    if (a.out) {
      for (int e = tid; e < nr * W; e += T) {
        const int r = div_magic((unsigned)e, magicW);
        const int j = e - r * W;
        const int orow = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
        a.out[(size_t)orow * a.ld_out + j] = X[r * tc.XS0 + j];
      }
    }
    if (a.st_sum) tile_col_stats(X, tc.XS0, W, nr, acc);
    __syncthreads();
  }
  if (a.st_sum) {
    for (int j = tid; j < W; j += T) {
      atomicAdd(a.st_sum + j, acc[j]);
      atomicAdd(a.st_sq + j, acc[W + j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Adj^T . S with thousands of independent threads: thread (qx, ry) owns the VEC-wide column chunk qx of
// rows ry, ry + rows_per_block * gridDim, ...; per-thread fp64 column partials, one block reduction, one
// fp64 atomicAdd per column and block.  Arc order inside a row is preserved (sequential fmaf).
// BT threads per block: the statistics end in one fp64 atomicAdd per column and BLOCK on the same 2 D addresses, and
// same-address atomics serialise in L2 (measured ~8 us per launch with 4 x 148 blocks), so the narrow instantiations
// run ONE 1024-thread block per SM.
template <int VEC, bool DIRECT, bool WGT, int BT>
__global__ void __launch_bounds__(BT) agg_stats_kernel(const __grid_constant__ AggArgs a, int QX) {
  if (a.gate && *a.gate == 0) return;
  __shared__ double red[BT * 2];
  // pad4 (VEC == 4, D % 4 == 2): rows are walked as ceil(D / 4) float4 slots; the last slot reads two foreign (finite)
  // columns, which are neither stored nor counted - half the threads / index arithmetic of the float2 walk
  const bool pad = VEC == 4 && a.pad4 != 0;
  const int nq = pad ? (a.D + 3) / 4 : a.D / VEC;
  const int qx = threadIdx.x % QX, ry = threadIdx.x / QX;
  const int rpb = blockDim.x / QX;
  double su[VEC], sq[VEC];
  float ts[VEC], tq[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { su[v] = 0.0; sq[v] = 0.0; ts[v] = 0.f; tq[v] = 0.f; }
  const bool want_stats = a.st_sum != nullptr;
  if (qx < nq && ry < rpb) {
    // NR rows per trip: their rowptr -> idx -> state-row load chains overlap (the kernel is latency bound otherwise);
    // arc order inside a row is kept (sequential fmaf), absent arcs are predicated off.  DIRECT (no CSR: the row's
    // only entry is S[row], weight 1) is the column-statistics / copy pass of a plain matrix.
    constexpr int NR = DIRECT ? 8 : 4;
    const int n_rows = a.n_rows_dev ? *a.n_rows_dev : a.n_rows;
    const int* __restrict__ rowlist = a.rowlist;
    const float* __restrict__ Sq = a.S + qx * VEC;
    const size_t ld = (size_t)a.ld;
    float* __restrict__ outq = a.out ? a.out + qx * VEC : nullptr;
    const size_t ldo = (size_t)(a.ld_out ? a.ld_out : a.D);
    const int stride = gridDim.x * rpb;
    for (int r = blockIdx.x * rpb + ry; r < n_rows; r += NR * stride) {
      int gr[NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int rj = r + j * stride;
        gr[j] = rj < n_rows ? (rowlist ? rowlist[rj] : rj) : -1;
      }
      float acc[NR][VEC];
      if (DIRECT) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          if (gr[j] >= 0) load_vec<VEC>(Sq + (size_t)gr[j] * ld, acc[j]);
          else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[j][v] = 0.f;
          }
        }
      } else {
        const int* __restrict__ rowptr = a.rowptr;
        const int* __restrict__ idx = a.idx;
        const float* __restrict__ wgt = a.wgt;
        int b[NR], n[NR];
        int nmax = 0;
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          b[j] = gr[j] >= 0 ? rowptr[gr[j]] : 0;
          n[j] = gr[j] >= 0 ? rowptr[gr[j] + 1] - b[j] : 0;
          nmax = n[j] > nmax ? n[j] : nmax;
        }
#pragma unroll
        for (int j = 0; j < NR; ++j)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[j][v] = 0.f;
        for (int q = 0; q < nmax; q += 2) {
          float tv[NR][2][VEC], w[NR][2];
          int ix[NR][2];
#pragma unroll
          for (int j = 0; j < NR; ++j)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const bool ok = q + u < n[j];
              ix[j][u] = ok ? idx[b[j] + q + u] : -1;
              if (WGT) w[j][u] = ok ? wgt[b[j] + q + u] : 0.0f;
              else w[j][u] = 1.0f;
            }
#pragma unroll
          for (int j = 0; j < NR; ++j)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (ix[j][u] >= 0) load_vec<VEC>(Sq + (size_t)ix[j][u] * ld, tv[j][u]);
              else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) tv[j][u][v] = 0.f;
              }
            }
#pragma unroll
          for (int j = 0; j < NR; ++j)
#pragma unroll
            for (int u = 0; u < 2; ++u)
              if (ix[j][u] >= 0) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[j][v] = fmaf(w[j][u], tv[j][u][v], acc[j][v]);
              }
        }
      }
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        if (gr[j] < 0) continue;
        if (outq) {
          float* o = outq + (size_t)gr[j] * ldo;
          if (VEC == 4 && pad) {
            *reinterpret_cast<float2*>(o) = make_float2(acc[j][0], acc[j][1 % VEC]);
            if (qx * 4 + 2 < a.D) *reinterpret_cast<float2*>(o + 2) = make_float2(acc[j][2 % VEC], acc[j][3 % VEC]);
          } else if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[j][0], acc[j][1 % VEC], acc[j][2 % VEC], acc[j][3 % VEC]);
          else if (VEC == 2) *reinterpret_cast<float2*>(o) = make_float2(acc[j][0], acc[j][1 % VEC]);
          else o[0] = acc[j][0];
        }
        if (want_stats) {                              // uniform branch: no fp64 work at all without statistics
#pragma unroll
          for (int v = 0; v < VEC; ++v) { ts[v] += acc[j][v]; tq[v] = fmaf(acc[j][v], acc[j][v], tq[v]); }
        }
      }
      if (want_stats) {                                // one conversion per trip (NR rows): fp32 partials of <= 8 values
#pragma unroll
        for (int v = 0; v < VEC; ++v) { su[v] += (double)ts[v]; sq[v] += (double)tq[v]; ts[v] = 0.f; tq[v] = 0.f; }
      }
    }
  }
  if (a.st_sum) {
    for (int v = 0; v < VEC; ++v) {
      red[threadIdx.x] = su[v];
      red[BT + threadIdx.x] = sq[v];
      __syncthreads();
      if (ry == 0 && qx < nq) {
        double s1 = 0.0, s2 = 0.0;
        for (int y2 = 0; y2 < rpb; ++y2) { s1 += red[y2 * QX + qx]; s2 += red[BT + y2 * QX + qx]; }
        if (qx * VEC + v < a.D) {
          atomicAdd(a.st_sum + qx * VEC + v, s1);
          atomicAdd(a.st_sq + qx * VEC + v, s2);
        }
      }
      __syncthreads();
    }
  }
}

template <int VEC, bool DIRECT, bool WGT>
static int launch_agg_t(const AggArgs& a, int QX, cudaStream_t s) {
  constexpr int NR = DIRECT ? 8 : 4;
  constexpr int BT = VEC == 4 ? 256 : 1024;            // VEC == 4 needs more than 64 registers per thread
  const int rpb = BT / QX;
  long long blocks = ((long long)a.n_rows + NR * rpb - 1) / (NR * rpb);
  static int occ = 0;                                  // resident blocks per SM of this instantiation: one full wave
  if (!occ) {
    int o = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, agg_stats_kernel<VEC, DIRECT, WGT, BT>, BT, 0);
    occ = o > 0 ? o : 4;
  }
  static const int waves = getenv("GNNFP_AGG_WAVES") ? atoi(getenv("GNNFP_AGG_WAVES")) : 1;
  const long long cap = (long long)gnnfp_num_sms() * occ * (waves > 0 ? waves : 1);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  agg_stats_kernel<VEC, DIRECT, WGT, BT><<<(int)blocks, BT, 0, s>>>(a, QX);
  return 0;
}

template <int VEC>
static int launch_agg_v(const AggArgs& a, int QX, cudaStream_t s) {
  if (!a.rowptr) return launch_agg_t<VEC, true, false>(a, QX, s);
  if (a.wgt) return launch_agg_t<VEC, false, true>(a, QX, s);
  return launch_agg_t<VEC, false, false>(a, QX, s);
}

int launch_agg_stats(const AggArgs& a, cudaStream_t s, int prof_cat) {
  if (a.n_rows <= 0) return GNNFP_OK;
  auto al = [&](const void* p, int m) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & (m - 1)) == 0; };
  int vec = 1;
  const int ldo = a.ld_out ? a.ld_out : a.D;
  int pad4 = 0;
  if (a.D % 4 == 0 && a.ld % 4 == 0 && ldo % 4 == 0 && al(a.S, 16) && al(a.out, 16)) vec = 4;
  // (padded float4 walk: measured slower than the float2 walk - fewer loads in flight - so only on request)
  else if (getenv("GNNFP_AGG_PAD4") && a.D % 4 == 2 && a.ld % 4 == 0 && ldo % 2 == 0 && al(a.S, 16) && al(a.out, 8) && a.ld >= a.D + 2) { vec = 4; pad4 = 1; }
  else if (a.D % 2 == 0 && a.ld % 2 == 0 && ldo % 2 == 0 && al(a.S, 8) && al(a.out, 8)) vec = 2;
  const int nq = pad4 ? (a.D + 3) / 4 : a.D / vec;
  if (nq > 256) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "state width %d too large for the aggregation kernel", a.D);
  const int QX = nq;                                  // threads per row (rows may straddle warps)
  ProfScope ps(prof_cat ? prof_cat : PC_AGG, s);
  if (vec == 4) { AggArgs b = a; b.pad4 = pad4; launch_agg_v<4>(b, QX, s); }
  else if (vec == 2) launch_agg_v<2>(a, QX, s);
  else launch_agg_v<1>(a, QX, s);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// ------------------------------------------------------------------------------------------------
// host side: tile geometry + launch
// ------------------------------------------------------------------------------------------------
static size_t fwd_smem_bytes(const NetDev& net, int R, int XS0, int XS1, int cap) {
  FwdLayout y;
  fwd_layout(net, R, XS0, XS1, cap, y);
  return (size_t)y.total * 4 + sizeof(FwdLayout) + 64;
}

int tile_cfg_fwd(const NetDev& net, int n_rows, TileCfg* tc) {
  int w0 = net.in_dim, w1 = 1, hpmax = 0;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Hpad = ceil_to(net.widths[l], GNNFP_JC);
    hpmax = Hpad > hpmax ? Hpad : hpmax;
    if (l % 2 == 0) w1 = Hpad > w1 ? Hpad : w1; else w0 = Hpad > w0 ? Hpad : w0;
  }
  tc->XS0 = tile_stride(w0);
  tc->XS1 = tile_stride(w1);
  const int cap_per_row = tc->cap_per_row > 0 ? tc->cap_per_row : 4;
  int CG = hpmax / GNNFP_JC;
  if (CG > 8) CG = 8;
  const int nsm = gnnfp_num_sms();
  const size_t cap = 216 * 1024, sm_budget = 224 * 1024;
  // pick the row-group count that keeps the most warps resident per SM (ties: the smaller tile), then
  // shrink further while the grid would not cover the GPU twice
  int best = 0, best_warps = -1;
  for (int RG = 1; RG <= 4; RG *= 2) {
    if (32 * RG * CG > 512) break;
    const size_t b = fwd_smem_bytes(net, 64 * RG, tc->XS0, tc->XS1, 64 * RG * cap_per_row);
    if (b > cap) break;
    int ctas = (int)(sm_budget / (b + 1024));
    const int by_thr = 2048 / (32 * RG * CG);
    if (ctas > by_thr) ctas = by_thr;
    if (ctas > 16) ctas = 16;
    const int warps = ctas * RG * CG;
    if (warps > best_warps) { best_warps = warps; best = RG; }
  }
  if (best == 0)
    GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "net too large for the shared-memory tile kernel (%zu bytes needed)",
               fwd_smem_bytes(net, 64, tc->XS0, tc->XS1, 64 * cap_per_row));
  int RG = best;
  while (RG > 1 && (n_rows + 64 * RG - 1) / (64 * RG) < 2 * nsm) RG /= 2;
  tc->RG = RG;
  tc->CG = CG;
  tc->R = 64 * RG;
  tc->threads = 32 * RG * CG;
  tc->cap = tc->R * cap_per_row;
  tc->smem_bytes = fwd_smem_bytes(net, tc->R, tc->XS0, tc->XS1, tc->cap);
  int per_sm = (int)(sm_budget / (tc->smem_bytes + 1024));
  const int by_threads = 2048 / tc->threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  const int n_tiles = (n_rows + tc->R - 1) / tc->R;
  tc->grid = n_tiles < nsm * per_sm ? n_tiles : nsm * per_sm;
  if (tc->grid < 1) tc->grid = 1;
  return GNNFP_OK;
}

int tile_cfg_pass(int in_dim, int n_rows, TileCfg* tc) {
  tc->RG = 1;
  tc->CG = 1;
  tc->XS0 = tile_stride(in_dim);
  tc->XS1 = 4;
  tc->threads = 256;
  const int cap_per_row = tc->cap_per_row > 0 ? tc->cap_per_row : 4;
  int R = 128;
  auto bytes = [&](int r) { return ((size_t)r * tc->XS0 + 4 * (size_t)ceil_to(in_dim, 4) + scratch_floats(r, r * cap_per_row)) * 4; };
  while (R > 16 && bytes(R) > 48 * 1024) R /= 2;
  tc->R = R;
  tc->cap = R * cap_per_row;
  tc->smem_bytes = bytes(R);
  const int nsm = gnnfp_num_sms();
  int per_sm = (int)((200 * 1024) / (tc->smem_bytes + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int n_tiles = (n_rows + R - 1) / R;
  tc->grid = n_tiles < nsm * per_sm ? n_tiles : nsm * per_sm;
  if (tc->grid < 1) tc->grid = 1;
  return GNNFP_OK;
}

int launch_tile_fwd(const FwdArgs& a, cudaStream_t s) {
  if (a.src.n_rows <= 0) return GNNFP_OK;
  static size_t attr_set = 0;
  if (a.tc.smem_bytes > attr_set) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(tile_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(220 * 1024)));
    attr_set = 220 * 1024;
  }
  ProfScope ps(a.prof_cat ? a.prof_cat : PC_OTHER, s);
  tile_fwd_kernel<<<a.tc.grid, a.tc.threads, a.tc.smem_bytes, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

int launch_tile_pass(const PassArgs& a, cudaStream_t s) {
  if (a.src.n_rows <= 0) return GNNFP_OK;
  // a single plain piece (one matrix, or one CSR gather of it) needs no tile: the streaming gather kernel
  // copies / aggregates it and takes the column statistics at several TB/s
  if (a.src.n_pieces == 1) {
    const Piece& pc = a.src.p[0];
    if (pc.col0 == 0 && pc.width == a.src.in_dim && !pc.accumulate && !pc.compact && !pc.map && !pc.rowscale &&
        !pc.gate && (pc.kind == PK_DIRECT || pc.kind == PK_GATHER) && pc.width <= 256) {
      AggArgs aa;
      memset(&aa, 0, sizeof(aa));
      aa.n_rows = a.src.n_rows; aa.rowlist = a.src.rowlist; aa.D = pc.width;
      aa.S = pc.ptr; aa.ld = pc.ld;
      if (pc.kind == PK_GATHER) { aa.rowptr = pc.rowptr; aa.idx = pc.idx; aa.wgt = pc.wgt; }
      aa.out = a.out; aa.ld_out = a.ld_out;
      aa.st_sum = a.st_sum; aa.st_sq = a.st_sq;
      aa.gate = a.gate;
      return launch_agg_stats(aa, s, PC_PASS);
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(tile_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(100 * 1024)));
    attr_set = true;
  }
  ProfScope ps(PC_PASS, s);
  tile_pass_kernel<<<a.tc.grid, a.tc.threads, a.tc.smem_bytes, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
