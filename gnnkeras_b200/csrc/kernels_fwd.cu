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
#include "tile.cuh"

// ------------------------------------------------------------------------------------------------
// one Dense layer on the tile: each warp (rg, cg) owns 64 rows x 16-column chunks
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dense_layer_tile(const float* __restrict__ Ain, int XSin,
                                                 float* __restrict__ Aout, int XSout,
                                                 const float* __restrict__ Wl, const float* __restrict__ bl,
                                                 int in_l, int Hpad, int act, int rg, int cg, int CG, int lane) {
  const int nch = Hpad / GNNFP_JC;
  const float* x0p = Ain + (rg * 64 + lane) * XSin;
  const float* x1p = x0p + 32 * XSin;
  for (int ch = cg; ch < nch; ch += CG) {
    float acc0[GNNFP_JC], acc1[GNNFP_JC];
#pragma unroll
    for (int j = 0; j < GNNFP_JC; ++j) {
      const float bj = bl[ch * GNNFP_JC + j];
      acc0[j] = bj;
      acc1[j] = bj;
    }
    const float4* wp = reinterpret_cast<const float4*>(Wl + ch * GNNFP_JC);
    const int wstride = Hpad / 4;
#pragma unroll 2
    for (int c = 0; c < in_l; ++c) {
      const float x0 = x0p[c], x1 = x1p[c];
      const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
      wp += wstride;
      acc0[0] = fmaf(x0, w0.x, acc0[0]);   acc1[0] = fmaf(x1, w0.x, acc1[0]);
      acc0[1] = fmaf(x0, w0.y, acc0[1]);   acc1[1] = fmaf(x1, w0.y, acc1[1]);
      acc0[2] = fmaf(x0, w0.z, acc0[2]);   acc1[2] = fmaf(x1, w0.z, acc1[2]);
      acc0[3] = fmaf(x0, w0.w, acc0[3]);   acc1[3] = fmaf(x1, w0.w, acc1[3]);
      acc0[4] = fmaf(x0, w1.x, acc0[4]);   acc1[4] = fmaf(x1, w1.x, acc1[4]);
      acc0[5] = fmaf(x0, w1.y, acc0[5]);   acc1[5] = fmaf(x1, w1.y, acc1[5]);
      acc0[6] = fmaf(x0, w1.z, acc0[6]);   acc1[6] = fmaf(x1, w1.z, acc1[6]);
      acc0[7] = fmaf(x0, w1.w, acc0[7]);   acc1[7] = fmaf(x1, w1.w, acc1[7]);
      acc0[8] = fmaf(x0, w2.x, acc0[8]);   acc1[8] = fmaf(x1, w2.x, acc1[8]);
      acc0[9] = fmaf(x0, w2.y, acc0[9]);   acc1[9] = fmaf(x1, w2.y, acc1[9]);
      acc0[10] = fmaf(x0, w2.z, acc0[10]); acc1[10] = fmaf(x1, w2.z, acc1[10]);
      acc0[11] = fmaf(x0, w2.w, acc0[11]); acc1[11] = fmaf(x1, w2.w, acc1[11]);
      acc0[12] = fmaf(x0, w3.x, acc0[12]); acc1[12] = fmaf(x1, w3.x, acc1[12]);
      acc0[13] = fmaf(x0, w3.y, acc0[13]); acc1[13] = fmaf(x1, w3.y, acc1[13]);
      acc0[14] = fmaf(x0, w3.z, acc0[14]); acc1[14] = fmaf(x1, w3.z, acc1[14]);
      acc0[15] = fmaf(x0, w3.w, acc0[15]); acc1[15] = fmaf(x1, w3.w, acc1[15]);
    }
    float* o0 = Aout + (rg * 64 + lane) * XSout + ch * GNNFP_JC;
    float* o1 = o0 + 32 * XSout;
#pragma unroll
    for (int j = 0; j < GNNFP_JC; ++j) {
      o0[j] = act_fwd(act, acc0[j]);
      o1[j] = act_fwd(act, acc1[j]);
    }
  }
}

// row-wise softmax over the first H columns of the tile (Keras softmax, last axis)
__device__ __forceinline__ void softmax_rows(float* A, int XS, int H, int R) {
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float* row = A + r * XS;
    float m = row[0];
    for (int j = 1; j < H; ++j) m = fmaxf(m, row[j]);
    float s = 0.f;
    for (int j = 0; j < H; ++j) {
      const float e = expf(row[j] - m);
      row[j] = e;
      s += e;
    }
    for (int j = 0; j < H; ++j) row[j] = row[j] / s;
  }
}

struct FwdSmem {
  float* W[GNNFP_MAX_LAYERS];
  float* b[GNNFP_MAX_LAYERS];
  float* bnA;
  float* bnB;
  float* buf0;
  float* buf1;
  double* ost;   // [2*H] output statistics accumulators
};

__device__ __forceinline__ void carve_fwd(const NetDev& net, const TileCfg& tc, float* base, FwdSmem& s) {
  float* p = base;
  s.ost = reinterpret_cast<double*>(p);
  p += 4 * ceil_to(net.widths[net.n_layers - 1], 4);   // 2*H doubles
  int in_l = net.in_dim;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Hpad = ceil_to(net.widths[l], GNNFP_JC);
    s.W[l] = p;
    p += in_l * Hpad;
    s.b[l] = p;
    p += Hpad;
    in_l = net.widths[l];
  }
  s.bnA = p;
  p += ceil_to(net.in_dim, 4);
  s.bnB = p;
  p += ceil_to(net.in_dim, 4);
  s.buf0 = p;
  p += tc.R * tc.XS0;
  s.buf1 = p;
}

__global__ void __launch_bounds__(512) tile_fwd_kernel(const __grid_constant__ FwdArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ __align__(16) float smem[];
  const NetDev& net = a.net;
  const TileCfg& tc = a.tc;
  FwdSmem s;
  carve_fwd(net, tc, smem, s);
  const int tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int rg = warp % tc.RG, cg = warp / tc.RG;
  const int L = net.n_layers;
  const int H = net.widths[L - 1];

  // ---- weights (zero padded to 16 columns), BN coefficients, statistics accumulators ----------
  {
    int in_l = net.in_dim;
    for (int l = 0; l < L; ++l) {
      const int Hl = net.widths[l], Hpad = ceil_to(Hl, GNNFP_JC);
      for (int e = tid; e < in_l * Hpad; e += T) {
        const int c = e / Hpad, j = e - c * Hpad;
        s.W[l][e] = j < Hl ? net.W[l][(size_t)c * Hl + j] : 0.0f;
      }
      for (int j = tid; j < Hpad; j += T) s.b[l][j] = j < Hl ? net.b[l][j] : 0.0f;
      in_l = Hl;
    }
  }
  float* bnA = nullptr;
  float* bnB = nullptr;
  if (net.bn_mode) {
    bnA = s.bnA;
    bnB = s.bnB;
    bn_coefficients(a.src, net, 1, bnA, bnB, nullptr, nullptr);
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
  }
  for (int j = tid; j < 2 * H; j += T) s.ost[j] = 0.0;
  __syncthreads();

  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  int notconv = 0;
  const unsigned magicH = (unsigned)((0x100000000ull + (unsigned)H - 1) / (unsigned)H);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    stage_tile(a.src, row0, nr, tc.R, s.buf0, tc.XS0, bnA, bnB);
    __syncthreads();
    // ---- the MLP --------------------------------------------------------------------------
    float* cur = s.buf0;
    float* nxt = s.buf1;
    int XSc = tc.XS0, XSn = tc.XS1;
    int in_l = net.in_dim;
    for (int l = 0; l < L; ++l) {
      const int Hl = net.widths[l], Hpad = ceil_to(Hl, GNNFP_JC);
      if (cg < Hpad / GNNFP_JC)
        dense_layer_tile(cur, XSc, nxt, XSn, s.W[l], s.b[l], in_l, Hpad, net.acts[l], rg, cg, tc.CG, lane);
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
    for (int e = tid; e < nr * H; e += T) {
      const int r = (int)__umulhi((unsigned)e, magicH);
      const int j = e - r * H;
      const int orow = a.out_compact ? (row0 + r) : (a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r);
      a.out[(size_t)orow * a.ld_out + j] = cur[r * XSc + j];
    }
    if (a.prev) {   // GNN.py:200-209: sqrt(sum (s-s_old)^2) > thr * sqrt(sum s_old^2), strict
      for (int r = tid; r < nr; r += T) {
        const int gr = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
        const float* pv = a.prev + (size_t)gr * a.ld_prev;
        float sd = 0.f, sp = 0.f;
        for (int j = 0; j < H; ++j) {
          const float p = pv[j];
          const float d = cur[r * XSc + j] - p;
          sd = fmaf(d, d, sd);
          sp = fmaf(p, p, sp);
        }
        if (sqrtf(sd) > a.thr * sqrtf(sp)) notconv = 1;
      }
    }
    if (a.ost_sum) {
      for (int j = tid; j < H; j += T) {
        double su = 0.0, sq = 0.0;
        for (int r = 0; r < nr; ++r) {
          const double v = (double)cur[r * XSc + j];
          su += v;
          sq += v * v;
        }
        s.ost[j] += su;
        s.ost[H + j] += sq;
      }
    }
    __syncthreads();
  }
  if (a.flag_next) {
    const int any = __syncthreads_or(notconv);
    if (tid == 0 && any) atomicOr(a.flag_next, 1);
  }
  if (a.ost_sum) {
    for (int j = tid; j < H; j += T) {
      atomicAdd(a.ost_sum + j, s.ost[j]);
      atomicAdd(a.ost_sq + j, s.ost[H + j]);
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
  float* X = smem + 4 * ceil_to(W, 4);
  const int tid = threadIdx.x, T = blockDim.x;
  for (int j = tid; j < 2 * W; j += T) acc[j] = 0.0;
  __syncthreads();
  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  const unsigned magicW = (unsigned)((0x100000000ull + (unsigned)W - 1) / (unsigned)W);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    stage_tile(a.src, row0, nr, tc.R, X, tc.XS0, nullptr, nullptr);
    __syncthreads();
    if (a.out) {
      for (int e = tid; e < nr * W; e += T) {
        const int r = (int)__umulhi((unsigned)e, magicW);
        const int j = e - r * W;
        a.out[(size_t)(row0 + r) * a.ld_out + j] = X[r * tc.XS0 + j];
      }
    }
    if (a.st_sum) {
      for (int j = tid; j < W; j += T) {
        double su = 0.0, sq = 0.0;
        for (int r = 0; r < nr; ++r) {
          const double v = (double)X[r * tc.XS0 + j];
          su += v;
          sq += v * v;
        }
        acc[j] += su;
        acc[W + j] += sq;
      }
    }
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
// host side: tile geometry + launch
// ------------------------------------------------------------------------------------------------
static size_t fwd_smem_floats(const NetDev& net, int R, int XS0, int XS1) {
  size_t f = 4 * (size_t)ceil_to(net.widths[net.n_layers - 1], 4);
  int in_l = net.in_dim;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Hpad = ceil_to(net.widths[l], GNNFP_JC);
    f += (size_t)in_l * Hpad + Hpad;
    in_l = net.widths[l];
  }
  f += 2 * (size_t)ceil_to(net.in_dim, 4);
  f += (size_t)R * XS0 + (size_t)R * XS1;
  return f;
}

int tile_cfg_fwd(const NetDev& net, int n_rows, TileCfg* tc) {
  int w0 = net.in_dim, w1 = 1, hpmax = 0;
  for (int l = 0; l < net.n_layers; ++l) {
    const int Hpad = ceil_to(net.widths[l], GNNFP_JC);
    hpmax = Hpad > hpmax ? Hpad : hpmax;
    if (l % 2 == 0) w1 = Hpad > w1 ? Hpad : w1; else w0 = Hpad > w0 ? Hpad : w0;
  }
  tc->XS0 = odd_stride(w0);
  tc->XS1 = odd_stride(w1);
  int CG = hpmax / GNNFP_JC;
  if (CG > 8) CG = 8;
  int RG = 8 / CG;
  if (RG < 1) RG = 1;
  if (RG > 4) RG = 4;
  const int nsm = gnnfp_num_sms();
  const size_t cap = 200 * 1024, want = 100 * 1024;
  // shrink the tile until it fits twice per SM (or at all), and until the grid fills the GPU
  while (RG > 1 && (fwd_smem_floats(net, 64 * RG, tc->XS0, tc->XS1) * 4 > want ||
                    (n_rows + 64 * RG - 1) / (64 * RG) < 2 * nsm))
    RG /= 2;
  tc->RG = RG;
  tc->CG = CG;
  tc->R = 64 * RG;
  tc->threads = 32 * RG * CG;
  tc->smem_bytes = fwd_smem_floats(net, tc->R, tc->XS0, tc->XS1) * 4;
  if (tc->smem_bytes > cap)
    GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "net too large for the shared-memory tile kernel (%zu bytes needed)",
               tc->smem_bytes);
  int per_sm = (int)((220 * 1024) / (tc->smem_bytes + 1024));
  const int by_threads = 2048 / tc->threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  const int n_tiles = (n_rows + tc->R - 1) / tc->R;
  tc->grid = n_tiles < nsm * per_sm ? n_tiles : nsm * per_sm;
  if (tc->grid < 1) tc->grid = 1;
  return GNNFP_OK;
}

int tile_cfg_pass(int in_dim, int n_rows, TileCfg* tc) {
  tc->RG = 1;
  tc->CG = 1;
  tc->XS0 = odd_stride(in_dim);
  tc->XS1 = 1;
  tc->threads = 256;
  int R = 128;
  while (R > 16 && ((size_t)R * tc->XS0 + 4 * (size_t)ceil_to(in_dim, 4)) * 4 > 48 * 1024) R /= 2;
  tc->R = R;
  tc->smem_bytes = ((size_t)R * tc->XS0 + 4 * (size_t)ceil_to(in_dim, 4)) * 4;
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
