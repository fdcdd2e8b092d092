// kernels_bwd.cu - hand-written backward of one net application on a row set (sm_100a, FP32).
//
// One launch = the backward of one fixed-point iteration t (or of net_output): what
// tf.GradientTape replays for convergence() (reference GNN.py:217-236 under GNN.py:284-294):
//   G_t    = dL/ds_t               assembled in shared memory from "gradient pieces"
//            (dOwn_{t+1} + Adj . dAgg_{t+1} via the source-grouped CSR, or dL/ds_final)
//   dz     = G_t * act'(s_t)       derivative through the saved output
//   dW    += X^T dz, db += sum dz  per-CTA shared-memory accumulators -> per-CTA partial slots
//                                  (deterministic: no float atomics on parameters)
//   dX     = dz W^T                -> written to the gradient destinations of the input pieces
//                                  (dOwn_t, dAgg_t, static-column accumulators, d_nodes ...)
// BatchNormalization in training mode needs two batch reductions per application
// (sum dy, sum dy*x~): this kernel writes the per-CTA partial sums and a_c*dy; bn_reduce_kernel
// finishes the sums and tile_bnfix_kernel applies the (linear) correction
// dx -= a_c * (mean(dy) + x~ * mean(dy x~)).
#include "tile.cuh"

#ifdef GNNFP_PHASE_TIMING
__device__ long long g_phase_cycles[32];
extern "C" int gnnfp_debug_phases(long long* out, int reset) {
  cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(long long) * 32);
  if (reset) { long long z[32] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)); }
  return 0;
}
#endif

struct BwdLayout {
  int L, recompute, bn;
  int inw[GNNFP_MAX_LAYERS + 1];    // width of activation l (0 = input)
  int XSa[GNNFP_MAX_LAYERS + 1];
  int XSdA, XSdB;                   // dz ping-pong buffers: A holds widths inw[L], inw[L-2]..; B holds inw[L-1], inw[L-3]..
  int regacc;                       // single big layer: dW accumulators live in registers (no smem copy)
  int groups[GNNFP_MAX_LAYERS];
  int oWT[GNNFP_MAX_LAYERS], oWf[GNNFP_MAX_LAYERS], obf[GNNFP_MAX_LAYERS], oAct[GNNFP_MAX_LAYERS + 1];
  int oAccW[GNNFP_MAX_LAYERS], oAccb[GNNFP_MAX_LAYERS];
  int obnA, obnB, obnS, oZero, odzA, odzB, oAccBN, oScr, oRaw, oBar, oOutRaw;
  int total;   // floats
};

__host__ __device__ inline void bwd_layout(const NetDev& net, int R, int T, int regacc, int cap, int dz_ready, int raw_per_row, int out_per_row, BwdLayout& y) {
  y.L = net.n_layers;
  y.regacc = regacc;
  y.recompute = net.n_layers > 1;
  y.bn = net.bn_mode != 0;
  y.inw[0] = net.in_dim;
  for (int l = 0; l < y.L; ++l) y.inw[l + 1] = net.widths[l];
  int dmax = 0, dA = 0, dB = 0;
  for (int l = 0; l <= y.L; ++l) {
    const int p = ceil_to(y.inw[l], 16);
    y.XSa[l] = tile_stride(p);
    dmax = p > dmax ? p : dmax;
    if (((y.L - l) & 1) == 0) dA = p > dA ? p : dA; else dB = p > dB ? p : dB;
  }
  y.XSdA = tile_stride(dA);
  y.XSdB = tile_stride(dB > 0 ? dB : 16);
  int o = 0;
  for (int l = 0; l < y.L; ++l) {
    const int in_l = y.inw[l], H = y.inw[l + 1];
    y.oWT[l] = o; o += ceil_to(H, 4) * ceil_to(in_l, 16);           // W^T: [ceil4(H)][ceil16(in)]
    if (y.recompute) {
      y.oWf[l] = o; o += ceil_to(in_l, 4) * ceil_to(H, 16);         // forward weights [ceil4(in)][ceil16(H)]
      y.obf[l] = o; o += ceil_to(H, 16);
    } else { y.oWf[l] = 0; y.obf[l] = 0; }
    const int U = ((in_l + 7) / 8) * ((H + 3) / 4);
    int g = U >= T ? 1 : T / U;
    if (g > 16) g = 16;
    y.groups[l] = g;
    y.oAccW[l] = o; o += regacc ? 0 : ceil_to(g * in_l * H, 4);
    y.oAccb[l] = o; o += ceil_to(H, 4);
  }
  y.obnA = o; o += ceil_to(net.in_dim, 4);
  y.obnB = o; o += ceil_to(net.in_dim, 4);
  y.obnS = o; o += ceil_to(net.in_dim, 4);
  y.oAccBN = o; o += 2 * ceil_to(net.in_dim, 4);
  y.oZero = o; o += dmax;
  y.oScr = o; o += (int)scratch_floats(R, cap);
  for (int l = 0; l <= y.L; ++l) { y.oAct[l] = o; o += (dz_ready && l == y.L) ? 0 : R * y.XSa[l]; }   // dz path: the saved output is not needed
  y.odzA = o; o += R * y.XSdA;
  y.odzB = o; o += R * y.XSdB;
  y.oBar = o; o += 4;                                   // one 8-byte mbarrier (16-byte slot)
  y.oRaw = o; o += ceil_to(R * raw_per_row, 4);         // landing buffer of the bulk copies
  y.oOutRaw = o; o += ceil_to(R * out_per_row, 4);      // staging buffer of the bulk stores
  y.total = o;
}

template <bool REGACC>
__global__ void __launch_bounds__(256, REGACC ? 1 : 2) tile_bwd_kernel(const __grid_constant__ BwdArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ __align__(16) float smem[];
  const NetDev& net = a.net;
  const TileCfg& tc = a.tc;
  const int tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int rg = warp % tc.RG, cg = warp / tc.RG;
  __shared__ BwdLayout y;
  __shared__ float s_dbred[256];
  if (tid == 0) bwd_layout(net, tc.R, T, REGACC ? 1 : 0, tc.cap, a.dz_ready, tc.raw_per_row, tc.out_per_row, y);
  __syncthreads();
  const int L = y.L;
  float racc[2][8][4];
  if (REGACC) {
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) racc[q][i][jj] = 0.f;
  }
  float* bnA = smem + y.obnA;     // rstd            (x~ = x*bnA + bnB)
  float* bnB = smem + y.obnB;     // -mean*rstd
  float* bnS = smem + y.obnS;     // gamma*rstd      (dx = bnS * dy - correction)
  float* accBN = smem + y.oAccBN;
  float* zero = smem + y.oZero;
  const StageScratch sc = carve_scratch(smem + y.oScr, tc.R, tc.cap);

  if (y.bn) bn_coefficients(a.src, net, 0, bnA, bnB, nullptr, nullptr);
  __syncthreads();
  // ---- one-time staging: W^T (true weights), forward weights for the recompute, accumulators --------
  for (int l = 0; l < L; ++l) {
    const int in_l = y.inw[l], H = y.inw[l + 1], inpad = ceil_to(in_l, 16), Hpad = ceil_to(H, 16);
    float* WT = smem + y.oWT[l];
    for (int e = tid; e < ceil_to(H, 4) * inpad; e += T) {
      const int j = e / inpad, c = e - j * inpad;
      WT[e] = (c < in_l && j < H) ? net.W[l][(size_t)c * H + j] : 0.0f;
    }
    if (y.recompute) {
      // the tile holds RAW inputs: x_hat = (gamma*rstd) x + (gamma*(-mean*rstd) + beta) folded into layer 0
      float* Wf = smem + y.oWf[l];
      const bool fold = (l == 0 && y.bn);
      for (int e = tid; e < ceil_to(in_l, 4) * Hpad; e += T) {
        const int c = e / Hpad, j = e - c * Hpad;
        float w = (j < H && c < in_l) ? net.W[l][(size_t)c * H + j] : 0.0f;
        if (fold && c < in_l) w *= net.gamma[c] * bnA[c];
        Wf[e] = w;
      }
      float* bf = smem + y.obf[l];
      for (int j = tid; j < Hpad; j += T) {
        float b = j < H ? net.b[l][j] : 0.0f;
        if (fold && j < H)
          for (int c = 0; c < in_l; ++c) b = fmaf(fmaf(net.gamma[c], bnB[c], net.beta[c]), net.W[l][(size_t)c * H + j], b);
        bf[j] = b;
      }
    }
    if (!REGACC) {
      float* aw = smem + y.oAccW[l];
      for (int e = tid; e < y.groups[l] * in_l * H; e += T) aw[e] = 0.0f;
    }
    float* ab = smem + y.oAccb[l];
    for (int j = tid; j < H; j += T) ab[j] = 0.0f;
  }
  for (int c = tid; c < 2 * ceil_to(net.in_dim, 4); c += T) accBN[c] = 0.0f;
  for (int c = tid; c < y.oScr - y.oZero; c += T) zero[c] = 0.0f;
  for (int c = tid; c < y.oBar - y.oAct[0]; c += T) smem[y.oAct[0] + c] = 0.0f;   // activation + dz tiles
  if (y.bn)
    for (int c = tid; c < net.in_dim; c += T) bnS[c] = net.gamma[c] * bnA[c];
  __syncthreads();

  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  const int HL = y.inw[L];
  const unsigned magicHL = (unsigned)((0x100000000ull + (unsigned)HL - 1) / (unsigned)HL);
  float* dzA = smem + y.odzA;
  float* dzB = smem + y.odzB;
  const int XSdA = y.XSdA, XSdB = y.XSdB;

  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_end = min(n_tiles, ((int)blockIdx.x + 1) * tiles_per_cta);
  PHASE_INIT();
  const bool use_bulk = tc.raw_per_row > 0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + y.oBar);
  float* raw = smem + y.oRaw;
  uint32_t bar_phase = 0;
  if (use_bulk) {
    if (tid == 0) { mbar_init(bar, 1); fence_proxy_async(); }
    __syncthreads();
    const int t0 = blockIdx.x * tiles_per_cta;
    if (tid == 0 && t0 < tile_end && n - t0 * tc.R >= tc.R) {
      mbar_expect_tx(bar, bulk_bytes(a.src, tc.bulk_src, tc.R) + bulk_bytes(a.gsrc, tc.bulk_g, tc.R));
      float* rp = bulk_issue(a.src, tc.bulk_src, t0 * tc.R, tc.R, raw, bar);
      bulk_issue(a.gsrc, tc.bulk_g, t0 * tc.R, tc.R, rp, bar);
    }
  }
  for (int tile = blockIdx.x * tiles_per_cta; tile < tile_end; ++tile) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    PHASE_MARK(0);
    if (use_bulk && nr == tc.R) {
      mbar_wait(bar, bar_phase);
      bar_phase ^= 1u;
      PHASE_MARK(11);
      const float* rp = bulk_relayout(a.src, tc.bulk_src, tc.R, raw, smem + y.oAct[0], y.XSa[0]);
      PHASE_MARK(12);
      bulk_relayout(a.gsrc, tc.bulk_g, tc.R, rp, dzA, XSdA);
      PHASE_MARK(13);
      stage_tile(a.src, row0, nr, tc.R, smem + y.oAct[0], y.XSa[0], sc, true, tc.bulk_src);
      stage_tile(a.gsrc, row0, nr, tc.R, dzA, XSdA, sc, true, tc.bulk_g);
    } else {
      stage_tile(a.src, row0, nr, tc.R, smem + y.oAct[0], y.XSa[0], sc);
      stage_tile(a.gsrc, row0, nr, tc.R, dzA, XSdA, sc);
    }
    PHASE_MARK(1);
    if (!y.recompute && !a.dz_ready) {
      TileSrc so;
      so.n_rows = a.src.n_rows; so.rowlist = a.saved_compact ? nullptr : a.src.rowlist; so.n_pieces = 1; so.in_dim = HL;
      Piece& sp = so.p[0];
      sp = Piece();
      sp.ptr = a.saved_out; sp.ld = a.ld_saved; sp.width = HL; sp.col0 = 0; sp.kind = PK_DIRECT; sp.magic = magicHL;
      sp.compact = (a.saved_compact && a.src.rowlist) ? 1 : 0;
      stage_tile(so, row0, nr, tc.R, smem + y.oAct[L], y.XSa[L], sc);
    }
    __syncthreads();
    if (use_bulk && tid == 0 && tile + 1 < tile_end && n - (tile + 1) * tc.R >= tc.R) {
      fence_proxy_async();                    // landing buffer: generic-proxy reads above, async-proxy writes below
      mbar_expect_tx(bar, bulk_bytes(a.src, tc.bulk_src, tc.R) + bulk_bytes(a.gsrc, tc.bulk_g, tc.R));
      float* rp = bulk_issue(a.src, tc.bulk_src, (tile + 1) * tc.R, tc.R, raw, bar);
      bulk_issue(a.gsrc, tc.bulk_g, (tile + 1) * tc.R, tc.R, rp, bar);
    }
    PHASE_MARK(2);
    if (y.recompute) {
      for (int l = 0; l < L; ++l) {
        const int Hpad = ceil_to(y.inw[l + 1], 16);
        if (cg < Hpad / GNNFP_JC)
          dense_tile(smem + y.oAct[l], y.XSa[l], smem + y.oAct[l + 1], y.XSa[l + 1], smem + y.oWf[l], smem + y.obf[l],
                     (y.inw[l] + 3) / 4, Hpad, net.acts[l], rg, cg, tc.CG, lane);
        __syncthreads();
        if (net.acts[l] == GNNFP_ACT_SOFTMAX) {
          softmax_rows(smem + y.oAct[l + 1], y.XSa[l + 1], y.inw[l + 1], tc.R);
          __syncthreads();
        }
      }
    }
    // ---- dz_L = G * act'(h_L) ----------------------------------------------------------------
    if (!a.dz_ready) {
      const float* aL = smem + y.oAct[L];
      const int XS = y.XSa[L];
      const int actL = net.acts[L - 1];
      if (actL == GNNFP_ACT_SOFTMAX) {
        for (int r = tid; r < tc.R; r += T) {
          float dot = 0.f;
          for (int j = 0; j < HL; ++j) dot = fmaf(dzA[r * XSdA + j], aL[r * XS + j], dot);
          for (int j = 0; j < HL; ++j) dzA[r * XSdA + j] = aL[r * XS + j] * (dzA[r * XSdA + j] - dot);
        }
      } else {
        for (int e0 = tid; e0 < tc.R * HL; e0 += 4 * T) {
          float yv[4], gv[4];
          int o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * T;
            o[u] = -1; yv[u] = 0.f; gv[u] = 0.f;
            if (e < tc.R * HL) {
              const int r = div_magic((unsigned)e, magicHL);
              const int j = e - r * HL;
              o[u] = r * XSdA + j;
              yv[u] = aL[r * XS + j];
              gv[u] = dzA[o[u]];
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (o[u] >= 0) dzA[o[u]] = act_bwd(actL, yv[u], gv[u]);
        }
      }
    }
    __syncthreads();
    PHASE_MARK(3);
    float* cur = dzA;
    float* oth = dzB;
    int XSc = XSdA, XSo = XSdB;
    for (int l = L - 1; l >= 0; --l) {
      const int in_l = y.inw[l], H = y.inw[l + 1];
      const float* al = smem + y.oAct[l];
      const int XSl = y.XSa[l];
      // (a) dW_l += act_l^T . dz   (register tile 8 x 4 per unit; row groups when the layer is small)
      {
        const int in8 = (in_l + 7) / 8, h4 = (H + 3) / 4, U = in8 * h4;
        const int groups = y.groups[l];
        int u0, g, ustride;
        if (groups == 1) { u0 = tid; g = 0; ustride = T; }
        else { u0 = tid % U; g = tid / U; ustride = U; if (g >= groups) u0 = U; }
        float* aw = REGACC ? nullptr : smem + y.oAccW[l] + g * in_l * H;
        int q = 0;
        for (int u = u0; u < U; u += ustride, ++q) {
          const int cu = u / h4, ju = u - cu * h4;
          const int c0 = cu * 8, j0 = ju * 4;
          float acc[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
          const float* ap = al + c0;
          const float* dp = cur + j0;
#pragma unroll 2
          for (int r = g; r < nr; r += groups) {
            const float4 a0 = *reinterpret_cast<const float4*>(ap + r * XSl);
            const float4 a1 = *reinterpret_cast<const float4*>(ap + r * XSl + 4);
            const float4 d4 = *reinterpret_cast<const float4*>(dp + r * XSc);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(av[i], dv[jj], acc[i][jj]);
          }
          if (REGACC) {
            if (q == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) racc[0][i][jj] += acc[i][jj];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) racc[1][i][jj] += acc[i][jj];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (c0 + i < in_l && j0 + jj < H) aw[(c0 + i) * H + j0 + jj] += acc[i][jj];
          }
          if (groups > 1) break;
        }
        // db_l += column sums of dz (all threads: column x row group; the row groups are summed in a FIXED order through
        // shared memory - shared float atomics made the bias gradient the one run-to-run varying output of this kernel)
        float* ab = smem + y.oAccb[l];
        if (a.skip_bias && !y.bn) {
        } else if (T >= H) {
          const int ngr = T / H, j = tid % H, gg = tid / H;
          float s2 = 0.f;
          if (gg < ngr)
            for (int r = gg; r < nr; r += ngr) s2 += cur[r * XSc + j];
          s_dbred[tid] = s2;
          __syncthreads();
          if (tid < H) {
            float t2 = 0.f;
            for (int g2 = 0; g2 < ngr; ++g2) t2 += s_dbred[g2 * H + tid];
            ab[tid] += t2;
          }
          __syncthreads();
        } else {
          for (int j = tid; j < H; j += T) {
            float s2 = 0.f;
            for (int r = 0; r < nr; ++r) s2 += cur[r * XSc + j];
            ab[j] += s2;
          }
        }
      }
      PHASE_MARK(4);
      // (b) dprev = dz . W_l^T
      {
        const int inpad = ceil_to(in_l, 16);
        if (cg < inpad / GNNFP_JC)
          dense_tile(cur, XSc, oth, XSo, smem + y.oWT[l], zero, (H + 3) / 4, inpad, GNNFP_ACT_LINEAR, rg, cg, tc.CG, lane);
      }
      PHASE_MARK(5);
      __syncthreads();
      PHASE_MARK(6);
      if (l > 0) {
        const int actp = net.acts[l - 1];
        const unsigned magic = (unsigned)((0x100000000ull + (unsigned)in_l - 1) / (unsigned)in_l);
#pragma unroll 4
        for (int e = tid; e < tc.R * in_l; e += T) {
          const int r = div_magic((unsigned)e, magic);
          const int c = e - r * in_l;
          oth[r * XSo + c] = act_bwd(actp, al[r * XSl + c], oth[r * XSo + c]);
        }
        __syncthreads();
      }
      float* t2 = cur; cur = oth; oth = t2;
      const int t3 = XSc; XSc = XSo; XSo = t3;
    }
    // ---- cur = dy (gradient w.r.t. the BN output / the raw input) ---------------------------------
    PHASE_MARK(7);
    // (BN batch sums sum dy and sum dy*x need no pass over the tile: dy = dz W^T is linear, so they follow from
    //  db and the raw dW accumulator at flush time: sum_r dy[r][c] = sum_j W[c][j] db[j],  sum_r dy x = sum_j W[c][j] acc[c][j])
    PHASE_MARK(8);
    const bool bulk_out = tc.bulk_out != 0u && nr == tc.R;
    if (bulk_out) {
      // dense staging buffer <- scaled dy columns of every bulk-stored piece; one bulk store per piece
      if (tid == 0) bulk_wait_read0();              // the previous tile's stores have finished reading the buffer
      __syncthreads();
      float* ob = smem + y.oOutRaw;
      for (int p = 0; p < a.src.n_pieces; ++p) {
        if (!(tc.bulk_out & (1u << p))) continue;
        const Piece& pc = a.src.p[p];
        const int w = pc.width, n4 = tc.R * w / 4;
        float4* o4 = reinterpret_cast<float4*>(ob);
#pragma unroll 2
        for (int i = tid; i < n4; i += T) {
          int r = div_magic((unsigned)(4 * i), pc.magic);
          int c = 4 * i - r * w;
          float vv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float v = cur[r * XSc + pc.col0 + c];
            if (y.bn) v *= bnS[pc.col0 + c];
            vv[k] = v;
            if (++c == w) { c = 0; ++r; }
          }
          o4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        }
        ob += tc.R * w;
      }
      PHASE_MARK(14);
      fence_proxy_async();
      __syncthreads();
      PHASE_MARK(15);
      if (tid == 0) {
        const float* ob2 = smem + y.oOutRaw;
        for (int p = 0; p < a.src.n_pieces; ++p) {
          if (!(tc.bulk_out & (1u << p))) continue;
          const Piece& pc = a.src.p[p];
          bulk_s2g(pc.gptr + (size_t)row0 * pc.width, ob2, (uint32_t)tc.R * pc.width * 4u);
          ob2 += tc.R * pc.width;
        }
        bulk_commit();
      }
    }
    for (int p = 0; p < a.src.n_pieces; ++p) {
      const Piece& pc = a.src.p[p];
      if (pc.gmode == GM_NONE) continue;
      if (bulk_out && (tc.bulk_out & (1u << p))) continue;
      const int w = pc.width;
      for (int e0 = tid; e0 < nr * w; e0 += 4 * T) {
        float v[4], old[4];
        float* d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * T;
          d[u] = nullptr; v[u] = 0.f; old[u] = 0.f;
          if (e < nr * w) {
            const int r = div_magic((unsigned)e, pc.magic);
            const int c = e - r * w;
            v[u] = cur[r * XSc + pc.col0 + c];
            if (y.bn) v[u] *= bnS[pc.col0 + c];
            const int gr = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
            const int drow = (pc.map && !pc.gdirect) ? pc.map[gr] : gr;
            d[u] = pc.gptr + (size_t)drow * pc.gld + c;
            if (pc.gmode == GM_ADD) old[u] = *d[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (d[u]) {
            if (pc.gmode == GM_STORE) *d[u] = v[u];
            else if (pc.gmode == GM_ADD) *d[u] = old[u] + v[u];
            else atomicAdd(d[u], v[u]);
          }
      }
    }
    PHASE_MARK(9);
    __syncthreads();
    PHASE_MARK(10);
  }
  if (tc.bulk_out != 0u && tid == 0) bulk_wait0();
  // ---- flush per-CTA accumulators to this CTA's partial slot (plain +=: the slot is private) ----
  // With BN the layer-0 accumulators were taken on RAW inputs x; the gradient w.r.t. the Dense kernel is
  //   dW0[c][j] = sum x_hat dz = gamma_c*(rstd_c*acc[c][j] - mean_c*rstd_c*db[j]) + beta_c*db[j]
  // and sum dy*x~ = rstd_c*Qraw_c - mean_c*rstd_c*P_c  (this launch's batch statistics).
  {
    float* part = a.partial + (size_t)blockIdx.x * a.n_params;
    int off = 0;
    for (int l = 0; l < L; ++l) {
      const int in_l = y.inw[l], H = y.inw[l + 1];
      const float* ab = smem + y.oAccb[l];
      const bool fix = (l == 0 && y.bn);
      if (REGACC) {
        const int h4 = (H + 3) / 4, U = ((in_l + 7) / 8) * h4;
        int q = 0;
        for (int u = tid; u < U && q < 2; u += T, ++q) {
          const int cu = u / h4, ju = u - cu * h4;
          const int c0 = cu * 8, j0 = ju * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              if (c0 + i < in_l && j0 + jj < H) {
                const int c = c0 + i, j = j0 + jj;
                float v = (q == 0 ? racc[0][i][jj] : racc[1][i][jj]);
                if (fix) v = net.gamma[c] * fmaf(bnA[c], v, bnB[c] * ab[j]) + net.beta[c] * ab[j];
                part[off + c * H + j] += v;
              }
        }
      } else {
        const float* aw = smem + y.oAccW[l];
        for (int e = tid; e < in_l * H; e += T) {
          float v = 0.f;
          for (int g = 0; g < y.groups[l]; ++g) v += aw[g * in_l * H + e];
          if (fix) {
            const int c = e / H, j = e - c * H;
            v = net.gamma[c] * fmaf(bnA[c], v, bnB[c] * ab[j]) + net.beta[c] * ab[j];
          }
          part[off + e] += v;
        }
      }
      off += in_l * H;
      if (a.dz_ready) off = a.bias_off;
      if (!a.skip_bias)
        for (int j = tid; j < H; j += T) part[off + j] += ab[j];
      off += H;
    }
    if (y.bn && a.bn_partial) {
      const int tot = a.bn_in_total > 0 ? a.bn_in_total : net.in_dim;
      float* bp = a.bn_partial + (size_t)blockIdx.x * 2 * tot + a.bn_c_off;
      const int in0 = y.inw[0], H0 = y.inw[1];
      const float* ab0 = smem + y.oAccb[0];
      const float* WT0 = smem + y.oWT[0];
      const int inpad0 = ceil_to(in0, 16);
      if (REGACC) {
        // racc lives in registers of the owning threads: stage sum_j W[c][j]*acc[c][j] through accBN (zeroed at start)
        const int h4 = (H0 + 3) / 4, U = ((in0 + 7) / 8) * h4;
        int q = 0;
        for (int u = tid; u < U && q < 2; u += T, ++q) {
          const int cu = u / h4, ju = u - cu * h4;
          const int c0 = cu * 8, j0 = ju * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float sQ = 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              if (c0 + i < in0 && j0 + jj < H0) sQ = fmaf(WT0[(j0 + jj) * inpad0 + c0 + i], (q == 0 ? racc[0][i][jj] : racc[1][i][jj]), sQ);
            if (c0 + i < in0) atomicAdd(accBN + c0 + i, sQ);
          }
        }
        __syncthreads();
        for (int c = tid; c < in0; c += T) {
          float P = 0.f;
          for (int j = 0; j < H0; ++j) P = fmaf(WT0[j * inpad0 + c], ab0[j], P);
          const float Qraw = accBN[c];
          bp[c] = P;
          bp[tot + c] = fmaf(bnA[c], Qraw, bnB[c] * P);
        }
      } else {
        const float* aw0 = smem + y.oAccW[0];
        for (int c = tid; c < in0; c += T) {
          float P = 0.f, Qraw = 0.f;
          for (int j = 0; j < H0; ++j) {
            const float wv = WT0[j * inpad0 + c];
            float av = 0.f;
            for (int g = 0; g < y.groups[0]; ++g) av += aw0[g * in0 * H0 + c * H0 + j];
            P = fmaf(wv, ab0[j], P);
            Qraw = fmaf(wv, av, Qraw);
          }
          bp[c] = P;
          bp[tot + c] = fmaf(bnA[c], Qraw, bnB[c] * P);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// finish the BN batch reductions of one application: d_gamma += sum dy*x~, d_beta += sum dy,
// constants c0 = a_c*mean(dy), c1 = a_c*mean(dy*x~) for the fix-up.
struct BnReduceArgs {
  TileSrc src;
  NetDev net;
  const float* bn_partial;
  int grid;
  float* bn_grad;     // [2*in_dim] accumulators (gamma then beta)
  float* bn_const;    // [4*in_dim] c0 | c1 | rstd | -mean*rstd
  float* static_acc;  // optional [2*in_dim]: running sums of c0, c1 over the iterations (static columns)
  const int* gate;
};
// one block per 8 columns: thread (cx, gy) sums partial rows gy, gy+32, ... of column cx (up to 4 loads in flight),
// shared-memory reduction over the 32 row groups in fixed order (deterministic).
#define BNR_COLS 8
__global__ void __launch_bounds__(256) bn_reduce_kernel(const __grid_constant__ BnReduceArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ float sm[];
  __shared__ double redp[32][BNR_COLS], redq[32][BNR_COLS];
  const int in = a.net.in_dim;
  float* A = sm;
  float* Bc = sm + in;
  const int cx = threadIdx.x & (BNR_COLS - 1), gy = threadIdx.x / BNR_COLS;
  const int c = blockIdx.x * BNR_COLS + cx;
  double p = 0.0, q = 0.0;
  if (c < in) {                                        // issue the partial loads before the (fp64) coefficient math
    for (int b0 = gy; b0 < a.grid; b0 += 128) {
      float pv[4], qv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int b = b0 + 32 * u;
        pv[u] = b < a.grid ? a.bn_partial[(size_t)b * 2 * in + c] : 0.f;
        qv[u] = b < a.grid ? a.bn_partial[(size_t)b * 2 * in + in + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { p += (double)pv[u]; q += (double)qv[u]; }
    }
  }
  redp[gy][cx] = p;
  redq[gy][cx] = q;
  bn_coefficients(a.src, a.net, 0, A, Bc, nullptr, nullptr);
  __syncthreads();
  if (gy == 0 && c < in) {
    p = 0.0; q = 0.0;
    for (int g2 = 0; g2 < 32; ++g2) { p += redp[g2][cx]; q += redq[g2][cx]; }
    a.bn_grad[c] += (float)q;
    a.bn_grad[in + c] += (float)p;
    const float ac = a.net.gamma[c] * A[c];
    float k0 = 0.f, k1 = 0.f;
    if (a.net.bn_mode == 1) {
      k0 = (float)((double)ac * p * a.net.inv_n);
      k1 = (float)((double)ac * q * a.net.inv_n);
    }
    a.bn_const[c] = k0;
    a.bn_const[in + c] = k1;
    a.bn_const[2 * in + c] = A[c];
    a.bn_const[3 * in + c] = Bc[c];
    if (a.static_acc) { a.static_acc[c] += k0; a.static_acc[in + c] += k1; }
  }
}

// dx -= c0 + x~ * c1 on every gradient destination of the input pieces
struct BnFixArgs {
  TileSrc src;
  NetDev net;
  TileCfg tc;
  const float* bn_const;
  const int* gate;
};
__global__ void __launch_bounds__(256) tile_bnfix_kernel(const __grid_constant__ BnFixArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ __align__(16) float smem[];
  const TileCfg& tc = a.tc;
  const int in = a.net.in_dim, pin = ceil_to(in, 4);
  float* bnA = smem;
  float* bnB = smem + pin;
  float* c0 = smem + 2 * pin;
  float* c1 = smem + 3 * pin;
  const StageScratch sc = carve_scratch(smem + 4 * pin, tc.R, tc.cap);
  float* X = smem + 4 * pin + scratch_floats(tc.R, tc.cap);
  const int tid = threadIdx.x, T = blockDim.x;
  bn_coefficients(a.src, a.net, 0, bnA, bnB, nullptr, nullptr);
  for (int c = tid; c < in; c += T) { c0[c] = a.bn_const[c]; c1[c] = a.bn_const[in + c]; }
  __syncthreads();
  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_end = min(n_tiles, ((int)blockIdx.x + 1) * tiles_per_cta);
  for (int tile = blockIdx.x * tiles_per_cta; tile < tile_end; ++tile) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    stage_tile(a.src, row0, nr, tc.R, X, tc.XS0, sc);
    __syncthreads();
    for (int p = 0; p < a.src.n_pieces; ++p) {
      const Piece& pc = a.src.p[p];
      if (pc.gmode == GM_NONE) continue;
      const int w = pc.width;
      for (int e0 = tid; e0 < nr * w; e0 += 4 * T) {      // 4 read-modify-writes in flight per thread
        float* dp[4];
        float corr[4], old[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = e0 + u * T;
          dp[u] = nullptr;
          corr[u] = 0.f; old[u] = 0.f;
          if (e < nr * w) {
            const int r = div_magic((unsigned)e, pc.magic);
            const int c = e - r * w;
            const int cc = pc.col0 + c;
            corr[u] = c0[cc] + fmaf(X[r * tc.XS0 + cc], bnA[cc], bnB[cc]) * c1[cc];
            const int gr = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
            const int drow = (pc.map && !pc.gdirect) ? pc.map[gr] : gr;
            dp[u] = pc.gptr + (size_t)drow * pc.gld + c;
            if (pc.gmode != GM_ATOMIC) old[u] = *dp[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (dp[u]) { if (pc.gmode == GM_ATOMIC) atomicAdd(dp[u], -corr[u]); else *dp[u] = old[u] - corr[u]; }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// d_params = scale * sum over CTAs of the partial slots (double accumulation, fixed order).
// With BN the layer-0 accumulator holds M = X~^T dz: dW0 = gamma (.) M + beta (x) db0.
struct ReduceArgs {
  NetDev net;            // widths + gamma/beta (device pointers)
  const float* partial;
  int grid, n_params;
  const float* bn_grad;  // [2*in_dim] or NULL
  float* dW[GNNFP_MAX_LAYERS];
  float* db[GNNFP_MAX_LAYERS];
  float* dgamma;
  float* dbeta;
  const int* flags;      // iteration flags -> k
  int max_iter;
  int average;           // divide by k (GNN.py:295)
};
__global__ void reduce_params_kernel(const __grid_constant__ ReduceArgs a) {
  float scale = 1.0f;
  if (a.average) {
    int k = 0;
    for (int t = 0; t < a.max_iter; ++t) k += a.flags[t] != 0;
    scale = 1.0f / (float)k;   // k == 0 -> inf, as dwbS / k in the reference
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int bn = a.net.bn_mode != 0;
  if (i < a.n_params) {
    double s = 0.0;
    for (int b = 0; b < a.grid; ++b) s += (double)a.partial[(size_t)b * a.n_params + i];
    // locate (layer, element)
    int off = 0, in_l = a.net.in_dim;
    for (int l = 0; l < a.net.n_layers; ++l) {
      const int H = a.net.widths[l];
      if (i < off + in_l * H) {
        const int e = i - off, c = e / H, j = e - c * H;
        (void)c; (void)j;
        a.dW[l][e] = (float)s * scale;
        break;
      }
      off += in_l * H;
      if (i < off + H) { a.db[l][i - off] = (float)s * scale; break; }
      off += H;
      in_l = H;
    }
  }
  if (bn && a.bn_grad && i < a.net.in_dim) {
    a.dgamma[i] = a.bn_grad[i] * scale;
    a.dbeta[i] = a.bn_grad[a.net.in_dim + i] * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// dz = act'(s_t) * G_t as a streaming kernel (single-layer state nets): the gather over the source-grouped
// CSR runs with thousands of independent threads instead of inside the persistent GEMM kernel.
// PAD (VEC == 4, D % 4 == 2, every leading dimension a multiple of 4): rows are walked as ceil(D / 4) float4 slots; the last
// slot reads two columns of padding / of the neighbouring block (finite values) and writes zeros into dz's padding - half
// the threads and index arithmetic of the float2 walk.  agg_next (the Adj^T S block of an X slot) starts 8-byte aligned.
#ifndef DZ_MINB
#define DZ_MINB 5
#endif
template <int VEC, bool PAD>
__global__ void __launch_bounds__(256, DZ_MINB) dz_kernel(const __grid_constant__ DzArgs a) {
  if (a.gate && *a.gate == 0) return;
  const bool last = a.always_last || a.last_flag == nullptr || *a.last_flag == 0;
  const int nq = PAD ? (a.D + 3) / 4 : a.D / VEC;
  const int ldg = a.ldg ? a.ldg : a.D, lda = a.ld_agg ? a.ld_agg : a.D, ldz = a.ld_dz ? a.ld_dz : ldg;
  // thread = (column chunk q, rows r0, r0 + rows_per_pass, ...): the chunk is fixed per thread, so the BN constants of its
  // columns are loaded ONCE (the kernel is bound by the L1 data path: ncu l1tex 76 % busy, 80 % hits - reloading 8 constants
  // per item was a third of its sectors)
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows_per_pass = (int)(((long long)gridDim.x * blockDim.x) / nq);
  const int q = gtid % nq, r0 = gtid / nq;
  if (r0 >= rows_per_pass) return;
  float k0o[VEC], k1o[VEC], Ao[VEC], Bo[VEC], k0a[VEC], k1a[VEC], Aa[VEC], Ba[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { k0o[v] = k1o[v] = Ao[v] = Bo[v] = k0a[v] = k1a[v] = Aa[v] = Ba[v] = 0.f; }
  if (a.cn && !last) {
    const int in = a.in_dim;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int cv = (PAD && q * VEC + v >= a.D) ? a.D - 1 : q * VEC + v;
      const int oc = a.own_col0 + cv, ac = a.agg_col0 + cv;
      k0o[v] = a.cn[oc]; k1o[v] = a.cn[in + oc]; Ao[v] = a.cn[2 * in + oc]; Bo[v] = a.cn[3 * in + oc];
      k0a[v] = a.cn[ac]; k1a[v] = a.cn[in + ac]; Aa[v] = a.cn[2 * in + ac]; Ba[v] = a.cn[3 * in + ac];
    }
  }
  for (int r = r0; r < a.n_rows; r += rows_per_pass) {
    const int gr = a.rowlist ? a.rowlist[r] : r;
    const size_t o = (size_t)gr * ldg + q * VEC;
    float g[VEC], yv[VEC];
    load_vec<VEC>(a.s_t + (size_t)gr * a.ld_s + q * VEC, yv);
    if (last) {
      load_vec<VEC>(a.dSfin + o, g);
    } else {
      load_vec<VEC>(a.dOwn + o, g);
      const float* cn = a.cn;
      if (cn) {   // BN-training correction of iteration t+1's raw gradients, applied where they are consumed:
                  // dx = a dy - (c0 + x~ c1),  x~ = x*rstd - mean*rstd
#pragma unroll
        for (int v = 0; v < VEC; ++v) g[v] -= k0o[v] + fmaf(yv[v], Ao[v], Bo[v]) * k1o[v];
      }
      if (a.pre) {
        float pv[VEC];
        load_vec<VEC>(a.pre + o, pv);
#pragma unroll
        for (int v = 0; v < VEC; ++v) g[v] += pv[v];
      }
      const int a0 = a.pre ? 0 : a.rowptr[gr], a1 = a.pre ? 0 : a.rowptr[gr + 1];
      for (int p = a0; p < a1; p += 4) {
        float t[4][VEC], xg[4][VEC], wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = p + u < a1;
          wv[u] = 0.0f;
#pragma unroll
          for (int v = 0; v < VEC; ++v) { t[u][v] = 0.f; xg[u][v] = 0.f; }
          if (ok) {   // predicated loads: an absent arc must not cost L1 sectors (the kernel is bound by the L1 data path)
            const int pi = p + u;
            const int di = a.idx[pi];
            wv[u] = a.wgt ? a.wgt[pi] : 1.0f;
            load_vec<VEC>(a.dAgg + (size_t)di * ldg + q * VEC, t[u]);
            if (cn) {
              if (PAD) {
                load_vec<2>(a.agg_next + (size_t)di * lda + q * VEC, xg[u]);
                load_vec<2>(a.agg_next + (size_t)di * lda + q * VEC + 2, xg[u] + (VEC > 2 ? 2 : 0));
              } else {
                load_vec<VEC>(a.agg_next + (size_t)di * lda + q * VEC, xg[u]);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (p + u < a1) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              float tv = t[u][v];
              if (cn) tv -= k0a[v] + fmaf(xg[u][v], Aa[v], Ba[v]) * k1a[v];
              g[v] = fmaf(wv[u], tv, g[v]);
            }
          }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) a.dz[(size_t)gr * ldz + q * VEC + v] = (PAD && q * VEC + v >= a.D) ? 0.f : act_bwd(a.act, yv[v], g[v]);
  }
}

static bool dz_pad4_enabled() { static const bool on = getenv("GNNFP_DZ_PAD4") != nullptr; return on; }
int launch_dz(const DzArgs& a, cudaStream_t s) {
  if (a.n_rows <= 0) return GNNFP_OK;
  auto al = [&](const void* p, int m) { return (reinterpret_cast<uintptr_t>(p) & (m - 1)) == 0; };
  int vec = 1;
  const int ldg = a.ldg ? a.ldg : a.D, lda = a.ld_agg ? a.ld_agg : a.D;
  const int ldz = a.ld_dz ? a.ld_dz : ldg;
  const bool g4 = ldg % 4 == 0 && ldz % 4 == 0 && (!a.agg_next || (lda % 4 == 0 && al(a.agg_next, 16))) && (!a.pre || al(a.pre, 16));
  const bool g2 = ldg % 2 == 0 && ldz % 2 == 0 && (!a.agg_next || (lda % 2 == 0 && al(a.agg_next, 8))) && (!a.pre || al(a.pre, 8));
  bool pad = false;
  const bool a16 = a.ld_s % 4 == 0 && al(a.s_t, 16) && al(a.dSfin, 16) && al(a.dOwn, 16) && al(a.dAgg, 16) && al(a.dz, 16);
  if (g4 && a.D % 4 == 0 && a16) vec = 4;
  // (the padded float4 walk measured 1.5x SLOWER than the float2 walk on B200 - the kernel is bound by loads in flight, not
  //  by instruction issue - so it is only taken on request: GNNFP_DZ_PAD4=1)
  else if (dz_pad4_enabled() && a.D % 4 == 2 && a16 && ldg % 4 == 0 && ldz % 4 == 0 && (!a.pre || al(a.pre, 16)) &&
           (!a.agg_next || (lda % 2 == 0 && al(a.agg_next, 8)))) { vec = 4; pad = true; }
  else if (g2 && a.D % 2 == 0 && a.ld_s % 2 == 0 && al(a.s_t, 8) && al(a.dSfin, 8) && al(a.dOwn, 8) && al(a.dAgg, 8) && al(a.dz, 8)) vec = 2;
  const long long items = (long long)a.n_rows * (pad ? (a.D + 3) / 4 : a.D / vec);
  long long blocks = (items + 255) / 256;
  static int occ[3] = {0, 0, 0};                      // resident blocks per SM: the grid is a whole number of waves
  const int oi = vec == 4 ? 2 : (vec == 2 ? 1 : 0);
  if (!occ[oi]) {
    int o = 0;
    if (vec == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dz_kernel<4, false>, 256, 0);
    else if (vec == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dz_kernel<2, false>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, dz_kernel<1, false>, 256, 0);
    occ[oi] = o > 0 ? o : 4;
  }
  const long long cap = (long long)gnnfp_num_sms() * occ[oi] * 3;
  if (blocks > cap) blocks = cap;
  ProfScope ps(PC_DZ, s);
  if (pad) dz_kernel<4, true><<<(int)blocks, 256, 0, s>>>(a);
  else if (vec == 4) dz_kernel<4, false><<<(int)blocks, 256, 0, s>>>(a);
  else if (vec == 2) dz_kernel<2, false><<<(int)blocks, 256, 0, s>>>(a);
  else dz_kernel<1, false><<<(int)blocks, 256, 0, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// ------------------------------------------------------------------------------------------------
int tile_cfg_bwd(const NetDev& net, int n_rows, int gwidth, TileCfg* tc) {
  (void)gwidth;
  const int dzr = tc->dz_ready;
  int maxch = 1;
  for (int l = 0; l < net.n_layers; ++l) {
    const int in_l = l == 0 ? net.in_dim : net.widths[l - 1];
    const int c1 = ceil_to(in_l, 16) / 16, c2 = ceil_to(net.widths[l], 16) / 16;
    maxch = c1 > maxch ? c1 : maxch;
    if (net.n_layers > 1) maxch = c2 > maxch ? c2 : maxch;
  }
  int CG = maxch >= 8 ? 8 : (maxch >= 4 ? 4 : (maxch >= 2 ? 2 : 1));
  int RG = 8 / CG;
  const int nsm = gnnfp_num_sms();
  BwdLayout y;
  const size_t cap = 216 * 1024, want = 100 * 1024;
  const int cap_per_row = tc->cap_per_row > 0 ? tc->cap_per_row : 4;
  int regacc = 0;
  for (;;) {
    bwd_layout(net, 64 * RG, 256, 0, 64 * RG * cap_per_row, dzr, tc->raw_per_row, tc->out_per_row, y);
    const bool too_big = (size_t)y.total * 4 > want;
    const bool underfill = (n_rows + 64 * RG - 1) / (64 * RG) < 2 * nsm;
    if (RG > 1 && (too_big || underfill)) { RG /= 2; CG = 8 / RG; continue; }
    break;
  }
  if (net.n_layers == 1) {
    // one Dense layer with 128 < units <= 512: keep the dW accumulators in registers (<= 2 units of 8x4 per thread)
    const int U = ((net.in_dim + 7) / 8) * ((net.widths[0] + 3) / 4);
    if (U > 128 && U <= 512) {
      regacc = 1;
      RG = 8 / CG;
      for (;;) {
        bwd_layout(net, 64 * RG, 256, 1, 64 * RG * cap_per_row, dzr, tc->raw_per_row, tc->out_per_row, y);
        const bool too_big = (size_t)y.total * 4 > want;
        const bool underfill = (n_rows + 64 * RG - 1) / (64 * RG) < 2 * nsm;
        if (RG > 1 && (too_big || underfill)) { RG /= 2; CG = 8 / RG; continue; }
        break;
      }
    }
  }
  tc->RG = RG; tc->CG = CG; tc->R = 64 * RG; tc->threads = 256;
  tc->XS0 = y.XSa[0]; tc->XS1 = regacc;   // XS1 doubles as the register-accumulation switch of the backward kernel
  tc->cap = tc->R * cap_per_row;
  tc->smem_bytes = (size_t)y.total * 4 + 64;
  if (tc->smem_bytes > cap)
    GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "net too large for the shared-memory backward tile kernel (%zu bytes needed)", tc->smem_bytes);
  int per_sm = (int)((224 * 1024) / (tc->smem_bytes + 1024));
  if (per_sm > 4) per_sm = 4;      // partial-sum slots are sized for 4 CTAs per SM (loop.cu grid_cap)
  if (per_sm < 1) per_sm = 1;
  const int n_tiles = (n_rows + tc->R - 1) / tc->R;
  tc->grid = n_tiles < nsm * per_sm ? n_tiles : nsm * per_sm;
  if (tc->grid < 1) tc->grid = 1;
  return GNNFP_OK;
}

int launch_tile_bwd(const BwdArgs& a, cudaStream_t s) {
  if (a.src.n_rows <= 0) return GNNFP_OK;
  static bool attr_set = false;
  if (!attr_set) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(tile_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(tile_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(tile_bnfix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(100 * 1024)));
    attr_set = true;
  }
  ProfScope ps(a.prof_cat ? a.prof_cat : PC_OTHER, s);
  if (a.tc.XS1) tile_bwd_kernel<true><<<a.tc.grid, a.tc.threads, a.tc.smem_bytes, s>>>(a);
  else tile_bwd_kernel<false><<<a.tc.grid, a.tc.threads, a.tc.smem_bytes, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// BN tail of one backward application: reduce + fix
int launch_bn_tail(const BwdArgs& a, float* bn_grad, float* bn_const, cudaStream_t s, int no_fix, float* static_acc) {
  if (a.src.n_rows <= 0) return GNNFP_OK;
  BnReduceArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.src = a.src; ra.net = a.net; ra.bn_partial = a.bn_partial; ra.grid = a.tc.grid;
  ra.bn_grad = bn_grad; ra.bn_const = bn_const; ra.gate = a.gate; ra.static_acc = static_acc;
  bn_reduce_kernel<<<(a.net.in_dim + BNR_COLS - 1) / BNR_COLS, 256, 2 * a.net.in_dim * sizeof(float), s>>>(ra);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  if (a.net.bn_mode != 1 || no_fix) return GNNFP_OK;
  bool any = false;
  for (int p = 0; p < a.src.n_pieces; ++p) any = any || a.src.p[p].gmode != GM_NONE;
  if (!any) return GNNFP_OK;
  BnFixArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.src = a.src; fa.net = a.net; fa.bn_const = bn_const; fa.gate = a.gate;
  fa.tc.cap_per_row = a.tc.cap_per_row;
  int rc = tile_cfg_pass(a.net.in_dim, a.src.n_rows, &fa.tc);
  if (rc) return rc;
  fa.tc.smem_bytes += 4 * (size_t)ceil_to(a.net.in_dim, 4) * sizeof(float);
  ProfScope ps(PC_BNFIX, s);
  tile_bnfix_kernel<<<fa.tc.grid, fa.tc.threads, fa.tc.smem_bytes, s>>>(fa);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

int launch_reduce_params(const NetDev& net, const float* partial, int grid, int n_params, const float* bn_grad,
                         const gnnfp_net_params& d, const int* flags, int max_iter, int average, cudaStream_t s) {
  ReduceArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.net = net; ra.partial = partial; ra.grid = grid; ra.n_params = n_params; ra.bn_grad = bn_grad;
  for (int l = 0; l < net.n_layers; ++l) { ra.dW[l] = d.W[l]; ra.db[l] = d.b[l]; }
  ra.dgamma = d.bn_gamma; ra.dbeta = d.bn_beta;
  ra.flags = flags; ra.max_iter = max_iter; ra.average = average;
  const int n = n_params > net.in_dim ? n_params : net.in_dim;
  reduce_params_kernel<<<(n + 127) / 128, 128, 0, s>>>(ra);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
