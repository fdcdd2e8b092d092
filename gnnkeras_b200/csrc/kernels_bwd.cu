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

struct BwdLayout {
  int L, recompute, bn;
  int inw[GNNFP_MAX_LAYERS + 1];    // width of activation l (0 = input)
  int XSa[GNNFP_MAX_LAYERS + 1];
  int XSdA, XSdB;                   // dz ping-pong buffers: A holds widths inw[L], inw[L-2]..; B holds inw[L-1], inw[L-3]..
  int regacc;                       // single big layer: dW accumulators live in registers (no smem copy)
  int groups[GNNFP_MAX_LAYERS];
  size_t oWT[GNNFP_MAX_LAYERS], oWf[GNNFP_MAX_LAYERS], obf[GNNFP_MAX_LAYERS], oAct[GNNFP_MAX_LAYERS + 1];
  size_t oAccW[GNNFP_MAX_LAYERS], oAccb[GNNFP_MAX_LAYERS];
  size_t obnA, obnB, obnS, oZero, odzA, odzB, oAccBN;
  size_t total;   // floats
};

__host__ __device__ inline void bwd_layout(const NetDev& net, int R, int T, int regacc, BwdLayout& y) {
  y.L = net.n_layers;
  y.regacc = regacc;
  y.recompute = net.n_layers > 1;
  y.bn = net.bn_mode != 0;
  y.inw[0] = net.in_dim;
  for (int l = 0; l < y.L; ++l) y.inw[l + 1] = net.widths[l];
  int dmax = 0, dA = 0, dB = 0;
  for (int l = 0; l <= y.L; ++l) {
    y.XSa[l] = odd_stride(ceil_to(y.inw[l], 16));
    const int p = ceil_to(y.inw[l], 16);
    dmax = p > dmax ? p : dmax;
    if (((y.L - l) & 1) == 0) dA = p > dA ? p : dA; else dB = p > dB ? p : dB;
  }
  y.XSdA = odd_stride(dA);
  y.XSdB = odd_stride(dB > 0 ? dB : 16);
  size_t o = 0;
  for (int l = 0; l < y.L; ++l) {
    const int in_l = y.inw[l], H = y.inw[l + 1];
    y.oWT[l] = o; o += (size_t)H * ceil_to(in_l, 16);
    if (y.recompute) {
      y.oWf[l] = o; o += (size_t)in_l * ceil_to(H, 16);
      y.obf[l] = o; o += ceil_to(H, 16);
    } else { y.oWf[l] = 0; y.obf[l] = 0; }
    const int U = ((in_l + 7) / 8) * ((H + 3) / 4);
    int g = U >= T ? 1 : T / U;
    if (g > 16) g = 16;
    y.groups[l] = g;
    y.oAccW[l] = o; o += regacc ? 0 : (size_t)ceil_to(g * in_l * H, 4);
    y.oAccb[l] = o; o += ceil_to(H, 4);
  }
  y.obnA = o; o += ceil_to(net.in_dim, 4);
  y.obnB = o; o += ceil_to(net.in_dim, 4);
  y.obnS = o; o += ceil_to(net.in_dim, 4);
  y.oAccBN = o; o += 2 * (size_t)ceil_to(net.in_dim, 4);
  y.oZero = o; o += dmax;
  for (int l = 0; l <= y.L; ++l) { y.oAct[l] = o; o += (size_t)R * y.XSa[l]; }
  y.odzA = o; o += (size_t)R * y.XSdA;
  y.odzB = o; o += (size_t)R * y.XSdB;
  y.total = o;
}

// forward-style dense layer used for both the recompute and dprev = dz . W^T (no activation, bias from `bl`)
__device__ __forceinline__ void dense_tile(const float* __restrict__ Ain, int XSin, float* __restrict__ Aout, int XSout,
                                           const float* __restrict__ Wl, const float* __restrict__ bl, int in_l, int Hpad,
                                           int act, int rg, int cg, int CG, int lane) {
  const int nch = Hpad / GNNFP_JC;
  const float* x0p = Ain + (rg * 64 + lane) * XSin;
  const float* x1p = x0p + 32 * XSin;
  for (int ch = cg; ch < nch; ch += CG) {
    float acc0[GNNFP_JC], acc1[GNNFP_JC];
#pragma unroll
    for (int j = 0; j < GNNFP_JC; ++j) { const float bj = bl[ch * GNNFP_JC + j]; acc0[j] = bj; acc1[j] = bj; }
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
    for (int j = 0; j < GNNFP_JC; ++j) { o0[j] = act_fwd(act, acc0[j]); o1[j] = act_fwd(act, acc1[j]); }
  }
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
  BwdLayout y;
  bwd_layout(net, tc.R, T, REGACC ? 1 : 0, y);
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
  float* bnA = y.bn ? smem + y.obnA : nullptr;
  float* bnB = y.bn ? smem + y.obnB : nullptr;
  float* bnS = smem + y.obnS;
  float* accBN = smem + y.oAccBN;
  float* zero = smem + y.oZero;

  // ---- one-time staging: W^T (true weights), forward weights for the recompute, BN, accumulators ----
  for (int l = 0; l < L; ++l) {
    const int in_l = y.inw[l], H = y.inw[l + 1], inpad = ceil_to(in_l, 16), Hpad = ceil_to(H, 16);
    float* WT = smem + y.oWT[l];
    for (int e = tid; e < H * inpad; e += T) {
      const int j = e / inpad, c = e - j * inpad;
      WT[e] = c < in_l ? net.W[l][(size_t)c * H + j] : 0.0f;
    }
    if (y.recompute) {
      float* Wf = smem + y.oWf[l];
      for (int e = tid; e < in_l * Hpad; e += T) {
        const int c = e / Hpad, j = e - c * Hpad;
        float w = j < H ? net.W[l][(size_t)c * H + j] : 0.0f;
        if (l == 0 && y.bn) w *= net.gamma[c];               // act0 holds x~: fold gamma/beta into layer 0
        Wf[e] = w;
      }
      float* bf = smem + y.obf[l];
      for (int j = tid; j < Hpad; j += T) {
        float b = j < H ? net.b[l][j] : 0.0f;
        if (l == 0 && y.bn && j < H)
          for (int c = 0; c < in_l; ++c) b = fmaf(net.beta[c], net.W[l][(size_t)c * H + j], b);
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
  if (y.bn) {
    bn_coefficients(a.src, net, 0, bnA, bnB, nullptr, nullptr);
  }
  for (int c = tid; c < 2 * ceil_to(net.in_dim, 4); c += T) accBN[c] = 0.0f;
  for (int c = tid; c < (int)(y.oAct[0] - y.oZero); c += T) zero[c] = 0.0f;
  __syncthreads();
  if (y.bn)
    for (int c = tid; c < net.in_dim; c += T) bnS[c] = net.gamma[c] * bnA[c];   // a_c = gamma * rstd
  __syncthreads();

  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  const int HL = y.inw[L];
  const unsigned magicHL = (unsigned)((0x100000000ull + (unsigned)HL - 1) / (unsigned)HL);
  float* dzA = smem + y.odzA;
  float* dzB = smem + y.odzB;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    stage_tile(a.src, row0, nr, tc.R, smem + y.oAct[0], y.XSa[0], bnA, bnB);
    stage_tile(a.gsrc, row0, nr, tc.R, dzA, y.XSdA, nullptr, nullptr);
    if (!y.recompute) {
      float* aL = smem + y.oAct[L];
      for (int e = tid; e < tc.R * HL; e += T) {
        const int r = (int)__umulhi((unsigned)e, magicHL);
        const int j = e - r * HL;
        float v = 0.0f;
        if (r < nr) {
          const int srow = a.saved_compact ? (row0 + r) : (a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r);
          v = a.saved_out[(size_t)srow * a.ld_saved + j];
        }
        aL[r * y.XSa[L] + j] = v;
      }
    }
    __syncthreads();
    if (y.recompute) {
      for (int l = 0; l < L; ++l) {
        const int Hpad = ceil_to(y.inw[l + 1], 16);
        if (cg < Hpad / GNNFP_JC)
          dense_tile(smem + y.oAct[l], y.XSa[l], smem + y.oAct[l + 1], y.XSa[l + 1], smem + y.oWf[l], smem + y.obf[l],
                     y.inw[l], Hpad, net.acts[l], rg, cg, tc.CG, lane);
        __syncthreads();
        if (net.acts[l] == GNNFP_ACT_SOFTMAX) {
          float* A = smem + y.oAct[l + 1];
          const int XS = y.XSa[l + 1], H = y.inw[l + 1];
          for (int r = tid; r < tc.R; r += T) {
            float* row = A + r * XS;
            float m = row[0];
            for (int j = 1; j < H; ++j) m = fmaxf(m, row[j]);
            float s = 0.f;
            for (int j = 0; j < H; ++j) { const float e2 = expf(row[j] - m); row[j] = e2; s += e2; }
            for (int j = 0; j < H; ++j) row[j] = row[j] / s;
          }
          __syncthreads();
        }
      }
    }
    // ---- dz_L = G * act'(h_L) ----------------------------------------------------------------
    {
      const float* aL = smem + y.oAct[L];
      const int XS = y.XSa[L];
      const int actL = net.acts[L - 1];
      if (actL == GNNFP_ACT_SOFTMAX) {
        for (int r = tid; r < tc.R; r += T) {
          float dot = 0.f;
          for (int j = 0; j < HL; ++j) dot = fmaf(dzA[r * y.XSdA + j], aL[r * XS + j], dot);
          for (int j = 0; j < HL; ++j) dzA[r * y.XSdA + j] = aL[r * XS + j] * (dzA[r * y.XSdA + j] - dot);
        }
      } else {
        for (int e = tid; e < tc.R * HL; e += T) {
          const int r = (int)__umulhi((unsigned)e, magicHL);
          const int j = e - r * HL;
          dzA[r * y.XSdA + j] = act_bwd(actL, aL[r * XS + j], dzA[r * y.XSdA + j]);
        }
      }
    }
    __syncthreads();
    float* cur = dzA;
    float* oth = dzB;
    int XSc = y.XSdA, XSo = y.XSdB;
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
        float* aw = REGACC ? nullptr : smem + y.oAccW[l] + (size_t)g * in_l * H;
        int q = 0;
        for (int u = u0; u < U; u += ustride, ++q) {
          const int cu = u / h4, ju = u - cu * h4;
          const int c0 = cu * 8, j0 = ju * 4;
          float acc[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
          for (int r = g; r < nr; r += groups) {
            float av[8], dv[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = al[r * XSl + c0 + i];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) dv[jj] = cur[r * XSc + j0 + jj];
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
        float* ab = smem + y.oAccb[l];
        for (int j = tid; j < H; j += T) {
          float s = 0.f;
          for (int r = 0; r < nr; ++r) s += cur[r * XSc + j];
          ab[j] += s;
        }
      }
      // (b) dprev = dz . W_l^T
      {
        const int inpad = ceil_to(in_l, 16);
        if (cg < inpad / GNNFP_JC)
          dense_tile(cur, XSc, oth, XSo, smem + y.oWT[l], zero, H, inpad, GNNFP_ACT_LINEAR, rg, cg, tc.CG, lane);
      }
      __syncthreads();
      if (l > 0) {
        const int actp = net.acts[l - 1];
        const unsigned magic = (unsigned)((0x100000000ull + (unsigned)in_l - 1) / (unsigned)in_l);
        for (int e = tid; e < tc.R * in_l; e += T) {
          const int r = (int)__umulhi((unsigned)e, magic);
          const int c = e - r * in_l;
          oth[r * XSo + c] = act_bwd(actp, al[r * XSl + c], oth[r * XSo + c]);
        }
        __syncthreads();
      }
      float* t2 = cur; cur = oth; oth = t2;
      const int t3 = XSc; XSc = XSo; XSo = t3;
    }
    // ---- cur = dy (gradient w.r.t. the BN output / the raw input) ---------------------------------
    if (y.bn) {
      const float* a0 = smem + y.oAct[0];
      const int pin = ceil_to(net.in_dim, 4);
      for (int c = tid; c < net.in_dim; c += T) {
        float p = 0.f, q = 0.f;
        for (int r = 0; r < nr; ++r) {
          const float dy = cur[r * XSc + c];
          p += dy;
          q = fmaf(dy, a0[r * y.XSa[0] + c], q);
        }
        accBN[c] += p;
        accBN[pin + c] += q;
      }
    }
    for (int p = 0; p < a.src.n_pieces; ++p) {
      const Piece& pc = a.src.p[p];
      if (pc.gmode == GM_NONE) continue;
      const int w = pc.width;
      for (int e = tid; e < nr * w; e += T) {
        const int r = (int)__umulhi((unsigned)e, pc.magic);
        const int c = e - r * w;
        float v = cur[r * XSc + pc.col0 + c];
        if (y.bn) v *= bnS[pc.col0 + c];
        const int gr = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
        const int drow = pc.map ? pc.map[gr] : gr;
        float* d = pc.gptr + (size_t)drow * pc.gld + c;
        if (pc.gmode == GM_STORE) *d = v;
        else if (pc.gmode == GM_ADD) *d += v;
        else atomicAdd(d, v);
      }
    }
    __syncthreads();
  }
  // ---- flush per-CTA accumulators to this CTA's partial slot (plain +=: the slot is private) ----
  {
    float* part = a.partial + (size_t)blockIdx.x * a.n_params;
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
      const int in_l = y.inw[l], H = y.inw[l + 1];
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
              if (c0 + i < in_l && j0 + jj < H) part[off + (size_t)(c0 + i) * H + j0 + jj] += (q == 0 ? racc[0][i][jj] : racc[1][i][jj]);
        }
      } else {
        const float* aw = smem + y.oAccW[l];
        for (int e = tid; e < in_l * H; e += T) {
          float s = 0.f;
          for (int g = 0; g < y.groups[l]; ++g) s += aw[(size_t)g * in_l * H + e];
          part[off + e] += s;
        }
      }
      off += (size_t)in_l * H;
      const float* ab = smem + y.oAccb[l];
      for (int j = tid; j < H; j += T) part[off + j] += ab[j];
      off += H;
    }
    if (y.bn && a.bn_partial) {
      const int pin = ceil_to(net.in_dim, 4);
      float* bp = a.bn_partial + (size_t)blockIdx.x * 2 * net.in_dim;
      for (int c = tid; c < net.in_dim; c += T) {
        bp[c] = accBN[c];
        bp[net.in_dim + c] = accBN[pin + c];
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
  float* bn_const;    // [2*in_dim] c0 then c1
  const int* gate;
};
__global__ void bn_reduce_kernel(const __grid_constant__ BnReduceArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ float sm[];
  const int in = a.net.in_dim;
  float* A = sm;
  float* Bc = sm + in;
  bn_coefficients(a.src, a.net, 0, A, Bc, nullptr, nullptr);
  __syncthreads();
  for (int c = threadIdx.x; c < in; c += blockDim.x) {
    double p = 0.0, q = 0.0;
    for (int b = 0; b < a.grid; ++b) {
      p += (double)a.bn_partial[(size_t)b * 2 * in + c];
      q += (double)a.bn_partial[(size_t)b * 2 * in + in + c];
    }
    a.bn_grad[c] += (float)q;
    a.bn_grad[in + c] += (float)p;
    const float ac = a.net.gamma[c] * A[c];
    if (a.net.bn_mode == 1) {
      a.bn_const[c] = (float)((double)ac * p * a.net.inv_n);
      a.bn_const[in + c] = (float)((double)ac * q * a.net.inv_n);
    } else {
      a.bn_const[c] = 0.f;
      a.bn_const[in + c] = 0.f;
    }
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
  float* X = smem + 4 * pin;
  const int tid = threadIdx.x, T = blockDim.x;
  bn_coefficients(a.src, a.net, 0, bnA, bnB, nullptr, nullptr);
  for (int c = tid; c < in; c += T) { c0[c] = a.bn_const[c]; c1[c] = a.bn_const[in + c]; }
  __syncthreads();
  const int n = a.src.n_rows;
  const int n_tiles = (n + tc.R - 1) / tc.R;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * tc.R;
    const int nr = min(tc.R, n - row0);
    stage_tile(a.src, row0, nr, tc.R, X, tc.XS0, bnA, bnB);
    __syncthreads();
    for (int p = 0; p < a.src.n_pieces; ++p) {
      const Piece& pc = a.src.p[p];
      if (pc.gmode == GM_NONE) continue;
      const int w = pc.width;
      for (int e = tid; e < nr * w; e += T) {
        const int r = (int)__umulhi((unsigned)e, pc.magic);
        const int c = e - r * w;
        const int cc = pc.col0 + c;
        const float corr = c0[cc] + X[r * tc.XS0 + cc] * c1[cc];
        const int gr = a.src.rowlist ? a.src.rowlist[row0 + r] : row0 + r;
        const int drow = pc.map ? pc.map[gr] : gr;
        float* d = pc.gptr + (size_t)drow * pc.gld + c;
        if (pc.gmode == GM_ATOMIC) atomicAdd(d, -corr);
        else *d -= corr;
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
        float v = (float)s;
        if (l == 0 && bn) {
          double sb = 0.0;
          const int ib = off + in_l * H + j;
          for (int b = 0; b < a.grid; ++b) sb += (double)a.partial[(size_t)b * a.n_params + ib];
          v = (float)((double)a.net.gamma[c] * s + (double)a.net.beta[c] * sb);
        }
        a.dW[l][e] = v * scale;
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
int tile_cfg_bwd(const NetDev& net, int n_rows, int gwidth, TileCfg* tc) {
  (void)gwidth;
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
  int regacc = 0;
  for (;;) {
    bwd_layout(net, 64 * RG, 256, 0, y);
    const bool too_big = y.total * 4 > want;
    const bool underfill = (n_rows + 64 * RG - 1) / (64 * RG) < 2 * nsm;
    if (RG > 1 && (too_big || underfill)) { RG /= 2; CG = 8 / RG; continue; }
    break;
  }
  if (y.total * 4 > want && net.n_layers == 1) {
    // one big Dense layer: keep the dW accumulators in registers (2 units of 8x4 per thread)
    const int U = ((net.in_dim + 7) / 8) * ((net.widths[0] + 3) / 4);
    if (U >= 256 && U <= 512) { regacc = 1; bwd_layout(net, 64 * RG, 256, 1, y); }
  }
  tc->RG = RG; tc->CG = CG; tc->R = 64 * RG; tc->threads = 256;
  tc->XS0 = y.XSa[0]; tc->XS1 = regacc;   // XS1 doubles as the register-accumulation switch of the backward kernel
  tc->smem_bytes = y.total * 4;
  if (tc->smem_bytes > cap)
    GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "net too large for the shared-memory backward tile kernel (%zu bytes needed)", tc->smem_bytes);
  int per_sm = (int)((220 * 1024) / (tc->smem_bytes + 1024));
  if (per_sm > 8) per_sm = 8;
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
int launch_bn_tail(const BwdArgs& a, float* bn_grad, float* bn_const, cudaStream_t s) {
  if (a.src.n_rows <= 0) return GNNFP_OK;
  BnReduceArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.src = a.src; ra.net = a.net; ra.bn_partial = a.bn_partial; ra.grid = a.tc.grid;
  ra.bn_grad = bn_grad; ra.bn_const = bn_const; ra.gate = a.gate;
  bn_reduce_kernel<<<1, 256, 2 * a.net.in_dim * sizeof(float), s>>>(ra);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  if (a.net.bn_mode != 1) return GNNFP_OK;
  bool any = false;
  for (int p = 0; p < a.src.n_pieces; ++p) any = any || a.src.p[p].gmode != GM_NONE;
  if (!any) return GNNFP_OK;
  BnFixArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.src = a.src; fa.net = a.net; fa.bn_const = bn_const; fa.gate = a.gate;
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
