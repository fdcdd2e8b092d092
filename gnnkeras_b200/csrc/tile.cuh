// tile.cuh - device helpers shared by the tile kernels: staging of input pieces into a
// shared-memory tile (row-major, odd row stride => conflict-free for lane==row access),
// activations and their derivatives, BN coefficient set-up.
#pragma once
#include "common.cuh"

#define SELU_SCALE_F 1.0507009873554805f
#define SELU_ALPHA_F 1.6732632423543772f

__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case GNNFP_ACT_TANH: return tanhf(z);
    case GNNFP_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    case GNNFP_ACT_RELU: return fmaxf(z, 0.0f);
    case GNNFP_ACT_SELU:  // TF functor: (x<0) ? scale_alpha*(exp(x)-1) : scale*x
      return z < 0.0f ? (SELU_SCALE_F * SELU_ALPHA_F) * (expf(z) - 1.0f) : SELU_SCALE_F * z;
    default: return z;   // linear; softmax is applied row-wise afterwards
  }
}

// d act / d z expressed through the OUTPUT y (all supported activations allow it; TF's
// TanhGrad/SigmoidGrad/ReluGrad/SeluGrad use the output too).
__device__ __forceinline__ float act_bwd(int act, float y, float g) {
  switch (act) {
    case GNNFP_ACT_TANH: return g * (1.0f - y * y);
    case GNNFP_ACT_SIGMOID: return g * y * (1.0f - y);
    case GNNFP_ACT_RELU: return y > 0.0f ? g : 0.0f;
    case GNNFP_ACT_SELU: return y < 0.0f ? g * (y + SELU_SCALE_F * SELU_ALPHA_F) : g * SELU_SCALE_F;
    default: return g;
  }
}

__device__ __forceinline__ bool piece_enabled(const Piece& pc) {
  if (pc.gate == nullptr) return true;
  return ((*pc.gate) != 0) == (pc.gate_pol != 0);
}

// Stage every piece of `ts` for tile rows [row0, row0+nr) into X[r*XS + col].  Rows nr..R-1 are
// zero-filled.  bnA/bnB (nullable): per-column affine applied while staging (BN as x*a+b,
// tf.nn.batch_normalization's own form).
__device__ __forceinline__ void stage_tile(const TileSrc& ts, int row0, int nr, int R, float* X, int XS,
                                           const float* bnA, const float* bnB) {
  for (int p = 0; p < ts.n_pieces; ++p) {
    const Piece& pc = ts.p[p];
    const bool on = piece_enabled(pc);
    if (pc.accumulate) {
      __syncthreads();          // the piece it adds to may have been staged with another mapping
      if (!on) continue;
    }
    const int w = pc.width;
    const int total = R * w;
    const int c0 = pc.col0;
    if (pc.kind == PK_DIRECT) {
#pragma unroll 4
      for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int r = (int)__umulhi((unsigned)e, pc.magic);
        const int c = e - r * w;
        float v = 0.0f;
        if (on && r < nr) {
          const int gr = ts.rowlist ? ts.rowlist[row0 + r] : row0 + r;
          const int sr = pc.compact ? (row0 + r) : (pc.map ? pc.map[gr] : gr);
          v = pc.ptr[(size_t)sr * pc.ld + c];
          if (pc.rowscale) v *= pc.rowscale[gr];
          if (bnA) v = fmaf(v, bnA[c0 + c], bnB[c0 + c]);
        }
        float* d = X + r * XS + c0 + c;
        if (pc.accumulate) *d += v; else *d = v;
      }
    } else {
      for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int r = (int)__umulhi((unsigned)e, pc.magic);
        const int c = e - r * w;
        float v = 0.0f;
        if (on && r < nr) {
          const int gr = ts.rowlist ? ts.rowlist[row0 + r] : row0 + r;
          const int a0 = pc.rowptr[gr], a1 = pc.rowptr[gr + 1];
          // sequential in arc order: the order TF-CPU SparseTensorDenseMatMul accumulates in
          for (int a = a0; a < a1; ++a) {
            const float wv = pc.wgt ? pc.wgt[a] : 1.0f;
            v = fmaf(wv, pc.ptr[(size_t)pc.idx[a] * pc.ld + c], v);
          }
          if (bnA) v = fmaf(v, bnA[c0 + c], bnB[c0 + c]);
        }
        float* d = X + r * XS + c0 + c;
        if (pc.accumulate) *d += v; else *d = v;
      }
    }
  }
}

// Per-column BN coefficients: x_hat = x*a + b.
//   affine=1: a = gamma*rsqrt(var+eps), b = beta - mean*a      (what the forward applies)
//   affine=0: a = rsqrt(var+eps),       b = -mean*a            (x_tilde, for the backward)
// Also returns mean/var through optional arrays.  bn_mode 1 = batch stats from the pieces'
// st_sum/st_sq (double sums), 2 = moving statistics.
__device__ __forceinline__ void bn_coefficients(const TileSrc& ts, const NetDev& net, int affine, float* A,
                                                float* B, float* meanOut, float* varOut) {
  for (int cc = threadIdx.x; cc < net.in_dim; cc += blockDim.x) {
    float mean = 0.f, var = 1.f;
    if (net.bn_mode == 1) {
      for (int p = 0; p < ts.n_pieces; ++p) {
        const Piece& pc = ts.p[p];
        if (pc.accumulate) continue;
        if (cc >= pc.col0 && cc < pc.col0 + pc.width && pc.st_sum) {
          const double m = pc.st_sum[cc - pc.col0] * net.inv_n;
          double v = pc.st_sq[cc - pc.col0] * net.inv_n - m * m;
          if (v < 0.0) v = 0.0;
          mean = (float)m;
          var = (float)v;
        }
      }
    } else {
      mean = net.mmean[cc];
      var = net.mvar[cc];
    }
    const float rs = 1.0f / sqrtf(var + net.bn_eps);
    const float a = affine ? rs * net.gamma[cc] : rs;
    A[cc] = a;
    B[cc] = affine ? (net.beta[cc] - mean * a) : (-mean * a);
    if (meanOut) meanOut[cc] = mean;
    if (varOut) varOut[cc] = var;
  }
}

__host__ __device__ __forceinline__ int ceil_to(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ __forceinline__ int odd_stride(int w) { return (w | 1); }
