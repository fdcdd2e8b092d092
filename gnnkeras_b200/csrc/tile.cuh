// tile.cuh - device helpers shared by the tile kernels: staging of input pieces into a
// shared-memory tile, activations and their derivatives, BN coefficient set-up.
//
// Tile layout: row-major [R][XS] floats with XS % 8 == 4, so that
//   * lane==row LDS.128 / STS.128 of 4 consecutive columns are bank-conflict free
//     (8 lanes per phase hit 8 distinct 16-byte bank groups because XS/4 is odd),
//   * staging writes (consecutive columns of one row) are conflict free.
// Staging never applies BatchNormalization: the forward folds the BN affine into the first Dense
// layer once per CTA, the backward works on raw inputs and corrects its reductions at flush time.
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

__host__ __device__ __forceinline__ int ceil_to(int x, int m) { return (x + m - 1) / m * m; }
// smallest stride >= w with stride % 8 == 4
__host__ __device__ __forceinline__ int tile_stride(int w) {
  const int p = ceil_to(w < 1 ? 1 : w, 4);
  return (p % 8 == 4) ? p : p + 4;
}

__device__ __forceinline__ bool piece_enabled(const Piece& pc) {
  if (pc.gate == nullptr) return true;
  return ((*pc.gate) != 0) == (pc.gate_pol != 0);
}

// shared-memory scratch for the CSR slice of one tile (identity row sets only)
struct StageScratch {
  int* rp;      // [R + 1]
  int* sidx;    // [cap]
  float* sw;    // [cap]
  int cap;
};
__host__ __device__ __forceinline__ size_t scratch_floats(int R, int cap) { return (size_t)ceil_to(R + 1, 4) + 2 * (size_t)ceil_to(cap, 4); }
__device__ __forceinline__ StageScratch carve_scratch(float* base, int R, int cap) {
  StageScratch s;
  s.rp = reinterpret_cast<int*>(base);
  s.sidx = s.rp + ceil_to(R + 1, 4);
  s.sw = reinterpret_cast<float*>(s.sidx + ceil_to(cap, 4));
  s.cap = cap;
  return s;
}

// ---- gather with the tile's CSR slice resident in shared memory ---------------------------------
// one item = (row r, VEC-wide column chunk q): all source addresses are known from shared memory, so
// the (up to 4) neighbour-row loads of an item are issued back to back; accumulation stays sequential
// in arc order (the order TF-CPU SparseTensorDenseMatMul uses).
template <int VEC>
__device__ __forceinline__ void load_vec(const float* sp, float* out) {
  if (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(sp);
    out[0] = t.x; out[1 % VEC] = t.y; out[2 % VEC] = t.z; out[3 % VEC] = t.w;
  } else if (VEC == 2) {
    const float2 t = *reinterpret_cast<const float2*>(sp);
    out[0] = t.x; out[1 % VEC] = t.y;
  } else {
    out[0] = *sp;
  }
}

// NI items (row, VEC-wide chunk) per thread are processed together: the first two arcs of every item are
// loaded before anything is consumed (2*NI independent loads in flight), longer rows continue sequentially.
template <int VEC, int NI>
__device__ __forceinline__ void gather_items(const Piece& pc, const StageScratch& sc, int abase, int nr, int R,
                                             float* X, int XS, bool has_w) {
  const int w = pc.width, nq = w / VEC;
  const int items = R * nq;
  const int T = blockDim.x;
  for (int it0 = threadIdx.x; it0 < items; it0 += NI * T) {
    float acc[NI][VEC];
    float v0[NI][VEC], v1[NI][VEC], w0[NI], w1[NI];
    int a0[NI], a1[NI], q[NI], rr[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int it = it0 + i * T;
      rr[i] = -1; a0[i] = 0; a1[i] = 0; q[i] = 0; w0[i] = 0.f; w1[i] = 0.f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { acc[i][v] = 0.f; v0[i][v] = 0.f; v1[i][v] = 0.f; }
      if (it < items) {
        const int r = it / nq;
        rr[i] = r; q[i] = it - r * nq;
        if (r < nr) { a0[i] = sc.rp[r] - abase; a1[i] = sc.rp[r + 1] - abase; }
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (a0[i] < a1[i]) {
        load_vec<VEC>(pc.ptr + (size_t)sc.sidx[a0[i]] * pc.ld + q[i] * VEC, v0[i]);
        w0[i] = has_w ? sc.sw[a0[i]] : 1.0f;
      }
      if (a0[i] + 1 < a1[i]) {
        load_vec<VEC>(pc.ptr + (size_t)sc.sidx[a0[i] + 1] * pc.ld + q[i] * VEC, v1[i]);
        w1[i] = has_w ? sc.sw[a0[i] + 1] : 1.0f;
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      if (a0[i] < a1[i]) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[i][v] = fmaf(w0[i], v0[i][v], acc[i][v]);
      }
      if (a0[i] + 1 < a1[i]) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[i][v] = fmaf(w1[i], v1[i][v], acc[i][v]);
      }
      for (int a = a0[i] + 2; a < a1[i]; ++a) {     // arcs beyond the second: sequential, still in arc order
        float t[VEC];
        load_vec<VEC>(pc.ptr + (size_t)sc.sidx[a] * pc.ld + q[i] * VEC, t);
        const float wv = has_w ? sc.sw[a] : 1.0f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[i][v] = fmaf(wv, t[v], acc[i][v]);
      }
      if (rr[i] >= 0) {
        float* d = X + rr[i] * XS + pc.col0 + q[i] * VEC;
        if (pc.accumulate) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[v] += acc[i][v];
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) d[v] = acc[i][v];
        }
      }
    }
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2-prefetch the byte range [p, p+bytes) cooperatively (one 128-byte line per thread-iteration)
__device__ __forceinline__ void prefetch_range(const void* p, size_t bytes) {
  const char* base = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)127);
  const size_t lines = (bytes + (reinterpret_cast<uintptr_t>(p) & 127) + 127) / 128;
  for (size_t i = threadIdx.x; i < lines; i += blockDim.x) prefetch_l2(base + i * 128);
}

// Stage every piece of `ts` for tile rows [row0, row0+nr) into X[r*XS + col] (raw values).  Rows
// nr..R-1 are zero-filled.  Ends WITHOUT a trailing __syncthreads (callers sync).
// CTAs walk CONSECUTIVE tiles, so while staging tile i the data of tile i+1 (the next R rows and the next
// slice of the CSR) is prefetched into L2: its DRAM latency is then hidden behind tile i's compute.
__device__ __forceinline__ void stage_tile(const TileSrc& ts, int row0, int nr, int R, float* X, int XS,
                                           const StageScratch& sc, bool prefetch_next = true, unsigned skip_mask = 0u) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int next0 = row0 + R;
  const int next_nr = prefetch_next ? max(0, min(R, ts.n_rows - next0)) : 0;
  for (int p = 0; p < ts.n_pieces; ++p) {
    const Piece& pc = ts.p[p];
    if (skip_mask & (1u << p)) continue;      // staged by the asynchronous bulk path
    const bool on = piece_enabled(pc);
    if (pc.accumulate) {
      __syncthreads();          // the piece it adds to may have been staged with another mapping
      if (!on) continue;
    }
    const int w = pc.width;
    const int total = R * w;
    const int c0 = pc.col0;
    if (pc.kind == PK_DIRECT) {
      const bool ident = ts.rowlist == nullptr && pc.map == nullptr;
      const bool flat = on && ident && pc.rowscale == nullptr && pc.ld == w &&
                        ((reinterpret_cast<uintptr_t>(pc.ptr + (size_t)row0 * w) & 15) == 0);
      if (on && ident && next_nr > 0) prefetch_range(pc.ptr + (size_t)next0 * pc.ld, (size_t)next_nr * pc.ld * sizeof(float));
      if (flat) {
        // the tile's rows are one contiguous, 16-byte aligned run: 128-bit loads, 4 in flight per thread
        const float* srcf = pc.ptr + (size_t)row0 * w;
        const float4* src4 = reinterpret_cast<const float4*>(srcf);
        const int nvalid = nr * w;
        const int n4 = (total + 3) / 4;
        for (int i0 = tid; i0 < n4; i0 += 4 * T) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * T;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n4) {
              if (4 * i + 3 < nvalid) v[u] = src4[i];
              else {
                if (4 * i < nvalid) v[u].x = srcf[4 * i];
                if (4 * i + 1 < nvalid) v[u].y = srcf[4 * i + 1];
                if (4 * i + 2 < nvalid) v[u].z = srcf[4 * i + 2];
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * T;
            if (i < n4) {
              const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
              int r = div_magic((unsigned)(4 * i), pc.magic);
              int c = 4 * i - r * w;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (4 * i + k < total) {
                  float* d = X + r * XS + c0 + c;
                  if (pc.accumulate) *d += vv[k]; else *d = vv[k];
                }
                if (++c == w) { c = 0; ++r; }
              }
            }
          }
        }
      } else {
        // 4 independent loads in flight per thread, then the 4 stores
        for (int e0 = tid; e0 < total; e0 += 4 * T) {
          float v[4];
          int off[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * T;
            v[u] = 0.f;
            off[u] = -1;
            if (e < total) {
              const int r = div_magic((unsigned)e, pc.magic);
              const int c = e - r * w;
              off[u] = r * XS + c0 + c;
              if (on && r < nr) {
                const int gr = ts.rowlist ? ts.rowlist[row0 + r] : row0 + r;
                const int sr = pc.compact ? (row0 + r) : (pc.map ? pc.map[gr] : gr);
                v[u] = pc.ptr[(size_t)sr * pc.ld + c];
                if (pc.rowscale) v[u] *= pc.rowscale[gr];
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (off[u] >= 0) { if (pc.accumulate) X[off[u]] += v[u]; else X[off[u]] = v[u]; }
        }
      }
    } else {
      // ---- gather: fast path = identity rows and the tile's arcs fit the scratch ----------------
      bool fast = on && ts.rowlist == nullptr && sc.cap > 0;
      int abase = 0;
      if (fast) {
        __syncthreads();                               // scratch may still be read by a previous piece
        for (int i = tid; i <= nr; i += T) sc.rp[i] = pc.rowptr[row0 + i];
        if (next_nr > 0) prefetch_range(pc.rowptr + next0, (size_t)(next_nr + 1) * sizeof(int));
        __syncthreads();
        abase = sc.rp[0];
        const int na = sc.rp[nr] - abase;
        fast = na <= sc.cap;                           // uniform across the CTA
        if (next_nr > 0 && pc.nnz > 0) {               // next tile's CSR slice follows this one
          const int nb = abase + na;
          const int len = min(na + 32, pc.nnz - nb);
          if (len > 0) {
            prefetch_range(pc.idx + nb, (size_t)len * sizeof(int));
            if (pc.wgt) prefetch_range(pc.wgt + nb, (size_t)len * sizeof(float));
          }
        }
        if (fast) {
          for (int i = tid; i < na; i += T) {
            sc.sidx[i] = pc.idx[abase + i];
            if (pc.wgt) sc.sw[i] = pc.wgt[abase + i];
          }
          __syncthreads();
        }
      }
      if (fast) {
        const bool al16 = ((reinterpret_cast<uintptr_t>(pc.ptr) & 15) == 0) && (pc.ld % 4 == 0) && (w % 4 == 0) && (c0 % 4 == 0);
        const bool al8 = ((reinterpret_cast<uintptr_t>(pc.ptr) & 7) == 0) && (pc.ld % 2 == 0) && (w % 2 == 0) && (c0 % 2 == 0);
        if (al16) gather_items<4, 4>(pc, sc, abase, nr, R, X, XS, pc.wgt != nullptr);
        else if (al8) gather_items<2, 4>(pc, sc, abase, nr, R, X, XS, pc.wgt != nullptr);
        else gather_items<1, 4>(pc, sc, abase, nr, R, X, XS, pc.wgt != nullptr);
      } else {
        for (int e = tid; e < total; e += T) {
          const int r = div_magic((unsigned)e, pc.magic);
          const int c = e - r * w;
          float v = 0.0f;
          if (on && r < nr) {
            const int gr = ts.rowlist ? ts.rowlist[row0 + r] : row0 + r;
            const int a0 = pc.rowptr[gr], a1 = pc.rowptr[gr + 1];
            for (int a = a0; a < a1; ++a) {
              const float wv = pc.wgt ? pc.wgt[a] : 1.0f;
              v = fmaf(wv, pc.ptr[(size_t)pc.idx[a] * pc.ld + c], v);
            }
          }
          float* d = X + r * XS + c0 + c;
          if (pc.accumulate) *d += v; else *d = v;
        }
      }
    }
  }
}

// ---- asynchronous staging: TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier -----------------------
// Direct pieces whose tile rows form one contiguous 16-byte aligned run are fetched by ONE bulk copy per piece
// into a dense landing buffer while the previous tile is being computed; after the mbarrier flips, all threads
// re-lay the landing buffer into the padded compute tile (shared -> shared, no global latency).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE%=;\n\tbra LAB_WAIT%=;\n\tDONE%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// thread 0: arm the barrier and issue one bulk copy per selected piece of `ts` (tile rows row0..row0+R, all valid)
__device__ __forceinline__ uint32_t bulk_bytes(const TileSrc& ts, unsigned mask, int R) {
  uint32_t b = 0;
  for (int p = 0; p < ts.n_pieces; ++p)
    if (mask & (1u << p)) b += (uint32_t)R * ts.p[p].width * 4u;
  return b;
}
__device__ __forceinline__ float* bulk_issue(const TileSrc& ts, unsigned mask, int row0, int R, float* raw, uint64_t* bar) {
  for (int p = 0; p < ts.n_pieces; ++p)
    if (mask & (1u << p)) {
      const Piece& pc = ts.p[p];
      bulk_g2s(raw, pc.ptr + (size_t)row0 * pc.width, (uint32_t)R * pc.width * 4u, bar);
      raw += R * pc.width;
    }
  return raw;
}
// all threads: landing buffer -> compute tile
__device__ __forceinline__ const float* bulk_relayout(const TileSrc& ts, unsigned mask, int R, const float* raw, float* X, int XS) {
  const int tid = threadIdx.x, T = blockDim.x;
  for (int p = 0; p < ts.n_pieces; ++p)
    if (mask & (1u << p)) {
      const Piece& pc = ts.p[p];
      const int w = pc.width, n4 = R * w / 4, c0 = pc.col0;
      const float4* r4 = reinterpret_cast<const float4*>(raw);
#pragma unroll 2
      for (int i = tid; i < n4; i += T) {
        const float4 v = r4[i];
        const float vv[4] = {v.x, v.y, v.z, v.w};
        int r = div_magic((unsigned)(4 * i), pc.magic);
        int c = 4 * i - r * w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float* d = X + r * XS + c0 + c;
          if (pc.accumulate) *d += vv[k]; else *d = vv[k];
          if (++c == w) { c = 0; ++r; }
        }
      }
      raw += R * w;
    }
  return raw;
}

// Per-column BN coefficients: x_hat = x*a + b.
//   affine=1: a = gamma*rsqrt(var+eps), b = beta - mean*a      (what the forward applies)
//   affine=0: a = rsqrt(var+eps),       b = -mean*a            (x_tilde, for the backward)
// bn_mode 1 = batch stats from the pieces' st_sum/st_sq (double sums), 2 = moving statistics.
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

// ---- one Dense layer on the tile -----------------------------------------------------------------
// warp (rg, cg) owns rows [rg*64, rg*64+64) (lane -> rows rg*64+lane and +32) and 16-column chunks
// ch = cg, cg+CG, ...  Inputs are read as LDS.128 (4 columns at a time), weights as 128-bit broadcasts
// from Wl[in_pad4][Hpad]; outputs are written as STS.128.  in4 = ceil(in/4) (padding columns of the tile
// and padding rows of Wl are zero).
__device__ __forceinline__ void dense_tile(const float* __restrict__ Ain, int XSin, float* __restrict__ Aout, int XSout,
                                           const float* __restrict__ Wl, const float* __restrict__ bl, int in4, int Hpad,
                                           int act, int rg, int cg, int CG, int lane) {
  const int nch = Hpad / GNNFP_JC;
  const float4* x0p = reinterpret_cast<const float4*>(Ain + (rg * 64 + lane) * XSin);
  const float4* x1p = reinterpret_cast<const float4*>(Ain + (rg * 64 + lane + 32) * XSin);
  const int wstride = Hpad / 4;
  for (int ch = cg; ch < nch; ch += CG) {
    float acc0[GNNFP_JC], acc1[GNNFP_JC];
#pragma unroll
    for (int j = 0; j < GNNFP_JC; ++j) { const float bj = bl[ch * GNNFP_JC + j]; acc0[j] = bj; acc1[j] = bj; }
    const float4* wp = reinterpret_cast<const float4*>(Wl + ch * GNNFP_JC);
#pragma unroll 1
    for (int c4 = 0; c4 < in4; ++c4) {
      const float4 xa = x0p[c4], xb = x1p[c4];
      const float xs0[4] = {xa.x, xa.y, xa.z, xa.w};
      const float xs1[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        wp += wstride;
        const float x0 = xs0[k], x1 = xs1[k];
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
    }
    float4* o0 = reinterpret_cast<float4*>(Aout + (rg * 64 + lane) * XSout + ch * GNNFP_JC);
    float4* o1 = reinterpret_cast<float4*>(Aout + (rg * 64 + lane + 32) * XSout + ch * GNNFP_JC);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      o0[j4] = make_float4(act_fwd(act, acc0[4 * j4]), act_fwd(act, acc0[4 * j4 + 1]), act_fwd(act, acc0[4 * j4 + 2]), act_fwd(act, acc0[4 * j4 + 3]));
      o1[j4] = make_float4(act_fwd(act, acc1[4 * j4]), act_fwd(act, acc1[4 * j4 + 1]), act_fwd(act, acc1[4 * j4 + 2]), act_fwd(act, acc1[4 * j4 + 3]));
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

// column statistics of a tile (first H columns, nr rows) accumulated into shared double accumulators
// acc[0..H) (sum) and acc[H..2H) (sum of squares) by ALL threads: thread -> (column, row group).
__device__ __forceinline__ void tile_col_stats(const float* A, int XS, int H, int nr, double* acc) {
  const int T = blockDim.x;
  const int ng = T / H > 0 ? T / H : 1;              // row groups
  const int tid = threadIdx.x;
  if (T >= H) {
    const int j = tid % H, g = tid / H;
    if (g < ng) {
      double su = 0.0, sq = 0.0;
      for (int r = g; r < nr; r += ng) {
        const double v = (double)A[r * XS + j];
        su += v;
        sq += v * v;
      }
      atomicAdd(acc + j, su);
      atomicAdd(acc + H + j, sq);
    }
  } else {
    for (int j = tid; j < H; j += T) {
      double su = 0.0, sq = 0.0;
      for (int r = 0; r < nr; ++r) {
        const double v = (double)A[r * XS + j];
        su += v;
        sq += v * v;
      }
      acc[j] += su;
      acc[H + j] += sq;
    }
  }
}
