// rows_tma.cu - TMA-fed row GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, FP32
// accumulators in TMEM).  One kernel, two uses:
//
//   RT_FWD  one fixed-point iteration of a single-Dense-layer state net (reference GNN.py:217-236 `convergence` +
//           the `condition` test of GNN.py:196-214):   S_t = act(BN([S_{t-1} | nodes? | Adj^T S_{t-1} | static]) W + b)
//           - BatchNormalization is folded into the shared-memory weights by every CTA from the fp64 batch sums the
//             previous launch left (no separate fold / coefficient launches),
//           - epilogue: bias + activation, per-row convergence test against the previous state, column statistics of
//             S_t for the next iteration's BatchNormalization, all from the output stage in shared memory.
//   RT_DX   backward of the same layer w.r.t. its input blocks:   [dOwn_t | dAgg_t] = dz (W^T . gamma rstd) - BN correction
//           (both blocks from ONE pass over dz: the MMA N covers both weight blocks).
//
// Data path per persistent CTA (one per SM, 15 warps):
//   warp 12 (one thread)  TMA producer: cp.async.bulk.tensor 2-D boxes of [128 rows x 32 columns] fp32 straight into the
//                         K-major SWIZZLE_128B operand tile "hi" of a ring stage (ragged rows / columns are zero-filled by
//                         the TMA unit: no predicates anywhere).  The tensor core reads the top 19 bits of an fp32 word,
//                         so the raw tile IS a_hi = trunc_tf32(a).
//   warps 4-7             converters, thread = row: a_lo = rn_tf32(a - a_hi) into the stage's "lo" tile (same swizzled
//                         position: conflict-free 128-bit accesses), fence.proxy.async, mbarrier arrive.
//   warp 13 (one thread)  MMA issuer: 3 tcgen05.mma (lo.hi, hi.lo, hi.hi) per 8-wide K step against the resident
//                         W_hi / W_lo tiles; tcgen05.commit frees the stage / publishes the accumulator (two TMEM buffers).
//   warps 0-3             epilogue, thread = row: tcgen05.ld 32 columns, activation / correction against the side input that
//                         the out thread preloaded INTO the output stage by TMA (previous state / BN input x), results
//                         written back in place (swizzled 128-bit stores).
//   warps 8-11            column warps (RT_FWD), lane = column: column sums / sums of squares of the finished output stage.
//   warp 14 (one thread)  out thread: TMA preload of the side input, TMA store of the finished stage (rows / columns outside
//                         the matrix are clipped by the TMA unit).
// SASS: UTMALDG / UTMASTG (tensor TMA), UTCHMMA, LDTM, UTCBAR.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "rows_tma.h"
#include "tc.cuh"

#define RT_THREADS 512
#define RT_W_EPI 0
#define RT_W_CONV 4
#define RT_W_COL 8
#define RT_W_PROD 12
#define RT_W_MMA 13
#define RT_W_OUT 14
#define RT_W_CSR 15
#define RT_CSR_RP_BYTES 528                                   // 132 row pointers
#define RT_CSR_LI_BYTES ((GNNFP_TILE_ARCS + 8) * 2)           // local source indices (16-byte aligned window)
#define RT_CSR_W_BYTES ((GNNFP_TILE_ARCS + 8) * 4)            // weights
#define RT_CSR_BUF (544 + 2080 + RT_CSR_W_BYTES)              // one staged CSR slice
#define RT_STAT_SLOTS 3

__device__ __forceinline__ void rt_tma_load(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void rt_tma_store(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void rt_tmem_ld32(uint32_t addr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// phase time stamps of the last launch (CTA x [entry, constants, weights ready, main loop done, exit], %globaltimer ns):
// read back by gnnfp_debug_rt_times (scratch/profiling only; five 8-byte stores per CTA)
__device__ long long g_rt_times[2][160][8];
__device__ __forceinline__ long long rt_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define RT_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_rt_times[MODE][blockIdx.x][i] = rt_now(); } while (0)

// D[tmem] (+)= A[tmem] . B[smem desc]^T: the A operand (M = 128 lanes x 8 TF32 columns) comes from tensor memory
__device__ __forceinline__ void rt_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void rt_tmem_st8(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// one lane of a converged warp (the single-thread roles run their loops warp-uniformly and predicate only the
// instruction itself: under `if (lane == 0)` every operand is a per-thread value and each tcgen05.mma / TMA instruction is
// wrapped in a ~20-instruction R2UR "waterfall" loop - measured ~3 us per 128-row tile for the MMA issuer alone)
__device__ __forceinline__ bool rt_elect() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void rt_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(RT_THREADS, 1) rows_tma_kernel(const __grid_constant__ RowsTmaArgs a) {
  if (a.gate && *a.gate == 0) return;
  RT_STAMP(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int BN = a.BN, NKC = a.n_kc, NOC = a.n_oc, NST = a.n_stages, NOS = a.n_ostages;
  const int wtile = BN * 128;                          // bytes of one [BN x 32] weight tile
  uint8_t* Whi = base;                                 // [NKC][BN x 32] hi, then lo
  uint8_t* Wlo = Whi + (size_t)NKC * wtile;
  uint8_t* ring = Wlo + (size_t)NKC * wtile;           // [NST] raw fp32 operand tiles = a_hi (a_lo lives in tensor memory)
  uint8_t* outst = ring + (size_t)NST * RT_STAGE_BYTES;       // [NOS]
  uint8_t* csrb = outst + (size_t)NOS * RT_STAGE_BYTES;       // [2] CSR slices of the fused aggregation (fuse_agg only)
  const uint32_t lo_col0 = (uint32_t)(2 * BN + 32);    // TMEM columns [lo_col0 + 32 * slot, +32): a_lo of ring slot `slot`
  __shared__ __align__(8) uint64_t slot_empty[8], hi_full[8], ops_full[8], tm_full[2], tm_empty[2], aux_full[4], out_full[4], col_done[4], csr_full[2], csr_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int notconv_s;
  __shared__ __align__(16) float s_c[4][RT_MAXIN];     // FWD: BN a | b | mean | var per input column
  __shared__ __align__(16) float s_tab[RT_MAXOC][4][32];   // per output chunk: FWD bias;  DX c0 | c1 | A | B
  __shared__ float s_part[4][128];
  __shared__ short s_wrow[RT_MAXKC * RT_CHUNK];        // copy of the chunks' W-row tables: lane-varying indices into kernel
                                                       // parameters serialise in the constant cache (~1 us per access)
  __shared__ double s_stat[4][RT_STAT_SLOTS][4][32];    // per column warp: sum / sum of squares of S_t and of Adj^T S_t
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_rows;
  const int n_tiles = (n + RT_ROWS - 1) / RT_ROWS;
  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int my_tiles = max(0, min(n_tiles, tile0 + tiles_per_cta) - tile0);
  const bool fuse = MODE == RT_FWD && a.fuse_agg != 0;
  const bool use_col = MODE == RT_FWD && (a.ost_sum != nullptr || fuse);

  if (warp == RT_W_MMA) { tmem_alloc(&tmem_base_s, (uint32_t)a.tmem_cols); tmem_relinquish(); }
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&slot_empty[i], 1); mbar_init(&hi_full[i], 1); mbar_init(&ops_full[i], 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], 128); }
    for (int i = 0; i < NOS; ++i) { mbar_init(&aux_full[i], 1); mbar_init(&out_full[i], 128); mbar_init(&col_done[i], 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(&csr_full[i], 1); mbar_init(&csr_empty[i], 128); }
    notconv_s = 0;
  }
  for (int e = tid; e < NKC * RT_CHUNK; e += RT_THREADS) s_wrow[e] = a.kc[e >> 5].wrow[e & 31];
  // ---- per-column constants ------------------------------------------------------------------------------------------
  if (MODE == RT_FWD) {
    if (a.net.bn_mode) {
      const bool upd = a.update_moving && a.net.bn_mode == 1 && blockIdx.x == 0;
      bn_coefficients(a.src, a.net, 1, s_c[0], s_c[1], upd ? s_c[2] : nullptr, upd ? s_c[3] : nullptr);
      if (a.coef_out && blockIdx.x == 0 && a.net.bn_mode == 1) {   // the same function the backward's bn_coef_kernel ran
        const int in = a.net.in_dim;
        bn_coefficients(a.src, a.net, 0, a.coef_out, a.coef_out + in, nullptr, nullptr);
        for (int c = tid; c < in; c += RT_THREADS) a.coef_out[2 * in + c] = a.net.gamma[c] * a.coef_out[c];   // own element: no barrier needed
      }
    }
  } else {
    for (int e = tid; e < NOC * 32; e += RT_THREADS) {
      const int o = e >> 5, j = e & 31;
      const int c = a.oc[o].cidx0 + j;
      const bool ok = a.corr != nullptr && j < a.oc[o].width;
      s_tab[o][0][j] = ok ? a.corr[c] : 0.f;
      s_tab[o][1][j] = ok ? a.corr[a.corr_in + c] : 0.f;
      s_tab[o][2][j] = ok ? a.corr[2 * a.corr_in + c] : 0.f;
      s_tab[o][3][j] = ok ? a.corr[3 * a.corr_in + c] : 0.f;
    }
  }
  __syncthreads();
  RT_STAMP(1);
  // ---- resident weights: TF32 hi / lo, K-major SWIZZLE_128B tiles per K chunk -------------------------------------------
  const int H = a.H;
  if (MODE == RT_FWD) {
    const float* Wg = a.net.W[0];
    const bool bn = a.net.bn_mode != 0;
    if (bn && a.update_moving && a.net.bn_mode == 1 && blockIdx.x == 0) {   // Keras _assign_moving_average: v -= (v - value) * (1 - momentum)
      const float decay = (float)(1.0 - (double)a.net.bn_momentum);
      for (int c = tid; c < a.net.in_dim; c += RT_THREADS) {
        a.net.mmean[c] -= (a.net.mmean[c] - s_c[2][c]) * decay;
        a.net.mvar[c] -= (a.net.mvar[c] - s_c[3][c]) * decay;
      }
    }
    {
      // 8 independent global loads in flight per thread (a one-at-a-time loop costs a full L2 round trip per element:
      // measured 28 us for the widest layer)
      // branch-free batches: addresses first (invalid elements read W[0] and are scaled by 0), then all loads, then the math
      const int E = NKC * 32 * BN;
      for (int e0 = tid; e0 < E; e0 += 8 * RT_THREADS) {
        float w[8], sc[8];
        int off[8], src[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = min(e0 + u * RT_THREADS, E - 1);
          // e = (kc * 32 + k) * BN + nn: ONE division by BN, by multiplication (runtime integer divisions run on the
          // conversion pipe and were a third of this prologue)
          const int R = (int)__umulhi((unsigned)e, a.bn_magic);
          const int nn = e - R * BN, kc = R >> 5, k = R & 31;
          const int c = s_wrow[kc * RT_CHUNK + k];
          const bool ok = c >= 0 && nn < H;
          off[u] = (e0 + u * RT_THREADS < E) ? kc * wtile + tc_sw128_off(nn, k) : -1;
          src[u] = ok ? c * H + nn : 0;
          sc[u] = ok ? (bn ? s_c[0][c >= 0 ? c : 0] : 1.0f) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(Wg + src[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (off[u] >= 0) {
            const float v = w[u] * sc[u];
            *reinterpret_cast<uint32_t*>(Whi + off[u]) = __float_as_uint(v);
            *reinterpret_cast<uint32_t*>(Wlo + off[u]) = tc_lo(__float_as_uint(v));
          }
      }
    }
    // folded bias b_j + sum_c B_c W[c][j]: 4 fixed groups of input columns per output column, combined in a fixed order
    // (the launcher keeps BN <= 112 in this mode: 4 * BN <= RT_THREADS)
    if (tid < 4 * BN) {
      const int g = tid / BN, nn = tid - g * BN;
      float part = 0.f;
      if (bn && nn < H) {
        const int in = a.net.in_dim;
        int c = g;
        for (; c + 28 < in; c += 32) {                 // 8 loads in flight, summed in a fixed order
          float wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = __ldg(Wg + (size_t)(c + 4 * u) * H + nn);
#pragma unroll
          for (int u = 0; u < 8; ++u) part = fmaf(s_c[1][c + 4 * u], wv[u], part);
        }
        for (; c < in; c += 4) part = fmaf(s_c[1][c], __ldg(Wg + (size_t)c * H + nn), part);
      }
      s_part[g][nn] = part;
    }
    __syncthreads();
    for (int e = tid; e < NOC * 32; e += RT_THREADS) {
      const int o = e >> 5, j = e & 31;
      const int nn = a.oc[o].cidx0 + j;
      float b = 0.f;
      if (j < a.oc[o].width && nn < H) {
        b = a.net.b[0][nn];
        if (bn) b += (s_part[0][nn] + s_part[1][nn]) + (s_part[2][nn] + s_part[3][nn]);
      }
      s_tab[o][0][j] = b;
    }
  } else {
    // accumulator column nn -> input column (row of W) and its scale gamma * rstd; then the tiles with 8 loads in flight
    int* s_ncol = reinterpret_cast<int*>(ring);        // scratch in the (still unused) operand ring
    float* s_nscale = reinterpret_cast<float*>(ring) + 256;
    for (int nn = tid; nn < BN; nn += RT_THREADS) {
      int c = -1;
      for (int b = 0; b < a.n_blk; ++b)
        if (nn >= a.blk_acc0[b] && nn < a.blk_acc0[b] + a.blk_w[b]) c = a.blk_in0[b] + (nn - a.blk_acc0[b]);
      s_ncol[nn] = c;
      s_nscale[nn] = (c >= 0 && a.colscale) ? a.colscale[c] : 1.0f;
    }
    __syncthreads();
    {
      const int E = NKC * 32 * BN;
      for (int e0 = tid; e0 < E; e0 += 8 * RT_THREADS) {
        float w[8], sc[8];
        int off[8], src[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = min(e0 + u * RT_THREADS, E - 1);
          // e = ((kc * BN + nn) << 5) + k
          const int q = e >> 5, k = e & 31;
          const int kc = (int)__umulhi((unsigned)q, a.bn_magic);
          const int nn = q - kc * BN;
          const int j = s_wrow[kc * RT_CHUNK + k], c = s_ncol[nn];
          const bool ok = j >= 0 && j < H && c >= 0;
          off[u] = (e0 + u * RT_THREADS < E) ? kc * wtile + tc_sw128_off(nn, k) : -1;
          src[u] = ok ? c * H + j : 0;
          sc[u] = ok ? s_nscale[nn] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(a.W + src[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (off[u] >= 0) {
            const float v = w[u] * sc[u];
            *reinterpret_cast<uint32_t*>(Whi + off[u]) = __float_as_uint(v);
            *reinterpret_cast<uint32_t*>(Wlo + off[u]) = tc_lo(__float_as_uint(v));
          }
      }
    }
    __syncthreads();                                   // the scratch in the ring is dead before the first TMA load lands
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  RT_STAMP(2);

  if (warp == RT_W_PROD) {
    // =================== TMA producer: operand "hi" tiles ==========================================================
    {
      int slot = 0;
      uint32_t use = 0;
      for (int tq = 0; tq < my_tiles; ++tq) {
        const int row0 = (tile0 + tq) * RT_ROWS;
        for (int kc = 0; kc < NKC; ++kc) {
          if (use > 0) mbar_wait_bounded(&slot_empty[slot], (use - 1) & 1);
          if (rt_elect()) {
            mbar_expect_tx(&hi_full[slot], RT_STAGE_BYTES);
            rt_tma_load(ring + (size_t)slot * RT_STAGE_BYTES, &a.maps[a.kc[kc].map], a.kc[kc].col0, row0, &hi_full[slot]);
          }
          __syncwarp();
          if (++slot == NST) { slot = 0; ++use; }
        }
      }
    }
  } else if (warp == RT_W_MMA) {
    // =================== MMA issuer ================================================================================
    {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(RT_ROWS >> 4) << 24);
      const uint32_t whi_addr = smem_u32(Whi), wlo_addr = smem_u32(Wlo), ring_addr = smem_u32(ring);
      int slot = 0;
      uint32_t ph = 0;
      for (int tq = 0; tq < my_tiles; ++tq) {
        const int tb = tq & 1;
        if (tq >= 2) { mbar_wait_bounded(&tm_empty[tb], (uint32_t)((tq >> 1) - 1) & 1u); tc_fence_after(); }
        const uint32_t dcol = tmem_d + (uint32_t)(tb * BN);
        uint32_t acc = 0;
        for (int kc = 0; kc < NKC; ++kc) {
          mbar_wait_bounded(&ops_full[slot], ph);
          tc_fence_after();
          const uint64_t dah = tc_desc_sw128(ring_addr + slot * RT_STAGE_BYTES);
          const uint32_t alo = tmem_d + lo_col0 + (uint32_t)(32 * slot);
          const uint64_t dbh = tc_desc_sw128(whi_addr + kc * wtile);
          const uint64_t dbl = tc_desc_sw128(wlo_addr + kc * wtile);
          const int k8n = a.kc[kc].k8;
          if (rt_elect()) {
            for (int k8 = 0; k8 < k8n; ++k8) {         // 8 fp32 = 32 bytes = 2 descriptor units along K inside the swizzle atom
              const uint64_t adv = (uint64_t)(2 * k8);
              rt_mma_tf32_ts(dcol, alo + (uint32_t)(8 * k8), dbh + adv, idesc, acc);
              acc = 1u;
              tc_mma_tf32(dcol, dah + adv, dbl + adv, idesc, 1u);
              tc_mma_tf32(dcol, dah + adv, dbh + adv, idesc, 1u);
            }
            tc_commit(&slot_empty[slot]);
            if (kc == NKC - 1) tc_commit(&tm_full[tb]);
          }
          __syncwarp();
          acc = 1u;
          if (++slot == NST) { slot = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == RT_W_OUT) {
    // =================== out thread: side-input preload + store of the finished output stages ============================
    {
      // chunk g's stage is released once the store of chunk g - NOS has read it: stores are issued NOS - 2 chunks behind the
      // releases, so "all but the latest store group have been read" is enough and one store is always in flight
      const int total = my_tiles * NOC, lag = NOS - 2;
      for (int g = 0; g < total + lag; ++g) {
        const bool leader = rt_elect();              // the same lane every time: bulk-store groups are per thread
        if (g < total) {
          const int s = g % NOS, tq = g / NOC, o = g - tq * NOC;
          if (leader) {
            if (g >= NOS) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (a.oc[o].aux_map >= 0) {
              mbar_expect_tx(&aux_full[s], RT_STAGE_BYTES);
              rt_tma_load(outst + (size_t)s * RT_STAGE_BYTES, &a.maps[a.oc[o].aux_map], a.oc[o].aux_col0, (tile0 + tq) * RT_ROWS, &aux_full[s]);
            } else {
              mbar_arrive(&aux_full[s]);
            }
          }
          __syncwarp();
        }
        const int h = g - lag;
        if (h >= 0 && h < total) {
          const int s = h % NOS, tq = h / NOC, o = h - tq * NOC;
          mbar_wait_bounded(use_col ? &col_done[s] : &out_full[s], (uint32_t)(h / NOS) & 1u);
          if (leader) {
            rt_tma_store(&a.maps[a.oc[o].out_map], a.oc[o].out_col0, (tile0 + tq) * RT_ROWS, outst + (size_t)s * RT_STAGE_BYTES);
            bulk_commit();
          }
          __syncwarp();
        }
      }
      if (rt_elect()) bulk_wait0();
    }
  } else if (warp == RT_W_CSR) {
    // =================== CSR loader: the tile's slice of the dst-CSR (row pointers, tile-local sources, weights) ==============
    if (fuse) {
      for (int tq = 0; tq < my_tiles; ++tq) {
        const int b = tq & 1, tile = tile0 + tq;
        if (tq >= 2) mbar_wait_bounded(&csr_empty[b], (uint32_t)((tq >> 1) - 1) & 1u);
        const int start = a.g_arc0[tile] & ~7;         // 16-byte aligned window start (shorts and floats)
        if (rt_elect()) {
          uint8_t* buf = csrb + (size_t)b * RT_CSR_BUF;
          mbar_expect_tx(&csr_full[b], RT_CSR_RP_BYTES + RT_CSR_LI_BYTES + (a.g_w ? RT_CSR_W_BYTES : 0));
          bulk_g2s(buf, a.g_rowptr + (size_t)tile * RT_ROWS, RT_CSR_RP_BYTES, &csr_full[b]);
          bulk_g2s(buf + 544, a.g_lidx + start, RT_CSR_LI_BYTES, &csr_full[b]);
          if (a.g_w) bulk_g2s(buf + 544 + 2080, a.g_w + start, RT_CSR_W_BYTES, &csr_full[b]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= RT_W_CONV && warp < RT_W_CONV + 4) {
    // =================== converters: thread = row, lo tile of the landed stage ==============================================
    const int r = tid - 32 * RT_W_CONV;                 // row of the tile = TMEM lane (warp w may touch lanes 32 (w & 3) ..)
    const int rbase = (r >> 3) * 1024 + (r & 7) * 128, rx = r & 7;
    const uint32_t lo_lane = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + lo_col0;
    int slot = 0;
    uint32_t ph = 0;
    for (int tq = 0; tq < my_tiles; ++tq) {
      for (int kc = 0; kc < NKC; ++kc) {
        mbar_wait_bounded(&hi_full[slot], ph);
        const uint8_t* hi = ring + (size_t)slot * RT_STAGE_BYTES + rbase;
        const int k8n = a.kc[kc].k8;                   // 8-column K steps that carry data
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < k8n) {
            const uint4 v0 = *reinterpret_cast<const uint4*>(hi + (((2 * q) ^ rx) << 4));
            const uint4 v1 = *reinterpret_cast<const uint4*>(hi + (((2 * q + 1) ^ rx) << 4));
            uint32_t w[8];
            w[0] = tc_lo(v0.x); w[1] = tc_lo(v0.y); w[2] = tc_lo(v0.z); w[3] = tc_lo(v0.w);
            w[4] = tc_lo(v1.x); w[5] = tc_lo(v1.y); w[6] = tc_lo(v1.z); w[7] = tc_lo(v1.w);
            rt_tmem_st8(lo_lane + (uint32_t)(32 * slot + 8 * q), w);
          }
        }
        rt_tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ops_full[slot]);
        if (++slot == NST) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // =================== epilogue: thread = row ==================================================================================
    const int r = tid;
    const int rbase = (r >> 3) * 1024 + (r & 7) * 128, rx = r & 7;
    const uint32_t lane_addr = (uint32_t)(32 * warp) << 16;
    const bool selu = a.act == GNNFP_ACT_SELU;
    const bool conv = MODE == RT_FWD && a.flag_next != nullptr;
    int os = 0;
    uint32_t oph = 0;
    int notconv = 0;
    for (int tq = 0; tq < my_tiles; ++tq) {
      const int tb = tq & 1;
      mbar_wait_bounded(&tm_full[tb], (uint32_t)(tq >> 1) & 1u);
      tc_fence_after();
      float sd = 0.f, sp = 0.f;
      for (int o = 0; o < NOC; ++o) {
        const int width = a.oc[o].width;
        const bool has_aux = a.oc[o].aux_map >= 0;
        float acc[32];
        rt_tmem_ld32(tmem_d + lane_addr + (uint32_t)(tb * BN + a.oc[o].acc_col0), acc);
        mbar_wait_bounded(&aux_full[os], oph);         // the stage is free, the side input (if any) has landed
        tmem_ld_wait();
        if (o == NOC - 1) { tc_fence_before(); mbar_arrive(&tm_empty[tb]); }
        uint8_t* st = outst + (size_t)os * RT_STAGE_BYTES + rbase;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          if (4 * l < width) {
            float4* p4 = reinterpret_cast<float4*>(st + ((l ^ rx) << 4));
            float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_aux) pv = *p4;
            const float pa[4] = {pv.x, pv.y, pv.z, pv.w};
            float v[4];
            if (MODE == RT_FWD) {
              const float4 b4 = *reinterpret_cast<const float4*>(&s_tab[o][0][4 * l]);
              const float ba[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const float z = acc[4 * l + jj] + ba[jj];
                float y = selu ? tc_selu(z) : act_fwd(a.act, z);
                const float pj = 4 * l + jj < width ? pa[jj] : 0.f;
                if (4 * l + jj >= width) y = 0.f;
                const float d = y - pj;
                sd = fmaf(d, d, sd);
                sp = fmaf(pj, pj, sp);
                v[jj] = y;
              }
            } else {
              const float4 k0 = *reinterpret_cast<const float4*>(&s_tab[o][0][4 * l]);
              const float4 k1 = *reinterpret_cast<const float4*>(&s_tab[o][1][4 * l]);
              const float4 kA = *reinterpret_cast<const float4*>(&s_tab[o][2][4 * l]);
              const float4 kB = *reinterpret_cast<const float4*>(&s_tab[o][3][4 * l]);
              v[0] = acc[4 * l + 0] - (k0.x + fmaf(pa[0], kA.x, kB.x) * k1.x);
              v[1] = acc[4 * l + 1] - (k0.y + fmaf(pa[1], kA.y, kB.y) * k1.y);
              v[2] = acc[4 * l + 2] - (k0.z + fmaf(pa[2], kA.z, kB.z) * k1.z);
              v[3] = acc[4 * l + 3] - (k0.w + fmaf(pa[3], kA.w, kB.w) * k1.w);
            }
            *p4 = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(&out_full[os]);
        if (++os == NOS) { os = 0; oph ^= 1u; }
      }
      if (conv && (tile0 + tq) * RT_ROWS + r < n && sqrtf(sd) > a.thr * sqrtf(sp)) notconv = 1;
    }
    if (conv && notconv) notconv_s = 1;
  } else if (warp >= RT_W_COL && warp < RT_W_COL + 4) {
    // =================== column warps: lane = column, statistics of the finished output stage ================================
    const int cw = warp - RT_W_COL;
    double s1[RT_STAT_SLOTS], s2[RT_STAT_SLOTS], g1[RT_STAT_SLOTS], g2[RT_STAT_SLOTS];
    float padv[3] = {0.f, 0.f, 0.f};                   // this row's first Adj^T S_t columns (see the pad patch below)
#pragma unroll
    for (int q = 0; q < RT_STAT_SLOTS; ++q) { s1[q] = 0.0; s2[q] = 0.0; g1[q] = 0.0; g2[q] = 0.0; }
    if (use_col) {
      int os = 0;
      uint32_t oph = 0;
      const int coff = ((lane >> 2) << 4), cin = (lane & 3) << 2;
      const bool want_s = a.ost_sum != nullptr;
      for (int tq = 0; tq < my_tiles; ++tq) {
        const int row0 = (tile0 + tq) * RT_ROWS;
        const int rfirst = row0 + 32 * cw;
        const int nv = min(32, max(0, n - rfirst));
        // staged CSR slice of this tile (fused aggregation)
        const int* rp = nullptr;
        const short* li = nullptr;
        const float* wv = nullptr;
        int arcs0 = 0;
        bool over = false;
        if (fuse) {
          const int b = tq & 1;
          mbar_wait_bounded(&csr_full[b], (uint32_t)(tq >> 1) & 1u);
          const uint8_t* buf = csrb + (size_t)b * RT_CSR_BUF;
          rp = reinterpret_cast<const int*>(buf) + 32 * cw;
          arcs0 = reinterpret_cast<const int*>(buf)[0];
          li = reinterpret_cast<const short*>(buf + 544) + (arcs0 & 7);
          wv = a.g_w ? reinterpret_cast<const float*>(buf + 544 + 2080) + (arcs0 & 7) : nullptr;
          over = reinterpret_cast<const int*>(buf)[RT_ROWS] - arcs0 > GNNFP_TILE_ARCS;   // graph.cu put the whole tile on the row list
        }
        for (int o = 0; o < NOC; ++o) {
          mbar_wait_bounded(&out_full[os], oph);
          const uint8_t* stg = outst + (size_t)os * RT_STAGE_BYTES;
          const bool colok = lane < a.oc[o].width;
          float p1 = 0.f, p2 = 0.f, q1 = 0.f, q2 = 0.f;
          if (want_s && colok) {
            const uint8_t* st = stg + (4 * cw) * 1024;
#pragma unroll 8
            for (int rr = 0; rr < nv; ++rr) {
              const float x = *reinterpret_cast<const float*>(st + (rr >> 3) * 1024 + (rr & 7) * 128 + (coff ^ ((rr & 7) << 4)) + cin);
              p1 += x;
              p2 = fmaf(x, x, p2);
            }
          }
          if (fuse && !over) {
            // Adj^T S_t, thread = row (row rfirst + lane): the row's in-arcs in dst-CSR order (ascending arc id, sequential
            // fmaf = the summation order of agg_stats_kernel and of TF's SparseTensorDenseMatMul); 128-bit swizzled reads of
            // the neighbours' rows.  (lane = column, as in the statistics pass, costs the index / pointer work once per row
            // and WARP instead of once per row and lane: measured 4x the whole kernel's time.)
            const bool rowok = lane < nv;
            const int a0 = rp[rowok ? lane : 0] - arcs0;
            const int na = rowok ? rp[lane + 1] - rp[lane] : 0;
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
            bool good = true;
            for (int q = 0; q < na; ++q) {
              int l = li[a0 + q];
              const float w = wv ? wv[a0 + q] : 1.0f;
              good = good && l >= 0;
              l = l < 0 ? 0 : l;
              const uint8_t* rowp = stg + (l >> 3) * 1024 + (l & 7) * 128;
              const int key = l & 7;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                const float4 v = *reinterpret_cast<const float4*>(rowp + ((ch ^ key) << 4));
                acc[4 * ch + 0] = fmaf(w, v.x, acc[4 * ch + 0]);
                acc[4 * ch + 1] = fmaf(w, v.y, acc[4 * ch + 1]);
                acc[4 * ch + 2] = fmaf(w, v.z, acc[4 * ch + 2]);
                acc[4 * ch + 3] = fmaf(w, v.w, acc[4 * ch + 3]);
              }
            }
            // (compute-sanitizer racecheck reports the 128-bit reads above against the padding patch below, which another
            //  column warp may apply to ITS rows of the same stage meanwhile: the overlap is confined to the padding columns
            //  of the last chunk, whose gathered values land in acc[j >= width] and are never stored or counted)
            const bool wr = rowok && good;
            const int width = a.oc[o].width;
            // The TMA store of S_t works in 16-byte units: when D is not a multiple of 4 the last unit of an S_t row also
            // covers the first (4 - D % 4) columns of the neighbouring Adj^T S_t block of the X slot and would overwrite them
            // with the stage's padding.  The row's own values are therefore patched into the stage's padding columns before
            // the store (rows left to the row-list pass get zeros here and their full row there, later in stream order).
            if (o == 0) { padv[0] = wr ? acc[0] : 0.f; padv[1] = wr ? acc[1] : 0.f; padv[2] = wr ? acc[2] : 0.f; }
            if (o == NOC - 1 && (width & 3) != 0 && rowok) {
              float* strow = reinterpret_cast<float*>(const_cast<uint8_t*>(stg) + ((4 * cw + (lane >> 3)) * 1024) + (lane & 7) * 128 +
                                                     (((width >> 2) ^ (lane & 7)) << 4));
              const int p0 = width & 3;
#pragma unroll
              for (int p = 1; p < 4; ++p) if (p >= p0) strow[p] = padv[p - p0];
              fence_proxy_async();
            }
            float* outp = a.agg_out + (size_t)(rfirst + lane) * a.ld_agg + a.oc[o].out_col0;
            const bool v2 = a.ld_agg % 2 == 0 && ((reinterpret_cast<uintptr_t>(a.agg_out) + 4 * (size_t)a.oc[o].out_col0) & 7) == 0;
            if (wr) {
              if (v2) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  if (j + 1 < width) *reinterpret_cast<float2*>(outp + j) = make_float2(acc[j], acc[j + 1]);
                  else if (j < width) outp[j] = acc[j];
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (j < width) outp[j] = acc[j];
              }
            }
            if (a.agg_sum) {                           // column sums over the warp's rows: lane j ends up with column j
              float sq[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) { acc[j] = (wr && j < width) ? acc[j] : 0.f; sq[j] = acc[j] * acc[j]; }
              q1 = warp_colsum32(acc, lane);
              q2 = warp_colsum32(sq, lane);
            }
          }
          mbar_arrive(&col_done[os]);
          const int slot = a.oc[o].st_slot;
#pragma unroll
          for (int q = 0; q < RT_STAT_SLOTS; ++q)
            if (q == slot) { s1[q] += (double)p1; s2[q] += (double)p2; g1[q] += (double)q1; g2[q] += (double)q2; }
          if (++os == NOS) { os = 0; oph ^= 1u; }
        }
        if (fuse) mbar_arrive(&csr_empty[tq & 1]);
      }
#pragma unroll
      for (int q = 0; q < RT_STAT_SLOTS; ++q) {
        s_stat[cw][q][0][lane] = s1[q]; s_stat[cw][q][1][lane] = s2[q];
        s_stat[cw][q][2][lane] = g1[q]; s_stat[cw][q][3][lane] = g2[q];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  RT_STAMP(3);
  if (MODE == RT_FWD) {
    if (a.flag_next && tid == 0 && notconv_s) atomicOr(a.flag_next, 1);
    if (use_col) {
      for (int e = tid; e < NOC * 32; e += RT_THREADS) {
        const int o = e >> 5, j = e & 31;
        if (j < a.oc[o].width) {
          const int q = a.oc[o].st_slot, c = a.oc[o].cidx0 + j;
          if (a.ost_sum) {
            atomicAdd(a.ost_sum + c, (s_stat[0][q][0][j] + s_stat[1][q][0][j]) + (s_stat[2][q][0][j] + s_stat[3][q][0][j]));
            atomicAdd(a.ost_sq + c, (s_stat[0][q][1][j] + s_stat[1][q][1][j]) + (s_stat[2][q][1][j] + s_stat[3][q][1][j]));
          }
          if (fuse && a.agg_sum) {
            atomicAdd(a.agg_sum + c, (s_stat[0][q][2][j] + s_stat[1][q][2][j]) + (s_stat[2][q][2][j] + s_stat[3][q][2][j]));
            atomicAdd(a.agg_sq + c, (s_stat[0][q][3][j] + s_stat[1][q][3][j]) + (s_stat[2][q][3][j] + s_stat[3][q][3][j]));
          }
        }
      }
    }
  }
  if (warp == RT_W_MMA) tmem_dealloc(tmem_d, (uint32_t)a.tmem_cols);
  RT_STAMP(4);
}

int dw_tma_times(long long* out);
extern "C" int gnnfp_debug_rt_times(long long* out, int mode) {      // out[160][8] of the last launch of `mode` (2 = dw_tma.cu)
  GNNFP_CHECK_CUDA(cudaDeviceSynchronize());
  if (mode == 2) return dw_tma_times(out);
  GNNFP_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_rt_times, (size_t)160 * 8 * sizeof(long long), (size_t)(mode ? 1 : 0) * 160 * 8 * sizeof(long long)));
  return GNNFP_OK;
}

// ---- host side ------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled rt_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}
int rows_tma_available() {
  static const int off = getenv("GNNFP_NO_TMA") ? 1 : 0;
  return !off && rt_encoder() != nullptr;
}
int rows_tma_ok(const float* ptr, int ld) { return ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 4 == 0 && ld > 0; }

int rows_tma_map(CUtensorMap* m, const float* ptr, int rows, int cols, int ld, int box_rows, int atom32) {
  PFN_cuTensorMapEncodeTiled fn = rt_encoder();
  if (!fn) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "rows_tma: cuTensorMapEncodeTiled is not available from this driver");
  if (!rows_tma_ok(ptr, ld) || rows < 1 || cols < 1 || cols > ld)
    GNNFP_FAIL(GNNFP_E_INVALID, "rows_tma: matrix %p [%d x %d, ld %d] cannot be described by a tensor map (16-byte alignment)", (const void*)ptr, rows, cols, ld);
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {RT_CHUNK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) GNNFP_FAIL(GNNFP_E_CUDA, "cuTensorMapEncodeTiled failed with %d ([%d x %d], ld %d)", (int)r, rows, cols, ld);
  return GNNFP_OK;
}

static size_t rt_smem_cap(int mode) {              // dynamic shared memory a launch of this mode may use next to the kernel's static arrays
  static size_t cap[2] = {0, 0};
  const int mi = mode == RT_DX ? 1 : 0;
  if (!cap[mi]) {
    cudaFuncAttributes fa;
    int dev = 0, optin = 0;
    const cudaError_t e = mi ? cudaFuncGetAttributes(&fa, rows_tma_kernel<RT_DX>) : cudaFuncGetAttributes(&fa, rows_tma_kernel<RT_FWD>);
    if (e != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    cap[mi] = (size_t)optin - fa.sharedSizeBytes;
  }
  return cap[mi];
}
size_t rows_tma_smem(const RowsTmaArgs& a) {
  return (size_t)2 * a.n_kc * a.BN * 128 + (size_t)a.n_stages * RT_STAGE_BYTES + (size_t)a.n_ostages * RT_STAGE_BYTES +
         (a.fuse_agg ? (size_t)2 * RT_CSR_BUF : 0) + 1024;
}
int rows_tma_finish(RowsTmaArgs& a) {
  if (a.n_kc < 1 || a.n_kc > RT_MAXKC || a.n_oc < 1 || a.n_oc > RT_MAXOC || a.BN < 16 || a.BN > RT_MAXBN || a.BN % 16 != 0)
    GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "rows_tma: shape outside the kernel's limits (K chunks %d, output chunks %d, N %d)", a.n_kc, a.n_oc, a.BN);
  const size_t cap = rt_smem_cap(a.mode);
  if (!cap) GNNFP_FAIL(GNNFP_E_CUDA, "rows_tma: cannot query the shared-memory budget");
  a.bn_magic = (unsigned)((0x100000000ull + (unsigned)a.BN - 1) / (unsigned)a.BN);   // exact for the < 2^17 indices of the weight build
  a.n_ostages = 2;
  a.n_stages = 2;
  if (rows_tma_smem(a) > cap) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "rows_tma: %zu bytes of shared memory needed, %zu available", rows_tma_smem(a), cap);
  static const int st_max = getenv("GNNFP_RT_STAGES") ? atoi(getenv("GNNFP_RT_STAGES")) : 6;
  // output stages first: the side-input preload -> epilogue -> statistics -> store chain of a chunk is ~2 us long, a deeper
  // operand ring measured no gain beyond 2 stages
  while (a.n_ostages < 4) { ++a.n_ostages; if (rows_tma_smem(a) > cap) { --a.n_ostages; break; } }
  // ring depth: bounded by shared memory (16 KB per stage) and by tensor memory (32 columns of a_lo per stage)
  while (a.n_stages < st_max && a.n_stages < 8 && 2 * a.BN + 32 + 32 * (a.n_stages + 1) <= 512) {
    ++a.n_stages;
    if (rows_tma_smem(a) > cap) { --a.n_stages; break; }
  }
  const int need = 2 * a.BN + 32 + 32 * a.n_stages;
  a.tmem_cols = need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512));
  return GNNFP_OK;
}

int launch_rows_tma(const RowsTmaArgs& a, cudaStream_t s, int prof_cat) {
  if (a.n_rows <= 0) return GNNFP_OK;
  if (a.mode == RT_FWD && a.BN > 112) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "rows_tma: forward mode takes at most 112 output columns (%d)", a.BN);
  const size_t smem = rows_tma_smem(a);
  static size_t attr[2] = {0, 0};
  const int mi = a.mode == RT_DX ? 1 : 0;
  if (smem > attr[mi]) {
    if (mi) GNNFP_CHECK_CUDA(cudaFuncSetAttribute(rows_tma_kernel<RT_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else GNNFP_CHECK_CUDA(cudaFuncSetAttribute(rows_tma_kernel<RT_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[mi] = smem;
  }
  const int n_tiles = (a.n_rows + RT_ROWS - 1) / RT_ROWS;
  const int nsm = gnnfp_num_sms();
  const int grid = n_tiles < nsm ? n_tiles : nsm;
  ProfScope ps(prof_cat, s);
  if (mi) rows_tma_kernel<RT_DX><<<grid, RT_THREADS, smem, s>>>(a);
  else rows_tma_kernel<RT_FWD><<<grid, RT_THREADS, smem, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
