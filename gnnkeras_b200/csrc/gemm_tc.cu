// gemm_tc.cu - the row GEMM of gemm.cu on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   OUT[rows, N] = epilogue( [piece_0 | piece_1 | ...][rows, K] . Wp[K, N] )        (same contract as gemm_rows_kernel)
//
// FP32 accuracy is kept with the 3-term TF32 split  a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi  (a_hi = tf32(a), a_lo =
// tf32(a - a_hi)): three MMAs per 8-wide K step accumulate into the same FP32 TMEM tile, the dropped a_lo.b_lo term is
// ~2^-22 relative.  Per persistent CTA (one per SM, 256 threads):
//   * the padded weights are split once into W_hi / W_lo, stored K-major in the canonical SWIZZLE_128B layout
//     (one [BN x 32] tile of 128-byte rows per 32-wide K block) and stay resident;
//   * activation rows stream through a 2-deep cp.async ring (raw fp32, 128 rows x 32 columns per stage); all threads
//     split a landed stage into A_hi / A_lo operand tiles (same swizzled layout, double buffered), fence the async proxy,
//     and ONE thread issues the 12 tcgen05.mma of the stage and commits them to an mbarrier (operand buffer free);
//   * after the last K block of a 128-row tile the accumulator tile (128 lanes x BN columns) is read back with
//     tcgen05.ld (thread = row), bias + activation / column scale applied, staged through shared memory for coalesced
//     stores, the per-row convergence test and the next BN's column statistics.
// Operands cannot come by TMA here: the state matrices have leading dimension D (e.g. 78 floats = 312 B), not a multiple
// of 16 bytes, so the tile is gathered with 8-byte cp.async and laid out by the split pass.
#include "tile.cuh"

#include "gemm.h"

#define TC_BM 128
#define TC_BK 32
#define TC_RAW_LD 32                 // floats per raw row (a quarter warp reads one 128-byte row: conflict free)
#define TC_RAW_STAGES 3
#define TC_TILE_BYTES (TC_BM * 128)  // one A operand tile: 128 rows x 128 B

__device__ __forceinline__ void tc_cp_async8(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tc_cp_async4(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tc_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bounded mbarrier wait: a protocol error traps (the launch fails) instead of hanging the device
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 26); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, TF32 inputs, FP32 accumulate; issued by one thread for the CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t tc_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);      // start address, 16-byte units          bits [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused: swizzled K-major)  [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: next 8-row group  bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)        bits [46,48)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B              bits [61,64)
  return d;
}
// byte offset of element (row, col) of a [rows x 32 fp32] K-major SWIZZLE_128B tile (16-byte chunk index XOR row%8)
__device__ __forceinline__ int tc_sw128_off(int row, int col) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 2) ^ (row & 7)) & 7) << 4) + ((col & 3) << 2);
}
__device__ __forceinline__ uint32_t tc_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float tc_selu(float z) {
  const float e = expf(fminf(z, 0.0f));
  return z < 0.0f ? (SELU_SCALE_F * SELU_ALPHA_F) * (e - 1.0f) : SELU_SCALE_F * z;
}

template <int BN, bool FWD>
__global__ void __launch_bounds__(256, 1) gemm_rows_tc_kernel(const __grid_constant__ GemmRowsArgs a) {
  if (a.gate && *a.gate == 0) return;
  constexpr int CH = BN / 2;                           // accumulator columns per epilogue warp (two warps share a lane quarter)
  constexpr int OLD = BN + 1;                          // leading dimension of the output staging tile
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; offset arithmetic (not an integer round trip) keeps the pointers in the
  // shared address space for the compiler (LDS/STS instead of generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NKB = (a.Kpad + TC_BK - 1) / TC_BK;
  const int wtile = BN * 128;                          // bytes of one [BN x 32] weight tile
  uint8_t* Whi = base;
  uint8_t* Wlo = Whi + (size_t)NKB * wtile;
  uint8_t* Aop = Wlo + (size_t)NKB * wtile;            // [2 buffers][hi, lo][TC_TILE_BYTES]
  float* raw = reinterpret_cast<float*>(Aop + 4 * TC_TILE_BYTES);     // [TC_RAW_STAGES][TC_BM][TC_RAW_LD]
  float* ost = reinterpret_cast<float*>(Aop);          // output staging [TC_BM][OLD] aliases the operand buffers
  __shared__ __align__(8) uint64_t bar_ops[2];
  __shared__ __align__(8) uint64_t bar_tile;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[BN];
  __shared__ double colacc[2][BN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int n = a.n_rows;
  const int n_tiles = (n + TC_BM - 1) / TC_BM;
  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int my_tiles = max(0, min(n_tiles, tile0 + tiles_per_cta) - tile0);

  if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }      // whole warp, converged (.sync.aligned)
  if (tid == 32) { mbar_init(&bar_ops[0], 1); mbar_init(&bar_ops[1], 1); mbar_init(&bar_tile, 1); }

  // ---- producer: raw stage (tile, K block) via 8-byte cp.async; thread copies column pair pq of rows ty + 16*i ------
  const int pq = tx;
  int i_tq = 0, i_kb = 0, i_st = 0;
  auto issue = [&]() {
    if (i_tq < my_tiles) {
      if (i_kb == 0 && i_tq + 1 < my_tiles) {
        // L2 prefetch of the NEXT tile's rows (all pieces): the cp.async of the coming stages then pay L2, not DRAM, latency
        const int prow = (tile0 + i_tq + 1) * TC_BM + (tid >> 1);
        if (prow < n) {
          for (int p = 0; p < a.n_pieces; ++p) {
            const char* b0 = reinterpret_cast<const char*>(a.p[p].ptr + (size_t)prow * a.p[p].ld);
            const char* l0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(b0) & ~(uintptr_t)127);
            const char* l1 = b0 + (size_t)a.p[p].width * 4;
            for (const char* l = l0 + 128 * (tid & 1); l < l1; l += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(l));
          }
        }
      }
      float* dst = raw + i_st * TC_BM * TC_RAW_LD + ty * TC_RAW_LD + 2 * pq;
      const int kcol = i_kb * TC_BK + 2 * pq;
      int p = 0;
      while (p + 1 < a.n_pieces && kcol >= a.p[p + 1].k8) ++p;
      const float* pptr = a.p[p].ptr;
      const int pld = a.p[p].ld;
      const int kk = kcol - a.p[p].k8;
      const int nv = kcol < a.Kpad ? a.p[p].width - kk : 0;
      const bool al8 = a.p[p].al8 != 0;
      const int row0 = (tile0 + i_tq) * TC_BM + ty;
      if (nv <= 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) tc_cp_async8(dst + i * 16 * TC_RAW_LD, pptr, 0);
      } else if (al8 && (tile0 + i_tq + 1) * TC_BM <= n) {
        const float* src = pptr + (size_t)row0 * pld + kk;
        const size_t step = (size_t)16 * pld;
        const int bytes = nv > 1 ? 8 : 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) { tc_cp_async8(dst + i * 16 * TC_RAW_LD, src, bytes); src += step; }
      } else {
#pragma unroll 2
        for (int i = 0; i < 8; ++i) {
          const int grow = row0 + 16 * i;
          const bool valid = grow < n;
          const float* src = valid ? pptr + (size_t)grow * pld + kk : pptr;
          if (al8) {
            tc_cp_async8(dst + i * 16 * TC_RAW_LD, src, valid ? (nv > 1 ? 8 : 4) : 0);
          } else {
            tc_cp_async4(dst + i * 16 * TC_RAW_LD, src, valid ? 4 : 0);
            tc_cp_async4(dst + i * 16 * TC_RAW_LD + 1, (valid && nv > 1) ? src + 1 : pptr, (valid && nv > 1) ? 4 : 0);
          }
        }
      }
      if (++i_kb == NKB) { i_kb = 0; ++i_tq; }
      if (++i_st == TC_RAW_STAGES) i_st = 0;
    }
    tc_cp_commit();
  };
  for (int i = 0; i < TC_RAW_STAGES - 1; ++i) issue();

  // ---- resident weights: split into TF32 hi / lo, K-major swizzled tiles per K block ------------------------------------
  for (int e = tid; e < NKB * TC_BK * BN; e += 256) {
    const int k = e / BN, nn = e - k * BN;
    const float w = k < a.Kpad ? a.Wp[(size_t)k * a.ldw + nn] : 0.f;
    const uint32_t hi = tc_tf32(w);
    const uint32_t lo = tc_tf32(w - __uint_as_float(hi));
    const int off = (k >> 5) * wtile + tc_sw128_off(nn, k & 31);
    *reinterpret_cast<uint32_t*>(Whi + off) = hi;
    *reinterpret_cast<uint32_t*>(Wlo + off) = lo;
  }
  for (int j = tid; j < BN; j += 256) sbias[j] = (FWD && a.bias && j < a.N) ? a.bias[j] : 0.f;
  for (int j = tid; j < 2 * BN; j += 256) (&colacc[0][0])[j] = 0.0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = BN, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  const uint32_t whi_addr = smem_u32(Whi), wlo_addr = smem_u32(Wlo), aop_addr = smem_u32(Aop);

  int notconv = 0;
  int c_st = 0;                                        // raw ring slot of the current stage
  uint32_t use0 = 0, use1 = 0;                         // how often each operand buffer has been handed to the tensor core
  int sidx = 0;                                        // running stage index (operand buffer = sidx & 1)
  for (int tq = 0; tq < my_tiles; ++tq) {
    for (int kb = 0; kb < NKB; ++kb, ++sidx) {
      issue();                                         // prefetch: TC_RAW_STAGES - 1 stages ahead
      tc_cp_wait<TC_RAW_STAGES - 1>();
      __syncthreads();                                 // (A) raw stage landed for every thread; previous epilogue done
      const int ob = sidx & 1;
      const uint32_t uses = ob ? use1 : use0;
      if (uses > 0) {                                  // MMAs of the previous use of this operand buffer must have completed
        mbar_wait_bounded(&bar_ops[ob], (uses - 1) & 1);
        tc_fence_after();
      }
      // ---- split raw fp32 -> TF32 hi / lo operand tiles (swizzled) ---------------------------------------------------
      const float* rs = raw + c_st * TC_BM * TC_RAW_LD;
      uint8_t* ahi = Aop + (size_t)ob * 2 * TC_TILE_BYTES;
      uint8_t* alo = ahi + TC_TILE_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int it = tid + 256 * j;
        const int row = it >> 3, ch = it & 7;
        const float4 v = *reinterpret_cast<const float4*>(rs + row * TC_RAW_LD + 4 * ch);
        uint4 h, l;
        h.x = tc_tf32(v.x); l.x = tc_tf32(v.x - __uint_as_float(h.x));
        h.y = tc_tf32(v.y); l.y = tc_tf32(v.y - __uint_as_float(h.y));
        h.z = tc_tf32(v.z); l.z = tc_tf32(v.z - __uint_as_float(h.z));
        h.w = tc_tf32(v.w); l.w = tc_tf32(v.w - __uint_as_float(h.w));
        const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(ahi + off) = h;
        *reinterpret_cast<uint4*>(alo + off) = l;
      }
      if (++c_st == TC_RAW_STAGES) c_st = 0;
      fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncthreads();                                 // (B)
      if (tid == 0) {
        tc_fence_after();
        const uint64_t dah = tc_desc_sw128(aop_addr + ob * 2 * TC_TILE_BYTES);
        const uint64_t dal = tc_desc_sw128(aop_addr + ob * 2 * TC_TILE_BYTES + TC_TILE_BYTES);
        const uint64_t dbh = tc_desc_sw128(whi_addr + kb * wtile);
        const uint64_t dbl = tc_desc_sw128(wlo_addr + kb * wtile);
#pragma unroll
        for (int k8 = 0; k8 < TC_BK / 8; ++k8) {       // 8 fp32 = 32 bytes = 2 descriptor units along K inside the swizzle atom
          const uint64_t adv = (uint64_t)(2 * k8);
          tc_mma_tf32(tmem_d, dal + adv, dbh + adv, idesc, (kb | k8) ? 1u : 0u);
          tc_mma_tf32(tmem_d, dah + adv, dbl + adv, idesc, 1u);
          tc_mma_tf32(tmem_d, dah + adv, dbh + adv, idesc, 1u);
        }
        tc_commit(&bar_ops[ob]);
        if (kb == NKB - 1) tc_commit(&bar_tile);
      }
      if (ob) ++use1; else ++use0;
    }
    // ---- epilogue of this tile ---------------------------------------------------------------------------------------
    mbar_wait_bounded(&bar_tile, (uint32_t)tq & 1u);
    tc_fence_after();
    {
      const int q = warp & 3, h = warp >> 2;
      const int row = 32 * q + lane;
      const int grow = (tile0 + tq) * TC_BM + row;
      const bool valid = grow < n;
      float acc[CH];
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * CH);
#pragma unroll
      for (int c8 = 0; c8 < CH / 8; ++c8) tmem_ld8(taddr + 8 * c8, acc + 8 * c8);
      tmem_ld_wait();
      const bool selu = a.act == GNNFP_ACT_SELU;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const int col = h * CH + j;
        float v = 0.f;
        if (col < a.N && valid) {
          if (FWD) {
            const float z = acc[j] + sbias[col];
            v = selu ? tc_selu(z) : act_fwd(a.act, z);
          } else {
            v = a.colscale ? acc[j] * a.colscale[col] : acc[j];
          }
        }
        ost[row * OLD + col] = v;
      }
    }
    tc_fence_before();
    __syncthreads();                                   // accumulator drained (next tile may overwrite it), staging complete
    // ---- coalesced pass over the staged tile: one row per warp, 4 rows in flight; stores, convergence test (FWD),
    // ---- BN-training correction / accumulation into the destination (backward) ---------------------------------------
    constexpr int NC = (BN + 31) / 32;                 // column trips of a lane
    for (int r0 = warp * 16; r0 < warp * 16 + 16; r0 += 4) {
      float v[4][NC], aux[4][NC];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u;
        const int grow = (tile0 + tq) * TC_BM + r;
        const bool rv = grow < n;
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          const int c = lane + 32 * cc;
          const bool ok = rv && c < a.N;
          v[u][cc] = ok ? ost[r * OLD + c] : 0.f;
          aux[u][cc] = 0.f;
          if (ok) {
            if (FWD) { if (a.prev) aux[u][cc] = a.prev[(size_t)grow * a.ld_prev + c]; }
            else {
              if (a.corr) aux[u][cc] = a.corr_x[(size_t)grow * a.corr_ld + c];
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u;
        const int grow = (tile0 + tq) * TC_BM + r;
        const bool rv = grow < n;
        float sd = 0.f, sp = 0.f;
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          const int c = lane + 32 * cc;
          if (rv && c < a.N) {
            float* o = a.out + (size_t)grow * a.ld_out + c;
            float x = v[u][cc];
            if (FWD) {
              const float dd = x - aux[u][cc];
              sd = fmaf(dd, dd, sd);
              sp = fmaf(aux[u][cc], aux[u][cc], sp);
            } else {
              if (a.corr) {
                const float* k = a.corr + a.corr_col0;
                x -= k[c] + fmaf(aux[u][cc], k[2 * a.corr_in + c], k[3 * a.corr_in + c]) * k[a.corr_in + c];
              }
              if (a.out_add) x += *o;
            }
            *o = x;
          }
        }
        if (FWD && a.prev) {
#pragma unroll
          for (int of = 16; of > 0; of >>= 1) {
            sd += __shfl_xor_sync(0xffffffffu, sd, of);
            sp += __shfl_xor_sync(0xffffffffu, sp, of);
          }
          if (rv && sqrtf(sd) > a.thr * sqrtf(sp)) notconv = 1;
        }
      }
    }
    if (FWD && a.ost_sum && tid < a.N) {               // column statistics of the tile (rows past n hold zeros)
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
      for (int r = 0; r < TC_BM; ++r) {
        const float x = ost[r * OLD + tid];
        s1 += x;
        s2 = fmaf(x, x, s2);
      }
      colacc[0][tid] += (double)s1;
      colacc[1][tid] += (double)s2;
    }
  }
  tc_cp_wait<0>();
  if (FWD && a.flag_next) {
    const int any = __syncthreads_or(notconv);
    if (tid == 0 && any) atomicOr(a.flag_next, 1);
  }
  __syncthreads();
  if (FWD && a.ost_sum) {
    for (int j = tid; j < a.N; j += 256) {
      atomicAdd(a.ost_sum + j, colacc[0][j]);
      atomicAdd(a.ost_sq + j, colacc[1][j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

// ------------------------------------------------------------------------------------------------------------
template <int BN, bool FWD>
static int launch_tc_t(const GemmRowsArgs& a, cudaStream_t s, int prof_cat) {
  const int NKB = (a.Kpad + TC_BK - 1) / TC_BK;
  const size_t smem = (size_t)2 * NKB * BN * 128 + 4 * TC_TILE_BYTES + (size_t)TC_RAW_STAGES * TC_BM * TC_RAW_LD * sizeof(float) + 1024;
  static size_t attr = 0;
  if (smem > attr) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(gemm_rows_tc_kernel<BN, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_tiles = (a.n_rows + TC_BM - 1) / TC_BM;
  const int nsm = gnnfp_num_sms();
  const int grid = n_tiles < nsm ? n_tiles : nsm;
  ProfScope ps(prof_cat, s);
  gemm_rows_tc_kernel<BN, FWD><<<grid, 256, smem, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}

// eligible shapes: plain row sets, 8-byte-copyable pieces, up to 160 padded K columns and 80 output columns
int gemm_rows_tc_supported(const GemmRowsArgs& a) {
  if (a.rowlist != nullptr || a.out_compact || a.act == GNNFP_ACT_SOFTMAX) return 0;
  if (a.N < 1 || a.N > 80 || a.Kpad < 8 || a.Kpad > 160) return 0;
  if (a.ldw != 16 * ((a.N + 15) / 16)) return 0;
  const int NKB = (a.Kpad + TC_BK - 1) / TC_BK;
  const int BN = a.ldw;
  const size_t smem = (size_t)2 * NKB * BN * 128 + 4 * TC_TILE_BYTES + (size_t)TC_RAW_STAGES * TC_BM * TC_RAW_LD * sizeof(float) + 1024;
  return smem <= 220 * 1024;
}

int launch_gemm_rows_tc(const GemmRowsArgs& a, cudaStream_t s, int prof_cat) {
  if (a.n_rows <= 0) return GNNFP_OK;
  if (!gemm_rows_tc_supported(a)) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows_tc: shape N=%d K=%d not supported", a.N, a.Kpad);
  const bool fwd = a.fwd != 0;
  switch (a.ldw) {
    case 16: return fwd ? launch_tc_t<16, true>(a, s, prof_cat) : launch_tc_t<16, false>(a, s, prof_cat);
    case 32: return fwd ? launch_tc_t<32, true>(a, s, prof_cat) : launch_tc_t<32, false>(a, s, prof_cat);
    case 48: return fwd ? launch_tc_t<48, true>(a, s, prof_cat) : launch_tc_t<48, false>(a, s, prof_cat);
    case 64: return fwd ? launch_tc_t<64, true>(a, s, prof_cat) : launch_tc_t<64, false>(a, s, prof_cat);
    case 80: return fwd ? launch_tc_t<80, true>(a, s, prof_cat) : launch_tc_t<80, false>(a, s, prof_cat);
    default: GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows_tc: %d output columns", a.ldw);
  }
}
