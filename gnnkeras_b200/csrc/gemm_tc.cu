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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// column sums of a [32 lanes (rows)] x [32 values per lane (columns)] block: after the 5 exchange rounds lane l holds the
// total of column l (31 shuffles instead of 32 x 5)
__device__ __forceinline__ float warp_colsum32(float (&x)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? x[i] : x[i + off];
      const float keep = up ? x[i + off] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return x[0];
}

#define TC_THREADS 288                                  // warps 0-3 converters, 4-7 epilogue, 8 MMA issuer

template <int BN, bool FWD>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_rows_tc_kernel(const __grid_constant__ GemmRowsArgs a) {
  if (a.gate && *a.gate == 0) return;
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
  __shared__ __align__(8) uint64_t ops_full[2], ops_empty[2], tm_full[2], tm_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[BN];
  __shared__ double colacc[4][2][BN];                  // per epilogue warp: column sums / sums of squares
  __shared__ float estage[4][32][17];                  // per epilogue warp: [32 rows x 16 columns] transpose staging
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_rows;
  const int n_tiles = (n + TC_BM - 1) / TC_BM;
  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int my_tiles = max(0, min(n_tiles, tile0 + tiles_per_cta) - tile0);
  constexpr uint32_t TMEM_COLS = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : 256));

  if (warp == 0) { tmem_alloc(&tmem_base_s, TMEM_COLS); tmem_relinquish(); }      // whole warp, converged (.sync.aligned)
  if (tid == 32) {
    mbar_init(&ops_full[0], 128); mbar_init(&ops_full[1], 128);
    mbar_init(&ops_empty[0], 1); mbar_init(&ops_empty[1], 1);
    mbar_init(&tm_full[0], 1); mbar_init(&tm_full[1], 1);
    mbar_init(&tm_empty[0], 128); mbar_init(&tm_empty[1], 128);
  }
  // ---- resident weights: split into TF32 hi / lo, K-major swizzled tiles per K block (all threads) ------------------
  for (int e = tid; e < NKB * TC_BK * BN; e += TC_THREADS) {
    const int k = e / BN, nn = e - k * BN;
    const float w = k < a.Kpad ? a.Wp[(size_t)k * a.ldw + nn] : 0.f;
    const uint32_t hi = tc_tf32(w);
    const uint32_t lo = tc_tf32(w - __uint_as_float(hi));
    const int off = (k >> 5) * wtile + tc_sw128_off(nn, k & 31);
    *reinterpret_cast<uint32_t*>(Whi + off) = hi;
    *reinterpret_cast<uint32_t*>(Wlo + off) = lo;
  }
  for (int j = tid; j < BN; j += TC_THREADS) sbias[j] = (FWD && a.bias && j < a.N) ? a.bias[j] : 0.f;
  for (int j = tid; j < 8 * BN; j += TC_THREADS) (&colacc[0][0][0])[j] = 0.0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  int notconv = 0;

  if (warp < 4) {
    // =================== converter warps: cp.async ring -> TF32 hi / lo operand tiles ==================================
    // thread copies the 8-byte column pair pq of rows ty + 8*i (i < 16) of a stage
    const int pq = tid & 15, ty = tid >> 4;
    int i_tq = 0, i_kb = 0, i_st = 0;
    auto issue = [&]() {
      if (i_tq < my_tiles) {
        if (i_kb == 0 && i_tq + 1 < my_tiles) {
          // L2 prefetch of the NEXT tile's rows (all pieces): the coming stages then pay L2, not DRAM, latency
          const int prow = (tile0 + i_tq + 1) * TC_BM + tid;
          if (prow < n) {
            for (int p = 0; p < a.n_pieces; ++p) {
              const char* b0 = reinterpret_cast<const char*>(a.p[p].ptr + (size_t)prow * a.p[p].ld);
              const char* l0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(b0) & ~(uintptr_t)127);
              const char* l1 = b0 + (size_t)a.p[p].width * 4;
              for (const char* l = l0; l < l1; l += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(l));
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
          for (int i = 0; i < 16; ++i) tc_cp_async8(dst + i * 8 * TC_RAW_LD, pptr, 0);
        } else if (al8 && (tile0 + i_tq + 1) * TC_BM <= n) {
          const float* src = pptr + (size_t)row0 * pld + kk;
          const size_t step = (size_t)8 * pld;
          const int bytes = nv > 1 ? 8 : 4;
#pragma unroll
          for (int i = 0; i < 16; ++i) { tc_cp_async8(dst + i * 8 * TC_RAW_LD, src, bytes); src += step; }
        } else {
#pragma unroll 2
          for (int i = 0; i < 16; ++i) {
            const int grow = row0 + 8 * i;
            const bool valid = grow < n;
            const float* src = valid ? pptr + (size_t)grow * pld + kk : pptr;
            if (al8) {
              tc_cp_async8(dst + i * 8 * TC_RAW_LD, src, valid ? (nv > 1 ? 8 : 4) : 0);
            } else {
              tc_cp_async4(dst + i * 8 * TC_RAW_LD, src, valid ? 4 : 0);
              tc_cp_async4(dst + i * 8 * TC_RAW_LD + 1, (valid && nv > 1) ? src + 1 : pptr, (valid && nv > 1) ? 4 : 0);
            }
          }
        }
        if (++i_kb == NKB) { i_kb = 0; ++i_tq; }
        if (++i_st == TC_RAW_STAGES) i_st = 0;
      }
      tc_cp_commit();
    };
    for (int i = 0; i < TC_RAW_STAGES - 1; ++i) issue();
    int c_st = 0;
    const int total = my_tiles * NKB;
    for (int sidx = 0; sidx < total; ++sidx) {
      tc_cp_wait<TC_RAW_STAGES - 2>();
      named_bar_sync(1, 128);                          // stage sidx has landed for all converter threads, and every thread
                                                       // has finished reading the slot of stage sidx-1 ...
      issue();                                         // ... which the prefetch (TC_RAW_STAGES - 1 stages ahead) refills
      const int ob = sidx & 1;
      const uint32_t use = (uint32_t)(sidx >> 1);      // how often this operand buffer has been filled before
      if (use > 0) {                                   // the MMAs that read its previous contents must have completed
        mbar_wait_bounded(&ops_empty[ob], (use - 1) & 1);
        tc_fence_after();
      }
      const float* rs = raw + c_st * TC_BM * TC_RAW_LD;
      uint8_t* ahi = Aop + (size_t)ob * 2 * TC_TILE_BYTES;
      uint8_t* alo = ahi + TC_TILE_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int it = tid + 128 * j;
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
      mbar_arrive(&ops_full[ob]);
    }
    tc_cp_wait<0>();
  } else if (warp == 8) {
    // =================== MMA issuer: one thread drives the tensor core ====================================================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t whi_addr = smem_u32(Whi), wlo_addr = smem_u32(Wlo), aop_addr = smem_u32(Aop);
      int sidx = 0;
      for (int tq = 0; tq < my_tiles; ++tq) {
        const int tb = tq & 1;
        if (tq >= 2) {                                 // the epilogue of the tile that used this accumulator must have drained it
          mbar_wait_bounded(&tm_empty[tb], (uint32_t)((tq >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t dcol = tmem_d + (uint32_t)(tb * BN);
        for (int kb = 0; kb < NKB; ++kb, ++sidx) {
          const int ob = sidx & 1;
          mbar_wait_bounded(&ops_full[ob], (uint32_t)(sidx >> 1) & 1u);
          tc_fence_after();
          const uint64_t dah = tc_desc_sw128(aop_addr + ob * 2 * TC_TILE_BYTES);
          const uint64_t dal = tc_desc_sw128(aop_addr + ob * 2 * TC_TILE_BYTES + TC_TILE_BYTES);
          const uint64_t dbh = tc_desc_sw128(whi_addr + kb * wtile);
          const uint64_t dbl = tc_desc_sw128(wlo_addr + kb * wtile);
#pragma unroll
          for (int k8 = 0; k8 < TC_BK / 8; ++k8) {     // 8 fp32 = 32 bytes = 2 descriptor units along K inside the swizzle atom
            const uint64_t adv = (uint64_t)(2 * k8);
            tc_mma_tf32(dcol, dal + adv, dbh + adv, idesc, (kb | k8) ? 1u : 0u);
            tc_mma_tf32(dcol, dah + adv, dbl + adv, idesc, 1u);
            tc_mma_tf32(dcol, dah + adv, dbh + adv, idesc, 1u);
          }
          tc_commit(&ops_empty[ob]);                   // operand buffer free once these MMAs have completed
          if (kb == NKB - 1) tc_commit(&tm_full[tb]);  // accumulator tile complete
        }
      }
    }
  } else {
    // =================== epilogue warps: TMEM -> registers (thread = row) -> warp-private staging -> global =================
    // The accumulator arrives one row per thread; global traffic goes through a [32 x 16] staging block per warp so that
    // every load / store instruction covers two 64-byte row segments (lanes 0-15 one row, lanes 16-31 the next).
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    float* stg = &estage[q][0][0];
    const int hr = lane >> 4, hc = lane & 15;          // staging <-> global mapping: row 2*rr + hr, column hc
    const bool selu = a.act == GNNFP_ACT_SELU;
    const float* kc = (!FWD && a.corr) ? a.corr + a.corr_col0 : nullptr;
    const float* auxsrc = FWD ? a.prev : (kc ? a.corr_x : nullptr);
    const int auxld = FWD ? a.ld_prev : a.corr_ld;
    for (int tq = 0; tq < my_tiles; ++tq) {
      const int tb = tq & 1;
      const int wrow0 = (tile0 + tq) * TC_BM + 32 * q; // first global row of this warp's 32 rows
      const bool valid = wrow0 + lane < n;
      mbar_wait_bounded(&tm_full[tb], (uint32_t)(tq >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(tb * BN);
      float sd = 0.f, sp = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {            // 16 accumulator columns per trip
        float acc[16], aux[16], old[FWD ? 1 : 16];
        tmem_ld8(taddr + c0, acc);
        tmem_ld8(taddr + c0 + 8, acc + 8);
        const bool cok = c0 + hc < a.N;
        if (auxsrc) {                                  // old values (FWD) / BN-correction inputs (backward), coalesced
          float t[16];
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int gr = wrow0 + 2 * rr + hr;
            t[rr] = (cok && gr < n) ? auxsrc[(size_t)gr * auxld + c0 + hc] : 0.f;
          }
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) stg[(2 * rr + hr) * 17 + hc] = t[rr];
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 16; ++j) aux[j] = stg[lane * 17 + j];
          __syncwarp();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) aux[j] = 0.f;
        }
        if (!FWD) {
          if (a.out_add) {                             // destination contents (accumulating blocks)
            float t[16];
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
              const int gr = wrow0 + 2 * rr + hr;
              t[rr] = (cok && gr < n) ? a.out[(size_t)gr * a.ld_out + c0 + hc] : 0.f;
            }
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) stg[(2 * rr + hr) * 17 + hc] = t[rr];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; ++j) old[j] = stg[lane * 17 + j];
            __syncwarp();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) old[j] = 0.f;
          }
        }
        tmem_ld_wait();
        if (c0 + 16 >= BN) {                           // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tm_empty[tb]);
        }
        float st[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c0 + j;
          float v = 0.f;
          if (valid && col < a.N) {
            if (FWD) {
              const float z = acc[j] + sbias[col];
              v = selu ? tc_selu(z) : act_fwd(a.act, z);
              if (a.prev) {
                const float dd = v - aux[j];
                sd = fmaf(dd, dd, sd);
                sp = fmaf(aux[j], aux[j], sp);
              }
            } else {
              v = a.colscale ? acc[j] * a.colscale[col] : acc[j];
              if (kc) v -= kc[col] + fmaf(aux[j], kc[2 * a.corr_in + col], kc[3 * a.corr_in + col]) * kc[a.corr_in + col];
              v += old[j];
            }
          }
          st[j] = v;
          stg[lane * 17 + j] = v;
        }
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {              // coalesced stores: two 64-byte row segments per instruction
          const int gr = wrow0 + 2 * rr + hr;
          if (cok && gr < n) a.out[(size_t)gr * a.ld_out + c0 + hc] = stg[(2 * rr + hr) * 17 + hc];
        }
        __syncwarp();
        if (FWD && a.ost_sum) {                        // column statistics of these 16 columns over the warp's 32 rows
          float x[32];
#pragma unroll
          for (int j = 0; j < 16; ++j) { x[j] = st[j]; x[16 + j] = st[j] * st[j]; }
          const float tot = warp_colsum32(x, lane);    // lane l < 16: sum of column c0+l; lane l >= 16: sum of squares of c0+l-16
          const int col = c0 + (lane & 15);
          if (col < a.N) colacc[q][lane >> 4][col] += (double)tot;
        }
      }
      if (FWD && a.prev && valid && sqrtf(sd) > a.thr * sqrtf(sp)) notconv = 1;
    }
  }
  if (FWD && a.flag_next) {
    const int any = __syncthreads_or(notconv);
    if (tid == 0 && any) atomicOr(a.flag_next, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (FWD && a.ost_sum) {
    for (int j = tid; j < a.N; j += TC_THREADS) {
      atomicAdd(a.ost_sum + j, colacc[0][0][j] + colacc[1][0][j] + colacc[2][0][j] + colacc[3][0][j]);
      atomicAdd(a.ost_sq + j, colacc[0][1][j] + colacc[1][1][j] + colacc[2][1][j] + colacc[3][1][j]);
    }
  }
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
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
  gemm_rows_tc_kernel<BN, FWD><<<grid, TC_THREADS, smem, s>>>(a);
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
