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
#include <type_traits>

#include "tile.cuh"
#include "tc.cuh"

#include "gemm.h"

#define TC_BM 128
#define TC_BK 32
#define TC_STAGES 3                  // minimum depth of the operand ring: [hi (raw fp32, filled by cp.async) | lo] x 16 KB per stage
#define TC_MAX_STAGES 6              // GNNFP_TC_STAGES=<n> deepens the ring while shared memory allows; measured SLOWER on B200 (13.7 vs 13.0
                                     // ms/step at 6 stages: the converter warps stall issuing the extra cp.async), so the default stays 3
#define TC_TILE_BYTES (TC_BM * 128)  // one A operand tile: 128 rows x 128 B


#define TC_THREADS 416                                  // warps 0-3 converters, 4-11 epilogue, 12 MMA issuer
#define TC_EPI_WARPS 8

template <int BN, bool FWD, int NB>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_rows_tc_kernel(const __grid_constant__ GemmRowsArgs a, const int NST) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; offset arithmetic (not an integer round trip) keeps the pointers in the
  // shared address space for the compiler (LDS/STS instead of generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NKB = (a.Kpad + TC_BK - 1) / TC_BK;
  const int wtile = BN * 128;                          // bytes of one [BN x 32] weight tile
  // resident weights: [NB output blocks][hi | lo][NKB][BN x 32]; NB == 2: a second output block computed from the same
  // A operand (backward dX of two equally wide pieces, e.g. dOwn and dAgg: dz is loaded and split once)
  uint8_t* Whi = base;
  uint8_t* Wlo = Whi + (size_t)NKB * wtile;
  uint8_t* Aop = base + (size_t)NB * 2 * NKB * wtile;  // [NST][hi, lo][TC_TILE_BYTES]
  __shared__ __align__(8) uint64_t ops_full[TC_MAX_STAGES], ops_empty[TC_MAX_STAGES], tm_full[2], tm_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[BN];
  __shared__ double colacc[FWD ? TC_EPI_WARPS : 1][2][BN];   // per epilogue warp: column sums / sums of squares (forward only)
  __shared__ float estage[TC_EPI_WARPS][32][17];       // per epilogue warp: [32 rows x 16 columns] transpose staging
  __shared__ float rowpart[2][TC_BM];                  // convergence sums of the second warp of each lane quarter
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_rows;
  const int n_tiles = (n + TC_BM - 1) / TC_BM;
  const int tiles_per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int my_tiles = max(0, min(n_tiles, tile0 + tiles_per_cta) - tile0);
  constexpr uint32_t TMEM_NEED = 2 * NB * BN;          // double-buffered accumulators of every output block
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));

  if (warp == 0) { tmem_alloc(&tmem_base_s, TMEM_COLS); tmem_relinquish(); }      // whole warp, converged (.sync.aligned)
  if (tid == 32) {
    for (int i = 0; i < NST; ++i) { mbar_init(&ops_full[i], 128); mbar_init(&ops_empty[i], 1); }
    mbar_init(&tm_full[0], 1); mbar_init(&tm_full[1], 1);
    mbar_init(&tm_empty[0], 32 * TC_EPI_WARPS); mbar_init(&tm_empty[1], 32 * TC_EPI_WARPS);
  }
  // ---- resident weights: split into TF32 hi / lo, K-major swizzled tiles per K block (all threads) ------------------
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float* Wsrc = b == 0 ? a.Wp : a.Wp2;
    uint8_t* whi = Whi + (size_t)b * 2 * NKB * wtile;
    uint8_t* wlo = Wlo + (size_t)b * 2 * NKB * wtile;
    for (int e = tid; e < NKB * TC_BK * BN; e += TC_THREADS) {
      const int k = e / BN, nn = e - k * BN;
      const float w = k < a.Kpad ? Wsrc[(size_t)k * a.ldw + nn] : 0.f;
      const int off = (k >> 5) * wtile + tc_sw128_off(nn, k & 31);
      *reinterpret_cast<uint32_t*>(whi + off) = __float_as_uint(w);
      *reinterpret_cast<uint32_t*>(wlo + off) = tc_lo(__float_as_uint(w));
    }
  }
  for (int j = tid; j < BN; j += TC_THREADS) sbias[j] = (FWD && a.bias && j < a.N) ? a.bias[j] : 0.f;
  for (int j = tid; j < (FWD ? TC_EPI_WARPS : 1) * 2 * BN; j += TC_THREADS) (&colacc[0][0][0])[j] = 0.0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  int notconv = 0;

  if (warp < 4) {
    // =================== converter warps: cp.async straight into the swizzled hi tile, then the lo tile ================
    // thread copies the 8-byte column pair pq of rows ty + 8*i (i < 16) of a stage
    const int pq = tid & 15, ty = tid >> 4;
    const int total = my_tiles * NKB;
    int i_tq = 0, i_kb = 0, i_idx = 0, i_slot = 0;
    uint32_t i_use = 0;                                // ring position of the next stage to issue: slot, and how often it was used
    auto issue = [&]() {
      if (i_idx < total) {
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
        const int slot = i_slot;
        const uint32_t use = i_use;
        if (++i_slot == NST) { i_slot = 0; ++i_use; }
        if (use > 0) {                                 // the MMAs that read this slot's previous contents must have completed
          mbar_wait_bounded(&ops_empty[slot], (use - 1) & 1);
          tc_fence_after();
        }
        uint8_t* hi = Aop + (size_t)slot * 2 * TC_TILE_BYTES;
        // swizzled destination of (row ty + 8*i, pair pq): row & 7 == ty for every i
        uint8_t* dst = hi + ty * 128 + ((((pq >> 1) ^ ty) & 7) << 4) + ((pq & 1) << 3);
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
          for (int i = 0; i < 16; ++i) tc_cp_async8(dst + i * 1024, pptr, 0);
        } else if (al8 && (tile0 + i_tq + 1) * TC_BM <= n) {
          const float* src = pptr + (size_t)row0 * pld + kk;
          const size_t step = (size_t)8 * pld;
          const int bytes = nv > 1 ? 8 : 4;
#pragma unroll
          for (int i = 0; i < 16; ++i) { tc_cp_async8(dst + i * 1024, src, bytes); src += step; }
        } else {
#pragma unroll 2
          for (int i = 0; i < 16; ++i) {
            const int grow = row0 + 8 * i;
            const bool valid = grow < n;
            const float* src = valid ? pptr + (size_t)grow * pld + kk : pptr;
            if (al8) {
              tc_cp_async8(dst + i * 1024, src, valid ? (nv > 1 ? 8 : 4) : 0);
            } else {
              tc_cp_async4(dst + i * 1024, src, valid ? 4 : 0);
              tc_cp_async4(dst + i * 1024 + 4, (valid && nv > 1) ? src + 1 : pptr, (valid && nv > 1) ? 4 : 0);
            }
          }
        }
        if (++i_kb == NKB) { i_kb = 0; ++i_tq; }
        ++i_idx;
      }
      tc_cp_commit();
    };
    for (int i = 0; i < NST - 1; ++i) issue();
    int c_slot = 0;
    for (int sidx = 0; sidx < total; ++sidx) {
      switch (NST) {                                   // all but the NST - 2 newest groups have landed
        case 3: tc_cp_wait<1>(); break;
        case 4: tc_cp_wait<2>(); break;
        case 5: tc_cp_wait<3>(); break;
        default: tc_cp_wait<4>(); break;
      }
      named_bar_sync(1, 128);                          // stage sidx has landed for all converter threads
      const int slot = c_slot;
      if (++c_slot == NST) c_slot = 0;
      uint8_t* ahi = Aop + (size_t)slot * 2 * TC_TILE_BYTES;
      uint8_t* alo = ahi + TC_TILE_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) {                    // lo tile: same swizzled positions, 16 bytes per item
        const int off = (tid + 128 * j) << 4;
        const uint4 v = *reinterpret_cast<const uint4*>(ahi + off);
        uint4 l;
        l.x = tc_lo(v.x); l.y = tc_lo(v.y); l.z = tc_lo(v.z); l.w = tc_lo(v.w);
        *reinterpret_cast<uint4*>(alo + off) = l;
      }
      fence_proxy_async();                             // cp.async / st.shared writes -> visible to the tensor core (async proxy)
      mbar_arrive(&ops_full[slot]);
      issue();                                         // refill the slot of stage sidx-1 once its MMAs have completed
    }
    tc_cp_wait<0>();
  } else if (warp == 4 + TC_EPI_WARPS) {
    // =================== MMA issuer: one thread drives the tensor core ====================================================
    {   // warp-uniform loop, one elected lane issues (a lane-0 branch wraps every tcgen05.mma in an R2UR waterfall loop)
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t whi_addr = smem_u32(Whi), wlo_addr = smem_u32(Wlo), aop_addr = smem_u32(Aop);
      int ob = 0;
      uint32_t oph = 0;                                // ring slot of the next stage and its mbarrier phase
      for (int tq = 0; tq < my_tiles; ++tq) {
        const int tb = tq & 1;
        if (tq >= 2) {                                 // the epilogue of the tile that used this accumulator must have drained it
          mbar_wait_bounded(&tm_empty[tb], (uint32_t)((tq >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t dcol0 = tmem_d + (uint32_t)(tb * NB * BN);
        for (int kb = 0; kb < NKB; ++kb) {
          mbar_wait_bounded(&ops_full[ob], oph);
          tc_fence_after();
          const uint64_t dah = tc_desc_sw128(aop_addr + ob * 2 * TC_TILE_BYTES);
          const uint64_t dal = tc_desc_sw128(aop_addr + ob * 2 * TC_TILE_BYTES + TC_TILE_BYTES);
          if (tc_elect()) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const uint32_t dcol = dcol0 + (uint32_t)(b * BN);
            const uint64_t dbh = tc_desc_sw128(whi_addr + (b * 2 * NKB + kb) * wtile);
            const uint64_t dbl = tc_desc_sw128(wlo_addr + (b * 2 * NKB + kb) * wtile);
#pragma unroll
            for (int k8 = 0; k8 < TC_BK / 8; ++k8) {   // 8 fp32 = 32 bytes = 2 descriptor units along K inside the swizzle atom
              const uint64_t adv = (uint64_t)(2 * k8);
              tc_mma_tf32(dcol, dal + adv, dbh + adv, idesc, (kb | k8) ? 1u : 0u);
              tc_mma_tf32(dcol, dah + adv, dbl + adv, idesc, 1u);
              tc_mma_tf32(dcol, dah + adv, dbh + adv, idesc, 1u);
            }
          }
          tc_commit(&ops_empty[ob]);                   // operand buffer free once these MMAs have completed
          if (kb == NKB - 1) tc_commit(&tm_full[tb]);  // accumulator tile complete
          }
          __syncwarp();
          if (++ob == NST) { ob = 0; oph ^= 1u; }
        }
      }
    }
  } else {
    // =================== epilogue warps: TMEM -> registers (thread = row) -> warp-private staging -> global =================
    // The accumulator arrives one row per thread; global traffic goes through a [32 x 16] staging block per warp so that
    // every load / store instruction covers two 64-byte row segments (lanes 0-15 one row, lanes 16-31 the next).
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int ew = warp - 4;                           // epilogue warp index; warps ew and ew+4 share a lane quarter and
    const int chalf = ew >> 2;                         // take alternate 16-column chunks
    float* stg = &estage[ew][0][0];
    const int hr = lane >> 4, hc = lane & 15;          // staging <-> global mapping: row 2*rr + hr, column hc
    const bool selu = a.act == GNNFP_ACT_SELU;
    for (int tq = 0; tq < my_tiles; ++tq) {
      const int tb = tq & 1;
      const int wrow0 = (tile0 + tq) * TC_BM + 32 * q; // first global row of this warp's 32 rows
      const bool valid = wrow0 + lane < n;
      mbar_wait_bounded(&tm_full[tb], (uint32_t)(tq >> 1) & 1u);
      tc_fence_after();
      float sdp[16], spp[16];                          // convergence partial sums of rows 2*rr + hr over this lane's columns
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) { sdp[rr] = 0.f; spp[rr] = 0.f; }
      bool handed_back = false;
#ifdef TC_DEBUG_SKIP_EPI
      { const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(tb * NB * BN); float acc[8]; tmem_ld8(taddr, acc); tmem_ld_wait(); tc_fence_before(); mbar_arrive(&tm_empty[tb]); if (acc[0] == 123.456f) notconv = 1; continue; }
#endif
      const bool rows_full = wrow0 + 32 <= n;
#pragma unroll
      for (int b = 0; b < NB; ++b) {                   // output blocks of this row tile (accumulators tb*NB + b)
        const bool second = NB == 2 && b == 1;
        const bool last_blk = b == NB - 1;
        float* const e_out = second ? a.out2 : a.out;
        const int e_ldo = second ? a.ld_out2 : a.ld_out;
        const int e_add = second ? a.out_add2 : a.out_add;
        const float* const e_cs = second ? a.colscale2 : a.colscale;
        const float* const kc = (!FWD && a.corr) ? a.corr + (second ? a.corr_col02 : a.corr_col0) : nullptr;
        const float* const auxsrc = FWD ? a.prev : (kc ? (second ? a.corr_x2 : a.corr_x) : nullptr);
        const int auxld = FWD ? a.ld_prev : (second ? a.corr_ld2 : a.corr_ld);
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)((tb * NB + b) * BN);
        // one 16-column chunk; FULL = all 32 rows of the warp and all 16 columns are inside the matrix (warp uniform), which
        // removes every per-element predicate from the common case
        auto chunk = [&](int c0, auto full_tag) {
          constexpr bool FULL = decltype(full_tag)::value;
          float acc[16];
          tmem_ld8(taddr + c0, acc);
          tmem_ld8(taddr + c0 + 8, acc + 8);
          // coalesced side inputs of this chunk, issued before the accumulator is consumed: lane (hr, hc) owns rows
          // 2*rr + hr, column c0 + hc (two 64-byte row segments per instruction)
          const bool cok = FULL || c0 + hc < a.N;
          const int nvr = FULL ? 32 : (cok ? n - wrow0 - hr : 0);      // rows 2*rr + hr with 2*rr < nvr are inside the matrix
          float t[16], o[FWD ? 1 : 16];
          if (auxsrc) {
            const float* ap = auxsrc + (size_t)(wrow0 + hr) * auxld + c0 + hc;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) { t[rr] = (FULL || 2 * rr < nvr) ? *ap : 0.f; ap += 2 * auxld; }
          } else {
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) t[rr] = 0.f;
          }
          if (!FWD) {
            if (e_add) {
              const float* op = e_out + (size_t)(wrow0 + hr) * e_ldo + c0 + hc;
#pragma unroll
              for (int rr = 0; rr < 16; ++rr) { o[rr] = (FULL || 2 * rr < nvr) ? *op : 0.f; op += 2 * e_ldo; }
            } else {
#pragma unroll
              for (int rr = 0; rr < 16; ++rr) o[rr] = 0.f;
            }
          }
          tmem_ld_wait();
          if (last_blk && c0 + 16 * (TC_EPI_WARPS / 4) >= BN) {      // this thread's last read of the accumulator: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tm_empty[tb]);
            handed_back = true;
          }
          // thread = row: bias + activation (forward) / column scale (backward)
          float v[16];
          if (FWD) {
            if (selu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = tc_selu(acc[j] + sbias[c0 + j]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = act_fwd(a.act, acc[j] + sbias[c0 + j]);
            }
          } else if (e_cs) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = acc[j] * ((FULL || c0 + j < a.N) ? e_cs[c0 + j] : 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = acc[j];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) stg[lane * 17 + j] = (FULL || (valid && c0 + j < a.N)) ? v[j] : 0.f;
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
          float k0 = 0.f, k1 = 0.f, kA = 0.f, kB = 0.f;
          if (!FWD && kc && cok) { k0 = kc[c0 + hc]; k1 = kc[a.corr_in + c0 + hc]; kA = kc[2 * a.corr_in + c0 + hc]; kB = kc[3 * a.corr_in + c0 + hc]; }
          float* outp = e_out + (size_t)(wrow0 + hr) * e_ldo + c0 + hc;
#pragma unroll
          for (int rr = 0; rr < 16; ++rr, outp += 2 * e_ldo) {   // coalesced: stores, convergence sums, statistics, corrections
            float x = stg[(2 * rr + hr) * 17 + hc];
            if (FULL || 2 * rr < nvr) {
              if (FWD) {
                const float dd = x - t[rr];
                sdp[rr] = fmaf(dd, dd, sdp[rr]);
                spp[rr] = fmaf(t[rr], t[rr], spp[rr]);
                s1 += x;
                s2 = fmaf(x, x, s2);
              } else {
                if (kc) x -= k0 + fmaf(t[rr], kA, kB) * k1;
                x += o[rr];
              }
              *outp = x;
            }
          }
          __syncwarp();
          if (FWD && a.ost_sum) {                        // column statistics: fold the two row halves, lanes 0-15 own a column
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
            if (hr == 0 && cok) { colacc[FWD ? ew : 0][0][c0 + hc] += (double)s1; colacc[FWD ? ew : 0][1][c0 + hc] += (double)s2; }
          }
        };
#pragma unroll 1
        for (int c0 = 16 * chalf; c0 < BN; c0 += 16 * (TC_EPI_WARPS / 4)) {      // 16 accumulator columns per trip
          if (rows_full && c0 + 16 <= a.N) chunk(c0, std::true_type{});
          else chunk(c0, std::false_type{});
        }
      }
      float sd = 0.f, sp = 0.f;                        // row sums: reduce over the 16 column lanes; lane rr of each half keeps row 2*rr + hr
      if (FWD && a.prev) {
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
          float x = sdp[rr], y = spp[rr];
#pragma unroll
          for (int of = 8; of > 0; of >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, of);
            y += __shfl_xor_sync(0xffffffffu, y, of);
          }
          if (hc == rr) { sd = x; sp = y; }
        }
      }
      const int myrow = 2 * hc + hr;                   // the row (within the warp's 32) whose sums this lane holds
      const bool myvalid = wrow0 + myrow < n;
      if (!handed_back) { tc_fence_before(); mbar_arrive(&tm_empty[tb]); }     // narrow tiles: this warp had no chunk
      if (FWD && a.prev) {                             // the row's sums are split over the two warps of its lane quarter
        if (chalf == 1) { rowpart[0][32 * q + myrow] = sd; rowpart[1][32 * q + myrow] = sp; }
        named_bar_sync(2, 32 * TC_EPI_WARPS);
        if (chalf == 0) {
          sd += rowpart[0][32 * q + myrow];
          sp += rowpart[1][32 * q + myrow];
          if (myvalid && sqrtf(sd) > a.thr * sqrtf(sp)) notconv = 1;
        }
        named_bar_sync(2, 32 * TC_EPI_WARPS);          // rowpart may be overwritten by the next tile
      }
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
      double s1 = 0.0, s2 = 0.0;
      for (int w = 0; w < TC_EPI_WARPS; ++w) { s1 += colacc[w][0][j]; s2 += colacc[w][1][j]; }
      atomicAdd(a.ost_sum + j, s1);
      atomicAdd(a.ost_sq + j, s2);
    }
  }
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
template <int BN, bool FWD, int NB>
static int launch_tc_t(const GemmRowsArgs& a, cudaStream_t s, int prof_cat) {
  const int NKB = (a.Kpad + TC_BK - 1) / TC_BK;
  static size_t smem_cap = 0;                          // dynamic shared memory this instantiation may use next to its static arrays
  if (!smem_cap) {
    cudaFuncAttributes fa;
    int dev = 0, optin = 0;
    GNNFP_CHECK_CUDA(cudaGetDevice(&dev));
    GNNFP_CHECK_CUDA(cudaFuncGetAttributes(&fa, gemm_rows_tc_kernel<BN, FWD, NB>));
    GNNFP_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    smem_cap = (size_t)optin - fa.sharedSizeBytes;
  }
  int nst = TC_STAGES;
  static const int nst_max = getenv("GNNFP_TC_STAGES") ? atoi(getenv("GNNFP_TC_STAGES")) : TC_STAGES;
  while (nst < nst_max && nst < TC_MAX_STAGES &&
         (size_t)NB * 2 * NKB * BN * 128 + (size_t)(nst + 1) * 2 * TC_TILE_BYTES + 1024 <= smem_cap) ++nst;
  const size_t smem = (size_t)NB * 2 * NKB * BN * 128 + (size_t)nst * 2 * TC_TILE_BYTES + 1024;
  if (smem > smem_cap) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows_tc: %zu bytes of shared memory needed, %zu available", smem, smem_cap);
  static size_t attr = 0;
  if (smem > attr) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(gemm_rows_tc_kernel<BN, FWD, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_tiles = (a.n_rows + TC_BM - 1) / TC_BM;
  const int nsm = gnnfp_num_sms();
  const int grid = n_tiles < nsm ? n_tiles : nsm;
  ProfScope ps(prof_cat, s);
  gemm_rows_tc_kernel<BN, FWD, NB><<<grid, TC_THREADS, smem, s>>>(a, nst);
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
  const int NB = a.nblk == 2 ? 2 : 1;
  if (NB == 2 && (a.fwd || BN > 64)) return 0;         // two output blocks: backward only; 80 columns do not fit next to the ring
  const size_t smem = (size_t)NB * 2 * NKB * BN * 128 + (size_t)TC_STAGES * 2 * TC_TILE_BYTES + 1024;
  return smem <= (NB == 2 ? 200 : 220) * 1024;
}

int launch_gemm_rows_tc(const GemmRowsArgs& a, cudaStream_t s, int prof_cat) {
  if (a.n_rows <= 0) return GNNFP_OK;
  if (!gemm_rows_tc_supported(a)) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows_tc: shape N=%d K=%d not supported", a.N, a.Kpad);
  const bool fwd = a.fwd != 0;
  if (a.nblk == 2) {
    switch (a.ldw) {
      case 16: return launch_tc_t<16, false, 2>(a, s, prof_cat);
      case 32: return launch_tc_t<32, false, 2>(a, s, prof_cat);
      case 48: return launch_tc_t<48, false, 2>(a, s, prof_cat);
      case 64: return launch_tc_t<64, false, 2>(a, s, prof_cat);
      default: GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows_tc: two output blocks of %d columns", a.ldw);
    }
  }
  switch (a.ldw) {
    case 16: return fwd ? launch_tc_t<16, true, 1>(a, s, prof_cat) : launch_tc_t<16, false, 1>(a, s, prof_cat);
    case 32: return fwd ? launch_tc_t<32, true, 1>(a, s, prof_cat) : launch_tc_t<32, false, 1>(a, s, prof_cat);
    case 48: return fwd ? launch_tc_t<48, true, 1>(a, s, prof_cat) : launch_tc_t<48, false, 1>(a, s, prof_cat);
    case 64: return fwd ? launch_tc_t<64, true, 1>(a, s, prof_cat) : launch_tc_t<64, false, 1>(a, s, prof_cat);
    case 80: return fwd ? launch_tc_t<80, true, 1>(a, s, prof_cat) : launch_tc_t<80, false, 1>(a, s, prof_cat);
    default: GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "gemm_rows_tc: %d output columns", a.ldw);
  }
}

// ------------------------------------------------------------------------------------------------------------
// dW = X^T dz on the tensor cores.  The reduction runs over the rows, so both operands are needed "transposed":
//     D[j, c] (+)= sum_r dz[r, j] * X[r, c]        A = dz^T [128 (H valid) x 32 rows],  B = X^T [NT x 32 rows], K-major
// with NT = 16*ceil((Kp + 1)/16) <= 256 output columns: the Kp input columns, one column of ones (so that D[j, Kp] = db_j
// comes out of the same MMAs) and zero padding.  Per persistent CTA (one per SM, 160 threads):
//   * 4 converter warps: 8-byte cp.async of 32 rows of all pieces and of dz into a raw stage (row-major, leading dimension
//     = 2 mod 4 floats so that the 8-byte column reads below are conflict free), then the transposing split: lane r reads
//     raw[r][c..c+1] and writes 4-byte hi / lo words into row c of the swizzled K-major operand tile (one full 128-byte
//     operand row per warp store: conflict free);
//   * 1 MMA warp: 12 tcgen05.mma (M = 128, N = NT, K = 8, 3xTF32) per stage, accumulating over ALL rows of the CTA in TMEM;
//   * no per-tile epilogue: after the last stage the accumulator is staged through shared memory once and flushed with
//     the same BN algebra as gemm_dw_kernel (per-CTA partial slot, P_c / Q_c sums).
#define DW_TC_CONV_WARPS 8
#define DW_TC_CONV_THREADS (32 * DW_TC_CONV_WARPS)
#define DW_TC_THREADS (DW_TC_CONV_THREADS + 32)
#define DW_TC_ROWS 32

__global__ void __launch_bounds__(DW_TC_THREADS, 1) gemm_dw_tc_kernel(const __grid_constant__ GemmDwArgs a, int NT, int RAWLD) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int btile = NT * 128;                          // bytes of one X^T operand tile
  const int slot_bytes = 2 * TC_TILE_BYTES + 2 * btile;       // [A_hi | A_lo | B_hi | B_lo]
  uint8_t* ops = base;                                 // [2 slots]
  float* raw = reinterpret_cast<float*>(ops + 2 * (size_t)slot_bytes);      // [2][DW_TC_ROWS][RAWLD]
  __shared__ __align__(8) uint64_t ops_full[2], ops_empty[2], done_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_gam[256], s_bet[256], s_bA[256], s_bB[256];
  __shared__ int s_kp[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_rows;
  const int n_chunks_all = (n + DW_TC_ROWS - 1) / DW_TC_ROWS;
  const int per_cta = (n_chunks_all + gridDim.x - 1) / gridDim.x;
  const int c0 = blockIdx.x * per_cta;
  const int total = max(0, min(n_chunks_all, c0 + per_cta) - c0);
  const int Kp = a.Kp, H = a.H;
  const int xpairs = Kp >> 1, zpairs = (H + 1) >> 1, npairs = xpairs + zpairs;   // raw row: [X pieces (Kp) | dz (H)]
  int K = 0;
  for (int p = 0; p < a.n_pieces; ++p) K += a.p[p].width;

  if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
  if (tid == 32) {
    mbar_init(&ops_full[0], DW_TC_CONV_THREADS); mbar_init(&ops_full[1], DW_TC_CONV_THREADS);
    mbar_init(&ops_empty[0], 1); mbar_init(&ops_empty[1], 1);
    mbar_init(&done_bar, 1);
  }
  for (int c = tid; c < K; c += DW_TC_THREADS) {       // per-column constants of the flush
    int kp = 0, coff = 0;
    for (int p = 0; p < a.n_pieces; ++p) {
      if (c >= coff && c < coff + a.p[p].width) kp = a.p[p].k8 + (c - coff);
      coff += a.p[p].width;
    }
    s_kp[c] = kp;
    s_gam[c] = a.bnA ? a.gamma[c] : 1.f; s_bet[c] = a.bnA ? a.beta[c] : 0.f;
    s_bA[c] = a.bnA ? a.bnA[c] : 1.f; s_bB[c] = a.bnA ? a.bnB[c] : 0.f;
  }
  // zero the operand slots once (padding rows of both operands are never written afterwards), then the row of ones
  for (int e = tid; e < 2 * slot_bytes / 16; e += DW_TC_THREADS) reinterpret_cast<uint4*>(ops)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (tid < 64) {
    const int sl = tid >> 5, r = tid & 31;
    *reinterpret_cast<float*>(ops + (size_t)sl * slot_bytes + 2 * TC_TILE_BYTES + tc_sw128_off(Kp, r)) = 1.0f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp < DW_TC_CONV_WARPS) {
    // ---- per-thread description of the (up to 4) column pairs this lane copies: pair lane + 32*q of every row -------
    const float* pbase[4];
    int pld[4], pbytes[4];
    bool pal8[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kcol = 2 * (lane + 32 * q);
      pbase[q] = a.dz; pld[q] = 0; pbytes[q] = -1;     // -1: pair outside the row
      pal8[q] = true;
      if (kcol < Kp) {
        int p = 0;
        while (p + 1 < a.n_pieces && kcol >= a.p[p + 1].k8) ++p;
        const int kk = kcol - a.p[p].k8, nv = a.p[p].width - kk;
        pbase[q] = a.p[p].ptr + (nv > 0 ? kk : 0); pld[q] = a.p[p].ld;
        pbytes[q] = nv <= 0 ? 0 : (nv > 1 ? 8 : 4);
        pal8[q] = a.p[p].al8 != 0;
      } else if (kcol < Kp + H) {
        const int j0 = kcol - Kp, nv = H - j0;
        pbase[q] = a.dz + j0; pld[q] = a.ld_dz; pbytes[q] = nv > 1 ? 8 : 4;
      } else if (kcol < 2 * npairs) {
        pbytes[q] = 0;
      }
    }
    auto issue = [&](int it) {
      if (it < total) {
        float* dst = raw + (it & 1) * DW_TC_ROWS * RAWLD;
        const int row0 = (c0 + it) * DW_TC_ROWS;
        if (it + 2 < total) {                          // L2 prefetch two stages ahead: one line per lane and row segment
          const int prow = (c0 + it + 2) * DW_TC_ROWS + lane;
          if (prow < n && warp < a.n_pieces) {
            const char* b0 = reinterpret_cast<const char*>(a.p[warp].ptr + (size_t)prow * a.p[warp].ld);
            const char* l1 = b0 + (size_t)a.p[warp].width * 4;
            for (const char* l = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(b0) & ~(uintptr_t)127); l < l1; l += 128)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(l));
          }
          if (prow < n && warp == DW_TC_CONV_WARPS - 1) {
            const char* b0 = reinterpret_cast<const char*>(a.dz + (size_t)prow * a.ld_dz);
            const char* l1 = b0 + (size_t)H * 4;
            for (const char* l = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(b0) & ~(uintptr_t)127); l < l1; l += 128)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(l));
          }
        }
        if (row0 + DW_TC_ROWS <= n) {                  // full stage: no row predicates, strength-reduced addresses
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (pbytes[q] < 0) continue;
            float* d = dst + warp * RAWLD + 2 * (lane + 32 * q);
            if (pbytes[q] == 0) {
#pragma unroll
              for (int i = 0; i < DW_TC_ROWS / DW_TC_CONV_WARPS; ++i) tc_cp_async8(d + i * DW_TC_CONV_WARPS * RAWLD, a.dz, 0);
            } else {
              const float* src = pbase[q] + (size_t)(row0 + warp) * pld[q];
              const size_t step = (size_t)DW_TC_CONV_WARPS * pld[q];
              if (pal8[q]) {
#pragma unroll
                for (int i = 0; i < DW_TC_ROWS / DW_TC_CONV_WARPS; ++i) { tc_cp_async8(d + i * DW_TC_CONV_WARPS * RAWLD, src, pbytes[q]); src += step; }
              } else {
#pragma unroll
                for (int i = 0; i < DW_TC_ROWS / DW_TC_CONV_WARPS; ++i) {
                  tc_cp_async4(d + i * DW_TC_CONV_WARPS * RAWLD, src, 4);
                  tc_cp_async4(d + i * DW_TC_CONV_WARPS * RAWLD + 1, pbytes[q] == 8 ? src + 1 : a.dz, pbytes[q] == 8 ? 4 : 0);
                  src += step;
                }
              }
            }
          }
        } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (pbytes[q] < 0) continue;
#pragma unroll
          for (int i = 0; i < DW_TC_ROWS / DW_TC_CONV_WARPS; ++i) {       // warp w copies rows w, w+8, ... of the stage
            const int r = warp + DW_TC_CONV_WARPS * i;
            const int grow = row0 + r;
            const bool ok = grow < n && pbytes[q] > 0;
            float* d = dst + r * RAWLD + 2 * (lane + 32 * q);
            const float* src = ok ? pbase[q] + (size_t)grow * pld[q] : a.dz;
            if (pal8[q]) tc_cp_async8(d, src, ok ? pbytes[q] : 0);
            else {                                     // piece whose rows are not 8-byte aligned: two 4-byte copies
              tc_cp_async4(d, src, ok ? 4 : 0);
              tc_cp_async4(d + 1, (ok && pbytes[q] == 8) ? src + 1 : a.dz, (ok && pbytes[q] == 8) ? 4 : 0);
            }
          }
        }
        }
      }
      tc_cp_commit();
    };
    issue(0);
    const int koff = ((lane >> 2) << 4) + ((lane & 3) << 2);            // unswizzled byte offset of stage row `lane` inside an operand row
    for (int it = 0; it < total; ++it) {
      tc_cp_wait<0>();
      named_bar_sync(1, DW_TC_CONV_THREADS);           // stage `it` landed; everyone is done reading the other raw slot
      issue(it + 1);
      const int slot = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      if (use > 0) { mbar_wait_bounded(&ops_empty[slot], (use - 1) & 1); tc_fence_after(); }
      uint8_t* Ahi = ops + (size_t)slot * slot_bytes;
      uint8_t* Bhi = Ahi + 2 * TC_TILE_BYTES;
      const float* rs = raw + slot * DW_TC_ROWS * RAWLD + lane * RAWLD;
      // transposing split: lane = row of the stage; one warp store writes a whole 128-byte operand row (conflict free)
      // pair pp = warp + 8k -> operand rows 2*warp + 16k (+1): the row's position inside its 8-row group is fixed per warp,
      // only the group advances (2 groups = 2048 bytes per k)
      {
        const int ro = (2 * warp) & 7;
        const int off0 = ((2 * warp) >> 3) * 1024 + ro * 128 + (koff ^ (ro << 4));
        const int off1 = ((2 * warp) >> 3) * 1024 + (ro + 1) * 128 + (koff ^ ((ro + 1) << 4));
        uint8_t* h0 = Bhi + off0;
        uint8_t* h1 = Bhi + off1;
        const float* rp = rs + 2 * warp;
#pragma unroll 4
        for (int pp = warp; pp < xpairs; pp += DW_TC_CONV_WARPS) {       // X columns 2pp, 2pp+1 -> rows of B = X^T
          const float2 v = *reinterpret_cast<const float2*>(rp);
          *reinterpret_cast<uint32_t*>(h0) = __float_as_uint(v.x);
          *reinterpret_cast<uint32_t*>(h0 + btile) = tc_lo(__float_as_uint(v.x));
          *reinterpret_cast<uint32_t*>(h1) = __float_as_uint(v.y);
          *reinterpret_cast<uint32_t*>(h1 + btile) = tc_lo(__float_as_uint(v.y));
          h0 += 2048; h1 += 2048; rp += 2 * DW_TC_CONV_WARPS;
        }
        uint8_t* g0 = Ahi + off0;
        uint8_t* g1 = Ahi + off1;
        rp = rs + Kp + 2 * warp;
#pragma unroll 2
        for (int pp = warp; pp < zpairs; pp += DW_TC_CONV_WARPS) {       // dz columns 2pp, 2pp+1 -> rows of A = dz^T
          const float2 v = *reinterpret_cast<const float2*>(rp);
          *reinterpret_cast<uint32_t*>(g0) = __float_as_uint(v.x);
          *reinterpret_cast<uint32_t*>(g0 + TC_TILE_BYTES) = tc_lo(__float_as_uint(v.x));
          if (2 * pp + 1 < H) {
            *reinterpret_cast<uint32_t*>(g1) = __float_as_uint(v.y);
            *reinterpret_cast<uint32_t*>(g1 + TC_TILE_BYTES) = tc_lo(__float_as_uint(v.y));
          }
          g0 += 2048; g1 += 2048; rp += 2 * DW_TC_CONV_WARPS;
        }
      }
      fence_proxy_async();
      mbar_arrive(&ops_full[slot]);
    }
    tc_cp_wait<0>();
  } else {
    // ---- MMA issuer (warp-uniform loop, one elected lane issues) ----------------------------------------------------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t ops_addr = smem_u32(ops);
    for (int it = 0; it < total; ++it) {
      const int slot = it & 1;
      mbar_wait_bounded(&ops_full[slot], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const uint32_t sa = ops_addr + slot * slot_bytes;
      const uint64_t dah = tc_desc_sw128(sa), dal = tc_desc_sw128(sa + TC_TILE_BYTES);
      const uint64_t dbh = tc_desc_sw128(sa + 2 * TC_TILE_BYTES), dbl = tc_desc_sw128(sa + 2 * TC_TILE_BYTES + btile);
      if (tc_elect()) {
#pragma unroll
        for (int k8 = 0; k8 < DW_TC_ROWS / 8; ++k8) {
          const uint64_t adv = (uint64_t)(2 * k8);
          tc_mma_tf32(tmem_d, dal + adv, dbh + adv, idesc, (it | k8) ? 1u : 0u);
          tc_mma_tf32(tmem_d, dah + adv, dbl + adv, idesc, 1u);
          tc_mma_tf32(tmem_d, dah + adv, dbh + adv, idesc, 1u);
        }
        tc_commit(&ops_empty[slot]);
      }
      __syncwarp();
    }
    if (tc_elect()) tc_commit(&done_bar);
    __syncwarp();
  }
  // ---- flush: accumulator -> shared memory (sD[c][j]) -> BN algebra -> this CTA's partial slot -------------------------
  constexpr int HS = 129;                              // leading dimension of sD (rows of D = 128 TMEM lanes)
  float* sD = reinterpret_cast<float*>(ops);           // [NT][HS] <= 2 operand slots
  if (warp < DW_TC_CONV_WARPS) {
    if (total > 0) {
      mbar_wait_bounded(&done_bar, 0);
      tc_fence_after();
      const int q = warp & 3;                          // TMEM lane quarter; the two warps of a quarter take alternate column groups
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16);
      for (int cc = 8 * (warp >> 2); cc < NT; cc += 16) {
        float v[8];
        tmem_ld8(taddr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) sD[(cc + i) * HS + 32 * q + lane] = v[i];
      }
    } else {
      for (int e = tid; e < NT * 128; e += DW_TC_CONV_THREADS) sD[(e >> 7) * HS + (e & 127)] = 0.f;
    }
    tc_fence_before();
    named_bar_sync(1, DW_TC_CONV_THREADS);
    float* part = a.partial + (size_t)blockIdx.x * a.n_params;
    const float* sdb = sD + Kp * HS;                   // db_j = D[j][Kp] (the column of ones)
    const int KH = K * H;
    // one reduction per element and launch into this CTA's slot: fire-and-forget RED.ADD (no read latency); the order of the
    // additions to one address is the launch order, so the sums stay deterministic
    for (int e = tid; e < KH; e += DW_TC_CONV_THREADS) {
      const int c = e / H, j = e - c * H;
      const float x = sD[s_kp[c] * HS + j];
      atomicAdd(part + e, a.bnA ? s_gam[c] * fmaf(s_bA[c], x, s_bB[c] * sdb[j]) + s_bet[c] * sdb[j] : x);
    }
    for (int j = tid; j < H; j += DW_TC_CONV_THREADS) atomicAdd(part + a.bias_off + j, sdb[j]);
    if (a.bn_partial) {
      float* bp = a.bn_partial + (size_t)blockIdx.x * 2 * K;
      for (int c = tid; c < K; c += DW_TC_CONV_THREADS) {
        const float* wr = a.W + (size_t)c * H;
        const float* dr = sD + s_kp[c] * HS;
        float P = 0.f, Q = 0.f;
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
          const float w = wr[j];
          P = fmaf(w, sdb[j], P);
          Q = fmaf(w, dr[j], Q);
        }
        bp[c] = P;
        bp[K + c] = fmaf(s_bA[c], Q, s_bB[c] * P);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}

static void dw_tc_geometry(const GemmDwArgs& a, int* NT, int* RAWLD, size_t* smem) {
  *NT = 16 * ((a.Kp + 1 + 15) / 16);
  int ld = a.Kp + a.H + (a.H & 1);
  while (ld % 4 != 2) ++ld;                            // = 2 mod 4 floats: 8-byte column reads of 32 rows are conflict free
  *RAWLD = ld;
  const size_t slot = 2 * (size_t)TC_TILE_BYTES + 2 * (size_t)(*NT) * 128;
  const size_t sd = (size_t)(*NT) * 129 * sizeof(float);
  const size_t opsz = 2 * slot > sd ? 2 * slot : sd;
  *smem = opsz + 2 * (size_t)DW_TC_ROWS * ld * sizeof(float) + 1024;
}

int gemm_dw_tc_supported(const GemmDwArgs& a) {
  if (a.rowlist != nullptr || a.dz_compact || a.Kp % 2 != 0 || a.H > 128 || a.Kp + 1 > 256 || a.n_pieces > 3) return 0;
  if (a.Kp + a.H > 256) return 0;                      // 4 column pairs per lane cover at most 256 raw columns
  if (((uintptr_t)a.dz & 7) != 0 || a.ld_dz % 2 != 0) return 0;
  int NT, RAWLD; size_t smem;
  dw_tc_geometry(a, &NT, &RAWLD, &smem);
  return smem <= 220 * 1024;
}

int launch_gemm_dw_tc(const GemmDwArgs& a, cudaStream_t s, int prof_cat, int* grid_out) {
  if (a.n_rows <= 0) { if (grid_out) *grid_out = 0; return GNNFP_OK; }
  int NT, RAWLD; size_t smem;
  dw_tc_geometry(a, &NT, &RAWLD, &smem);
  static size_t attr = 0;
  if (smem > attr) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(gemm_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_chunks = (a.n_rows + DW_TC_ROWS - 1) / DW_TC_ROWS;
  const int nsm = gnnfp_num_sms();
  int grid = (n_chunks + 7) / 8;                       // >= 256 rows per CTA
  if (grid > nsm) grid = nsm;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = grid;
  ProfScope ps(prof_cat, s);
  gemm_dw_tc_kernel<<<grid, DW_TC_THREADS, smem, s>>>(a, NT, RAWLD);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
