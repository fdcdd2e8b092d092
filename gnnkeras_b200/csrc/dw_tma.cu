// dw_tma.cu - dW = X^T dz of a single-Dense-layer state net on the tensor cores, operands fed by TMA with NO transposing pass.
//
// The reduction of dW runs over the rows, so the MMA needs both operands with the ROW index as K:
//     D[j, c] (+)= sum_r dz[r, j] * X[r, c]        A = dz^T (M = 128, H valid),  B = X^T (N = 32 * chunks of X columns)
// A row-major [rows x 32 columns] fp32 box, as the TMA unit writes it with SWIZZLE_128B_ATOM_32B, IS the canonical "MN-major"
// operand layout tcgen05 accepts for TF32 (128-byte swizzle with 32-byte atomicity, the only MN-major layout of 32-bit
// operands: 32 contiguous M/N elements per 128-byte line, 4 K lines per swizzle atom, the 32-byte units of a line XORed with
// the line index mod 4; atoms of further M/N groups one leading-byte-offset apart) - so the raw boxes are used as a_hi / b_hi as they land (the tensor core
// reads the top 19 bits of an fp32 word), and gemm_dw_tc_kernel's transposing split (cp.async into a raw stage, 4-byte
// scatter stores into K-major tiles, one stage in flight) disappears.  Per persistent CTA (one per SM, 6 warps):
//   warp 4   TMA producer: per 32-row stage one box per 32-column chunk of X and of dz into a ring of `n_stages` stages
//   warps 0-3 converters: lo = rn_tf32(v - trunc_tf32(v)) of every element into the same swizzled position of one of two
//            "lo" slots (128-bit conflict-free accesses) + the column sums of dz (= db) in registers
//   warp 5   MMA issuer: per 8-row K step  lo.hi + hi.lo + hi.hi  (3xTF32), accumulating over ALL rows of the CTA in TMEM
//   flush    as gemm_dw_tc_kernel: accumulator -> shared memory -> BN algebra -> this CTA's slot of the partial sums
// Reference: the weight gradients of GNN.py:284-297 (tape.gradient through `convergence`, GNN.py:217-236).
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "gemm.h"
#include "rows_tma.h"
#include "tc.cuh"

// rows per stage (a.rows: 32, 64 or 128 = 4, 8 or 16 K steps): chosen by the launcher so that a stage carries up to 40 KB.
// Measured on B200 (N = 248 k rows, 148 CTAs): main loop 25 us at D = 14 and 57 us at D = 78, i.e. ~5.3 us per 32-column
// chunk (= the HBM rate) on top of ~14 us that scale with the rows only (3 dependent MMAs per 8-row K step); flush 1.5-10 us.
#define DT_CONV_THREADS 128
#define DT_THREADS 192
#define DT_MAXSTAGES 8
#define DT_MAXLO 4

__device__ __forceinline__ void dt_tma_load(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// MN-major SWIZZLE_128B_BASE32B operand (cute::UMMA::SmemDescriptor, layout type 1): atoms of 32 M/N elements x 4 K lines
// (512 bytes); the next M/N atom lies `lbo` bytes further, the next K atom 512 bytes further
__device__ __forceinline__ uint64_t dt_desc_mn(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

// phase time stamps of the last launch (CTA x [entry, prologue done, main loop done (converters), accumulator staged, exit],
// %globaltimer ns): read back by gnnfp_debug_rt_times(out, 2) (scratch/profiling only)
__device__ long long g_dt_times[160][8];
__device__ __forceinline__ long long dt_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DT_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 160) g_dt_times[blockIdx.x][i] = dt_now(); } while (0)
int dw_tma_times(long long* out) {
  GNNFP_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_dt_times, sizeof(long long) * 160 * 8));
  return GNNFP_OK;
}

__global__ void __launch_bounds__(DT_THREADS, 1) dw_tma_kernel(const __grid_constant__ DwTmaArgs a) {
  if (a.gate && *a.gate == 0) return;
  DT_STAMP(0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NXC = a.n_xc, NZC = a.n_zc, NS = a.n_stages, NLO = a.n_lo;
  const int DT_ROWS = a.rows, DT_CHUNK_BYTES = a.rows * 128;
  const int stage_bytes = (NXC + NZC) * DT_CHUNK_BYTES;
  uint8_t* hi = base;                                  // [NS] stages: [X chunks | dz chunks], raw fp32 = the hi operands
  uint8_t* lo = hi + (size_t)NS * stage_bytes;         // [NLO] slots of the same shape
  __shared__ __align__(8) uint64_t hi_full[DT_MAXSTAGES], hi_empty[DT_MAXSTAGES], lo_full[DT_MAXLO], lo_empty[DT_MAXLO], done_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_gam[256], s_bet[256], s_bA[256], s_bB[256], s_db[128];
  __shared__ float s_zred[16][128];
  __shared__ int s_kp[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_rows, H = a.H, K = a.K, NT = 32 * NXC;
  const int n_chunks_all = (n + DT_ROWS - 1) / DT_ROWS;
  const int per_cta = (n_chunks_all + gridDim.x - 1) / gridDim.x;
  const int c0 = blockIdx.x * per_cta;
  const int total = max(0, min(n_chunks_all, c0 + per_cta) - c0);

  if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
  if (tid == 32) {
    for (int s = 0; s < NS; ++s) { mbar_init(&hi_full[s], 1); mbar_init(&hi_empty[s], 1); }
    for (int b = 0; b < NLO; ++b) { mbar_init(&lo_full[b], DT_CONV_THREADS); mbar_init(&lo_empty[b], 1); }
    mbar_init(&done_bar, 1);
  }
  for (int c = tid; c < K; c += DT_THREADS) {          // per input column: its accumulator column and the constants of the flush
    int kp = 0;
    for (int p = 0; p < a.n_pieces; ++p)
      if (c >= a.p_in0[p] && c < a.p_in0[p] + a.p_w[p]) kp = a.p_acc0[p] + (c - a.p_in0[p]);
    s_kp[c] = kp;
    s_gam[c] = a.bnA ? a.gamma[c] : 1.f; s_bet[c] = a.bnA ? a.beta[c] : 0.f;
    s_bA[c] = a.bnA ? a.bnA[c] : 1.f; s_bB[c] = a.bnA ? a.bnB[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;
  DT_STAMP(1);

  if (warp == 4) {
    // ---- TMA producer ---------------------------------------------------------------------------------------------------
    for (int it = 0; it < total; ++it) {
      const int s = it % NS;
      if (it >= NS) mbar_wait_bounded(&hi_empty[s], (uint32_t)(it / NS - 1) & 1u);
      if (tc_elect()) {
        mbar_expect_tx(&hi_full[s], (uint32_t)stage_bytes);
        uint8_t* st = hi + (size_t)s * stage_bytes;
        const int row0 = (c0 + it) * DT_ROWS;
        for (int x = 0; x < NXC; ++x) dt_tma_load(st + x * DT_CHUNK_BYTES, &a.xmap[a.xc_map[x]], a.xc_col0[x], row0, &hi_full[s]);
        for (int z = 0; z < NZC; ++z) dt_tma_load(st + (NXC + z) * DT_CHUNK_BYTES, &a.zmap, 32 * z, row0, &hi_full[s]);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ---- MMA issuer -------------------------------------------------------------------------------------------------------
    // kind::tf32, D = f32, A and B MN-major (bits 15 / 16), N = NT, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi_addr = smem_u32(hi), lo_addr = smem_u32(lo);
    for (int it = 0; it < total; ++it) {
      const int s = it % NS, b = it % NLO;
      mbar_wait_bounded(&lo_full[b], (uint32_t)(it / NLO) & 1u);      // converters done => the stage's boxes have landed too
      tc_fence_after();
      const uint32_t hs = hi_addr + (uint32_t)s * stage_bytes, ls = lo_addr + (uint32_t)b * stage_bytes;
      if (tc_elect()) {
#pragma unroll 4
        for (int k8 = 0; k8 < DT_ROWS / 8; ++k8) {
          const uint32_t ko = (uint32_t)k8 * 1024u;
          const uint64_t dah = dt_desc_mn(hs + NXC * DT_CHUNK_BYTES + ko, DT_CHUNK_BYTES), dal = dt_desc_mn(ls + NXC * DT_CHUNK_BYTES + ko, DT_CHUNK_BYTES);
          const uint64_t dbh = dt_desc_mn(hs + ko, DT_CHUNK_BYTES), dbl = dt_desc_mn(ls + ko, DT_CHUNK_BYTES);
          tc_mma_tf32(tmem_d, dal, dbh, idesc, (it | k8) ? 1u : 0u);
          tc_mma_tf32(tmem_d, dah, dbl, idesc, 1u);
          tc_mma_tf32(tmem_d, dah, dbh, idesc, 1u);
        }
        tc_commit(&hi_empty[s]);
        tc_commit(&lo_empty[b]);
      }
      __syncwarp();
    }
    if (tc_elect()) tc_commit(&done_bar);
    __syncwarp();
  } else {
    // ---- converters: thread = (16-byte unit u of a 128-byte line, lines rr and rr + 16 of every chunk) ---------------------
    const int u = tid & 7, rr = tid >> 3;
    const int off0 = rr * 128 + ((((u >> 1) ^ (rr & 3)) << 5) | ((u & 1) << 4));   // lines rr + 16 i share rr & 3: same swizzled unit
    const int NL2 = DT_ROWS / 32;                        // pairs of lines per chunk and thread
    float zs[4][4];
#pragma unroll
    for (int z = 0; z < 4; ++z)
#pragma unroll
      for (int v = 0; v < 4; ++v) zs[z][v] = 0.f;
    for (int it = 0; it < total; ++it) {
      const int s = it % NS, b = it % NLO;
      mbar_wait_bounded(&hi_full[s], (uint32_t)(it / NS) & 1u);
      if (it >= NLO) mbar_wait_bounded(&lo_empty[b], (uint32_t)(it / NLO - 1) & 1u);
      const uint8_t* hs = hi + (size_t)s * stage_bytes;
      uint8_t* ls = lo + (size_t)b * stage_bytes;
      for (int x = 0; x < NXC; ++x) {
        for (int i = 0; i < NL2; ++i) {
          const int o = x * DT_CHUNK_BYTES + off0 + i * 4096;
          const uint4 v0 = *reinterpret_cast<const uint4*>(hs + o);
          const uint4 v1 = *reinterpret_cast<const uint4*>(hs + o + 2048);
          *reinterpret_cast<uint4*>(ls + o) = make_uint4(tc_lo(v0.x), tc_lo(v0.y), tc_lo(v0.z), tc_lo(v0.w));
          *reinterpret_cast<uint4*>(ls + o + 2048) = make_uint4(tc_lo(v1.x), tc_lo(v1.y), tc_lo(v1.z), tc_lo(v1.w));
        }
      }
#pragma unroll
      for (int z = 0; z < 4; ++z) {
        if (z < NZC) {
          for (int i = 0; i < NL2; ++i) {
            const int o = (NXC + z) * DT_CHUNK_BYTES + off0 + i * 4096;
            const uint4 v0 = *reinterpret_cast<const uint4*>(hs + o);
            const uint4 v1 = *reinterpret_cast<const uint4*>(hs + o + 2048);
            *reinterpret_cast<uint4*>(ls + o) = make_uint4(tc_lo(v0.x), tc_lo(v0.y), tc_lo(v0.z), tc_lo(v0.w));
            *reinterpret_cast<uint4*>(ls + o + 2048) = make_uint4(tc_lo(v1.x), tc_lo(v1.y), tc_lo(v1.z), tc_lo(v1.w));
            zs[z][0] += __uint_as_float(v0.x) + __uint_as_float(v1.x);
            zs[z][1] += __uint_as_float(v0.y) + __uint_as_float(v1.y);
            zs[z][2] += __uint_as_float(v0.z) + __uint_as_float(v1.z);
            zs[z][3] += __uint_as_float(v0.w) + __uint_as_float(v1.w);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&lo_full[b]);
    }
    DT_STAMP(2);
    // db_j = column sums of dz: 16 partial sums (one per line group rr) per column
#pragma unroll
    for (int z = 0; z < 4; ++z)
#pragma unroll
      for (int v = 0; v < 4; ++v) s_zred[rr][32 * z + 4 * u + v] = zs[z][v];
    named_bar_sync(1, DT_CONV_THREADS);
    {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) t += s_zred[i][tid];
      s_db[tid] = tid < H ? t : 0.f;
    }
  }
  // ---- flush: accumulator -> shared memory (sD[c][j]) -> BN algebra -> this CTA's partial slot -------------------------
  constexpr int HS = 129;
  float* sD = reinterpret_cast<float*>(base);          // [NT][HS] over the (drained) ring
  if (warp < 4) {
    if (total > 0) {
      mbar_wait_bounded(&done_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16);
      for (int cc = 0; cc < NT; cc += 8) {
        float v[8];
        tmem_ld8(taddr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) sD[(cc + i) * HS + 32 * warp + lane] = v[i];
      }
    } else {
      for (int e = tid; e < NT * 128; e += DT_CONV_THREADS) sD[(e >> 7) * HS + (e & 127)] = 0.f;
    }
    tc_fence_before();
  }
  __syncthreads();
  DT_STAMP(3);
  {
    float* part = a.partial + (size_t)blockIdx.x * a.n_params;
    const float* sdb = s_db;
    const int KH = K * H;
    // one reduction per element and launch into this CTA's slot: fire-and-forget RED.ADD (no read latency); the order of the
    // additions to one address is the launch order, so the sums stay deterministic
    for (int e = tid; e < KH; e += DT_THREADS) {
      const int c = (int)__umulhi((unsigned)e, a.h_magic), j = e - c * H;     // e / H without the conversion pipe (e < 2^16)
      const float x = sD[s_kp[c] * HS + j];
      atomicAdd(part + e, a.bnA ? s_gam[c] * fmaf(s_bA[c], x, s_bB[c] * sdb[j]) + s_bet[c] * sdb[j] : x);
    }
    for (int j = tid; j < H; j += DT_THREADS) atomicAdd(part + a.bias_off + j, sdb[j]);
    if (a.bn_partial) {
      float* bp = a.bn_partial + (size_t)blockIdx.x * 2 * K;
      for (int c = tid; c < K; c += DT_THREADS) {
        const float* wr = a.W + (size_t)c * H;
        const float* dr = sD + s_kp[c] * HS;
        float P = 0.f, Q = 0.f;
        for (int j0 = 0; j0 < H; j0 += 16) {           // 16 weight loads in flight (L2 latency), then the two dot products
          float w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = j0 + i < H ? wr[j0 + i] : 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int j = j0 + i < H ? j0 + i : 0;
            P = fmaf(w[i], sdb[j], P);
            Q = fmaf(w[i], dr[j], Q);
          }
        }
        bp[c] = P;
        bp[K + c] = fmaf(s_bA[c], Q, s_bB[c] * P);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  DT_STAMP(4);
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}

static size_t dt_smem(const DwTmaArgs& a) {
  const size_t chunk = (size_t)a.rows * 128;
  const size_t stage = (size_t)(a.n_xc + a.n_zc) * chunk;
  const size_t ring = (size_t)(a.n_stages + a.n_lo) * stage + 4 * chunk;    // + the A operand's 4 M atoms may reach past the last slot
  const size_t sd = (size_t)32 * a.n_xc * 129 * sizeof(float);
  return (ring > sd ? ring : sd) + 1024;
}
static size_t dt_smem_cap() {
  static size_t cap = 0;
  if (!cap) {
    cudaFuncAttributes fa;
    int dev = 0, optin = 0;
    if (cudaFuncGetAttributes(&fa, dw_tma_kernel) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    cap = (size_t)optin - fa.sharedSizeBytes;
  }
  return cap;
}

int dw_tma_rows(int n_xc, int n_zc) {                 // rows per stage: the largest box with a stage of <= 40 KB
  static const int forced = getenv("GNNFP_DW_ROWS") ? atoi(getenv("GNNFP_DW_ROWS")) : 0;
  if (forced == 32 || forced == 64 || forced == 128) return forced;
  for (int r = 128; r > 32; r /= 2)
    if ((n_xc + n_zc) * r * 128 <= 40 * 1024) return r;
  return 32;
}

int dw_tma_finish(DwTmaArgs& a) {
  if (a.n_xc < 1 || a.n_xc > DT_MAXXC || a.n_zc < 1 || a.n_zc > 4 || a.H > 128 || a.K > 256 || a.n_pieces < 1 || a.n_pieces > DT_MAXP)
    GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "dw_tma: shape outside the kernel's limits (%d X chunks, %d dz chunks, H %d, K %d)", a.n_xc, a.n_zc, a.H, a.K);
  const size_t cap = dt_smem_cap();
  if (!cap) GNNFP_FAIL(GNNFP_E_CUDA, "dw_tma: cannot query the shared-memory budget");
  static const int max_st = getenv("GNNFP_DW_STAGES") ? atoi(getenv("GNNFP_DW_STAGES")) : DT_MAXSTAGES;
  a.h_magic = (unsigned)((0x100000000ull + (unsigned)a.H - 1) / (unsigned)a.H);
  static const int max_lo = getenv("GNNFP_DW_LO") ? atoi(getenv("GNNFP_DW_LO")) : 3;
  a.n_stages = 2; a.n_lo = 2;
  if (a.rows != 32 && a.rows != 64 && a.rows != 128) GNNFP_FAIL(GNNFP_E_INVALID, "dw_tma: %d rows per stage", a.rows);
  if (dt_smem(a) > cap) GNNFP_FAIL(GNNFP_E_UNSUPPORTED, "dw_tma: %zu bytes of shared memory needed, %zu available", dt_smem(a), cap);
  // grow in the order: 3 stages, 3 lo slots, then stages while they fit
  auto grow = [&](int& v, int lim) { if (v >= lim) return false; ++v; if (dt_smem(a) > cap) { --v; return false; } return true; };
  grow(a.n_stages, max_st < DT_MAXSTAGES ? max_st : DT_MAXSTAGES);
  grow(a.n_lo, max_lo < DT_MAXLO ? max_lo : DT_MAXLO);
  while (grow(a.n_stages, max_st < DT_MAXSTAGES ? max_st : DT_MAXSTAGES)) {}
  while (grow(a.n_lo, max_lo < DT_MAXLO ? max_lo : DT_MAXLO)) {}
  return GNNFP_OK;
}

int launch_dw_tma(const DwTmaArgs& a, cudaStream_t s, int prof_cat, int* grid_out) {
  if (a.n_rows <= 0) { if (grid_out) *grid_out = 0; return GNNFP_OK; }
  const size_t smem = dt_smem(a);
  static size_t attr = 0;
  if (smem > attr) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(dw_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int n_chunks = (a.n_rows + a.rows - 1) / a.rows;
  const int nsm = gnnfp_num_sms();
  int grid = (a.n_rows + 255) / 256;                   // >= 256 rows per CTA
  if (grid > n_chunks) grid = n_chunks;
  if (grid > nsm) grid = nsm;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = grid;
  ProfScope ps(prof_cat, s);
  dw_tma_kernel<<<grid, DT_THREADS, smem, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
