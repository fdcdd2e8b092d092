// agg_tile.cu - Adj^T . S as a TMA-pipelined tile kernel (reference GNN.py:228 `tf.sparse.sparse_dense_matmul(adjacency,
// state, adjoint_a=True)`; replaces the streaming agg_stats_kernel for the interleaved X-slot layout).
//
// Batches of small graphs are block diagonal: almost every in-neighbour of a row lies in the row's own 128-row tile.
// Per persistent CTA (two per SM):
//   warp 8 (one lane)  producer: TMA box loads of S chunk tiles [128 rows x 32 columns] (SWIZZLE_128B) into an input ring,
//                      bulk copies of the tile's slice of the dst-CSR (row pointers, tile-local sources, weights);
//   warps 0-3          gather, thread = row: in-arcs in dst-CSR order (ascending arc id, sequential fmaf - the summation
//                      order of TF's SparseTensorDenseMatMul and of agg_stats_kernel), neighbours' rows read from the
//                      staged tile with swizzled 128-bit loads; the rare out-of-tile neighbour (graphs straddling a tile
//                      boundary, graphs larger than a tile) is read from global memory; result row -> output stage;
//   warps 4-7          store, lane = column: coalesced 128-byte row stores of the output stage (the Adj^T S block of an X
//                      slot starts at column D, which a tensor-map store cannot address when D % 4 != 0) + fp64 column
//                      statistics of the result (the next BatchNormalization's batch sums).
#include <cudaTypedefs.h>

#include "rows_tma.h"
#include "tc.cuh"

#define AT_THREADS 288
#define AT_W_GATHER 0
#define AT_W_STORE 4
#define AT_W_PROD 8
#define AT_NS 3                                               // input stages
#define AT_NO 2                                               // output stages
#define AT_RP_BYTES 528
#define AT_LI_BYTES ((GNNFP_TILE_ARCS + 8) * 2)
#define AT_W_BYTES ((GNNFP_TILE_ARCS + 8) * 4)
#define AT_CSR_BUF (544 + 2080 + AT_W_BYTES)
#define AT_SLOTS 4                                            // statistics slots (column chunks): D <= 128

struct AggTileArgs {
  CUtensorMap map_s;         // S [n_rows x D], row pitch ld_s
  int n_rows, D;
  const float* S; int ld_s;  // the same matrix for the out-of-tile reads
  const int* rowptr; const int* src; const float* wgt;          // dst-CSR (wgt NULL = 1)
  const short* lidx; const int* arc0;                           // tile-local view (graph.cu)
  float* out; int ld_out;
  double* st_sum; double* st_sq;                                // [D] or NULL
  const int* gate;
};

__device__ __forceinline__ void at_tma_load(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(AT_THREADS, 2) agg_tile_kernel(const __grid_constant__ AggTileArgs a) {
  if (a.gate && *a.gate == 0) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* inst = base;                                        // [AT_NS] S chunk tiles
  uint8_t* outst = inst + AT_NS * RT_STAGE_BYTES;              // [AT_NO] result tiles
  uint8_t* csrb = outst + AT_NO * RT_STAGE_BYTES;              // [2] CSR slices
  __shared__ __align__(8) uint64_t in_full[AT_NS], in_empty[AT_NS], out_full[AT_NO], out_empty[AT_NO], csr_full[2], csr_empty[2];
  __shared__ double s_stat[4][AT_SLOTS][2][32];
  __shared__ uint8_t s_bnd[2][RT_ROWS];                       // row has an in-neighbour outside the tile: handled by the store warps
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = a.n_rows, D = a.D;
  const int n_tiles = (n + RT_ROWS - 1) / RT_ROWS;
  const int NCH = (D + 31) / 32;
  if (tid == 0) {
    for (int i = 0; i < AT_NS; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 128); }
    for (int i = 0; i < AT_NO; ++i) { mbar_init(&out_full[i], 128); mbar_init(&out_empty[i], 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(&csr_full[i], 1); mbar_init(&csr_empty[i], 256); }
  }
  fence_proxy_async();
  __syncthreads();
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == AT_W_PROD) {
    int slot = 0;
    uint32_t use = 0;
    for (int tq = 0; tq < my_tiles; ++tq) {
      const int tile = blockIdx.x + tq * gridDim.x, b = tq & 1;
      if (tq >= 2) mbar_wait_bounded(&csr_empty[b], (uint32_t)((tq >> 1) - 1) & 1u);
      const int start = a.arc0[tile] & ~7;
      if (tc_elect()) {
        uint8_t* buf = csrb + (size_t)b * AT_CSR_BUF;
        mbar_expect_tx(&csr_full[b], AT_RP_BYTES + AT_LI_BYTES + (a.wgt ? AT_W_BYTES : 0));
        bulk_g2s(buf, a.rowptr + (size_t)tile * RT_ROWS, AT_RP_BYTES, &csr_full[b]);
        bulk_g2s(buf + 544, a.lidx + start, AT_LI_BYTES, &csr_full[b]);
        if (a.wgt) bulk_g2s(buf + 544 + 2080, a.wgt + start, AT_W_BYTES, &csr_full[b]);
      }
      __syncwarp();
      for (int c = 0; c < NCH; ++c) {
        if (use > 0) mbar_wait_bounded(&in_empty[slot], (use - 1) & 1);
        if (tc_elect()) {
          mbar_expect_tx(&in_full[slot], RT_STAGE_BYTES);
          at_tma_load(inst + (size_t)slot * RT_STAGE_BYTES, &a.map_s, 32 * c, tile * RT_ROWS, &in_full[slot]);
        }
        __syncwarp();
        if (++slot == AT_NS) { slot = 0; ++use; }
      }
    }
  } else if (warp < AT_W_STORE) {
    // =================== gather: thread = row ============================================================================
    const int r = tid;
    const int rbase = (r >> 3) * 1024 + (r & 7) * 128, rx = r & 7;
    int islot = 0, oslot = 0;
    uint32_t iph = 0, ouse = 0;
    for (int tq = 0; tq < my_tiles; ++tq) {
      const int tile = blockIdx.x + tq * gridDim.x, b = tq & 1;
      const int row0 = tile * RT_ROWS;
      mbar_wait_bounded(&csr_full[b], (uint32_t)(tq >> 1) & 1u);
      const uint8_t* buf = csrb + (size_t)b * AT_CSR_BUF;
      const int* rp = reinterpret_cast<const int*>(buf);
      const int arcs0 = rp[0];
      const short* li = reinterpret_cast<const short*>(buf + 544) + (arcs0 & 7);
      const float* wv = a.wgt ? reinterpret_cast<const float*>(buf + 544 + 2080) + (arcs0 & 7) : nullptr;
      const bool over = rp[RT_ROWS] - arcs0 > GNNFP_TILE_ARCS;  // the staged slice does not cover the tile: read the CSR from global
      const bool rowok = row0 + r < n;
      const int p0 = rp[r];                                    // global arc positions p0 .. p0 + na
      int na = rowok ? rp[r + 1] - p0 : 0;
      // a row with a neighbour outside the tile (graphs straddling a tile boundary, oversized graphs / tiles) is left to the
      // store warps, which read global memory with coalesced lane = column accesses: inside this thread = row loop ONE such
      // lane would drag its whole warp through 32 scalar loads per arc (12 % of the rows, i.e. nearly every warp)
      bool good = !over;
      for (int q = 0; q < na && good; ++q) good = li[p0 + q - arcs0] >= 0;
      s_bnd[b][r] = (rowok && !good) ? 1 : 0;
      if (!good) na = 0;
      for (int c = 0; c < NCH; ++c) {
        mbar_wait_bounded(&in_full[islot], iph);
        const uint8_t* stg = inst + (size_t)islot * RT_STAGE_BYTES;
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        for (int q = 0; q < na; ++q) {
          const int p = p0 + q;
          const int l = li[p - arcs0];
          const float w = wv ? wv[p - arcs0] : 1.0f;
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
        mbar_arrive(&in_empty[islot]);
        if (++islot == AT_NS) { islot = 0; iph ^= 1u; }
        if (ouse > 0) mbar_wait_bounded(&out_empty[oslot], (ouse - 1) & 1);
        uint8_t* od = outst + (size_t)oslot * RT_STAGE_BYTES + rbase;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<float4*>(od + ((ch ^ rx) << 4)) = make_float4(acc[4 * ch], acc[4 * ch + 1], acc[4 * ch + 2], acc[4 * ch + 3]);
        mbar_arrive(&out_full[oslot]);
        if (++oslot == AT_NO) { oslot = 0; ++ouse; }
      }
      mbar_arrive(&csr_empty[b]);
    }
  } else {
    // =================== store + statistics: lane = column ============================================================
    const int cw = warp - AT_W_STORE;
    const int coff = ((lane >> 2) << 4), cin = (lane & 3) << 2;
    double s1[AT_SLOTS], s2[AT_SLOTS];
#pragma unroll
    for (int q = 0; q < AT_SLOTS; ++q) { s1[q] = 0.0; s2[q] = 0.0; }
    int oslot = 0;
    uint32_t oph = 0;
    for (int tq = 0; tq < my_tiles; ++tq) {
      const int tile = blockIdx.x + tq * gridDim.x, b = tq & 1;
      const int rfirst = tile * RT_ROWS + 32 * cw;
      const int nv = min(32, max(0, n - rfirst));
      mbar_wait_bounded(&csr_full[b], (uint32_t)(tq >> 1) & 1u);
      const int* rp = reinterpret_cast<const int*>(csrb + (size_t)b * AT_CSR_BUF) + 32 * cw;
      for (int c = 0; c < NCH; ++c) {
        mbar_wait_bounded(&out_full[oslot], oph);             // (also publishes s_bnd of this tile)
        const uint8_t* st = outst + (size_t)oslot * RT_STAGE_BYTES + (4 * cw) * 1024;
        const bool colok = 32 * c + lane < D;
        float* op = a.out + (size_t)rfirst * a.ld_out + 32 * c + lane;
        float p1 = 0.f, p2 = 0.f;
        unsigned bm = 0;                                       // rows of this warp left to it by the gather threads
        if (lane < nv && s_bnd[b][32 * cw + lane]) bm = 1u;
        bm = __ballot_sync(0xffffffffu, bm != 0);
        if (colok) {
#pragma unroll 8
          for (int rr = 0; rr < nv; ++rr) {
            const float x = *reinterpret_cast<const float*>(st + (rr >> 3) * 1024 + (rr & 7) * 128 + (coff ^ ((rr & 7) << 4)) + cin);
            if (!((bm >> rr) & 1u)) {
              op[(size_t)rr * a.ld_out] = x;
              p1 += x;
              p2 = fmaf(x, x, p2);
            }
          }
        }
        while (bm) {                                           // boundary rows: the whole CSR row from global memory, coalesced
          const int rr = __ffs(bm) - 1;
          bm &= bm - 1;
          float x = 0.f;
          for (int p = rp[rr]; p < rp[rr + 1]; ++p) {
            const float w = a.wgt ? a.wgt[p] : 1.0f;
            const float v = colok ? __ldg(a.S + (size_t)a.src[p] * a.ld_s + 32 * c + lane) : 0.f;
            x = fmaf(w, v, x);
          }
          if (colok) {
            op[(size_t)rr * a.ld_out] = x;
            p1 += x;
            p2 = fmaf(x, x, p2);
          }
        }
        mbar_arrive(&out_empty[oslot]);
#pragma unroll
        for (int q = 0; q < AT_SLOTS; ++q)
          if (q == c) { s1[q] += (double)p1; s2[q] += (double)p2; }
        if (++oslot == AT_NO) { oslot = 0; oph ^= 1u; }
      }
      mbar_arrive(&csr_empty[b]);
    }
#pragma unroll
    for (int q = 0; q < AT_SLOTS; ++q) { s_stat[cw][q][0][lane] = s1[q]; s_stat[cw][q][1][lane] = s2[q]; }
  }
  __syncthreads();
  if (a.st_sum) {
    for (int e = tid; e < NCH * 32; e += AT_THREADS) {
      const int q = e >> 5, j = e & 31, col = 32 * q + j;
      if (col < D) {
        atomicAdd(a.st_sum + col, (s_stat[0][q][0][j] + s_stat[1][q][0][j]) + (s_stat[2][q][0][j] + s_stat[3][q][0][j]));
        atomicAdd(a.st_sq + col, (s_stat[0][q][1][j] + s_stat[1][q][1][j]) + (s_stat[2][q][1][j] + s_stat[3][q][1][j]));
      }
    }
  }
}

// eligible: the interleaved layout's state block (16-byte aligned rows), D <= 128, tile view present
int agg_tile_supported(const float* S, int ld_s, int D) {
  // Measured on B200 (C2 widths, 248 k rows): 38 .. 118 us per launch with the out-of-tile reads inside the thread = row
  // loop, 111 .. 277 us with them moved to the store warps, against 28 .. 92 us of the streaming agg_stats_kernel, whose
  // thousands of independent threads hide the dependent rowptr -> idx -> row chain better than 8 gather warps per SM do.
  // The tile kernel is therefore only taken on request (GNNFP_AGG_TILE=1) until boundary rows get their own wide pass.
  static const int on = getenv("GNNFP_AGG_TILE") ? 1 : 0;
  return on && rows_tma_available() && rows_tma_ok(S, ld_s) && D >= 1 && D <= 32 * AT_SLOTS;
}

int launch_agg_tile(const float* S, int ld_s, int n_rows, int D, const int* rowptr, const int* src, const float* wgt,
                    const short* lidx, const int* arc0, float* out, int ld_out, double* st_sum, double* st_sq,
                    const int* gate, cudaStream_t s, int prof_cat) {
  if (n_rows <= 0) return GNNFP_OK;
  AggTileArgs a;
  memset(&a, 0, sizeof(a));
  int rc;
  if ((rc = rows_tma_map(&a.map_s, S, n_rows, D, ld_s))) return rc;
  a.n_rows = n_rows; a.D = D; a.S = S; a.ld_s = ld_s;
  a.rowptr = rowptr; a.src = src; a.wgt = wgt; a.lidx = lidx; a.arc0 = arc0;
  a.out = out; a.ld_out = ld_out; a.st_sum = st_sum; a.st_sq = st_sq; a.gate = gate;
  const size_t smem = (size_t)(AT_NS + AT_NO) * RT_STAGE_BYTES + 2 * AT_CSR_BUF + 1024;
  static bool attr = false;
  if (!attr) {
    GNNFP_CHECK_CUDA(cudaFuncSetAttribute(agg_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int n_tiles = (n_rows + RT_ROWS - 1) / RT_ROWS;
  int grid = 2 * gnnfp_num_sms();
  if (grid > n_tiles) grid = n_tiles;
  ProfScope ps(prof_cat ? prof_cat : PC_AGG, s);
  agg_tile_kernel<<<grid, AT_THREADS, smem, s>>>(a);
  GNNFP_COUNT_LAUNCH();
  GNNFP_CHECK_CUDA(cudaGetLastError());
  return GNNFP_OK;
}
