// tc.cuh - tcgen05 / TMEM / mbarrier / TF32-split helpers shared by the tensor-core kernels (gemm_tc.cu, rows_tma.cu)
#pragma once
#include "tile.cuh"

__device__ __forceinline__ void tc_cp_async8(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tc_cp_async4(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void tc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tc_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bounded mbarrier wait: try_wait suspends the thread in hardware for a bounded time slice (no busy polling that would
// steal issue slots from the working warps); a protocol error traps (the launch fails) instead of hanging the device
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  for (int spin = 0;; ++spin) {
    __nanosleep(40);                                   // a waiting role must not eat the issue slots of the working warps
    if (mbar_try(bar, parity)) return;
    if ((spin & 1023) == 1023 && clock64() - t0 > 120000000000ll) __trap();   // ~60 s at 2 GHz: a protocol error must not hang
                                                       // the device for ever, yet time-slicing with other contexts must not trip it
  }
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// one lane of a converged warp (elect.sync): the issue roles keep their loops warp-uniform and predicate only the instruction
__device__ __forceinline__ bool tc_elect() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
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

// TF32 split without a conversion of the high part: the tensor core reads the top 19 bits of an fp32 operand (the low 13
// mantissa bits are ignored), so a itself serves as a_hi = trunc(a); a_lo = a - trunc(a) is exact in fp32 and is rounded to
// 11 significant bits by adding half a TF32 ulp before the hardware truncation
__device__ __forceinline__ uint32_t tc_lo(uint32_t abits) {
  const float lo = __uint_as_float(abits) - __uint_as_float(abits & 0xFFFFE000u);
  return __float_as_uint(lo) + 0x1000u;
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
