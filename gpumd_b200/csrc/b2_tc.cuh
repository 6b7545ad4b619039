// b2_tc.cuh -- minimal hand-written tcgen05 / TMEM layer for sm_100a (raw PTX, no CUTLASS types).
//
// Used by the tensor-core NEP MLP: D[128 x N] (FP32 in TMEM) += A[128 x K] . B[N x K]^T with
// A, B in shared memory as K-major, non-swizzled ("interleaved") core matrices:
//   byte offset of element (row r, column k), 4-byte elements:
//     (r % 8) * 16  +  (r / 8) * SBO  +  (k / 4) * LBO  +  (k % 4) * 4
// i.e. a core matrix is 8 rows x 16 bytes stored contiguously (128 B); SBO is the stride between
// 8-row groups, LBO the stride between 16-byte K chunks.  One tcgen05.mma.kind::tf32 consumes
// K = 8 (two chunks).  FP32 accuracy comes from the 3xTF32 split (hi*hi + hi*lo + lo*hi).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, descriptor version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // version = 1
  return d;                                             // base_offset 0, lbo_mode 0, layout NONE
}

// 32-bit instruction descriptor: TF32 x TF32 -> F32, both operands K-major, dense
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N)
{
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 2u << 7;                    // a_format = TF32
  d |= 2u << 10;                   // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17;   // n_dim
  d |= (uint32_t)(M >> 4) << 24;   // m_dim
  return d;                        // a_major = b_major = K (0), no negate, no sparsity
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                 smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void fence_before_sync()
{
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync()
{
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core reads smem)
__device__ __forceinline__ void fence_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (!done) {
    if (++spins > (1u << 22))
      __trap(); // a lost arrival would otherwise hang the GPU; fail loudly instead
    asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(addr), "r"(parity)
      : "memory");
  }
}

// D[tmem] (+)= A[smem desc] . B[smem desc]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(
  uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
  const uint32_t zero = 0;
  asm volatile(
    "{\n\t"
    ".reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
    "}"
    :
    : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(zero), "r"(zero),
      "r"(zero), "r"(zero)
    : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                 smem_u32(bar))
               : "memory");
}

// this warp's 32 TMEM lanes x 32 consecutive columns -> 32 registers per thread (thread = lane/row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v)
{
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
      "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
      "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
    : "r"(taddr)
    : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 3xTF32 split: x = hi + lo with hi = tf32(x) (round to nearest), lo = tf32(x - hi)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo)
{
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float r = x - hi;
  uint32_t l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}

// byte offset of element (r, k) in the K-major interleaved layout described above
__device__ __forceinline__ uint32_t kmajor_offset(int r, int k, uint32_t lbo, uint32_t sbo)
{
  return (uint32_t)(r & 7) * 16u + (uint32_t)(r >> 3) * sbo + (uint32_t)(k >> 2) * lbo +
         (uint32_t)(k & 3) * 4u;
}

// 16-column variant of tmem_ld32
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v)
{
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
      "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
      "=r"(v[14]), "=r"(v[15])
    : "r"(taddr)
    : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// TMA bulk copy global -> shared (bytes a multiple of 16, both addresses 16-byte aligned); the
// mbarrier receives complete_tx(bytes).  Issued by one thread after mbar_expect_tx.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
      smem_u32(smem_dst)),
    "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
    : "memory");
}

} // namespace b2tc
