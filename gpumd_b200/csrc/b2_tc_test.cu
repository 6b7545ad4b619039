// b2_tc_test.cu -- self-test of the tcgen05 layer (b2_tc.cuh): one CTA computes
// D[128 x N] = A[128 x K] . B[N x K]^T with the 3xTF32 split and returns it, so that the
// descriptor / TMEM plumbing can be validated against a host GEMM before the MLP depends on it.
#include "../../include/b200md.h"
#include "b2_host.h"
#include "b2_tc.cuh"

namespace b2 {
namespace {

// smem: A_hi, A_lo [128 x K], B_hi, B_lo [N x K] in the interleaved K-major layout
// layout 0: 8-row groups strided by a whole K row of core matrices (LBO 128 B, SBO K/4 * 128 B)
// layout 1: K chunks strided by a whole column of core matrices (SBO 128 B, LBO rows * 16 B) --
//           the one k_mlp_tc uses, because a thread that owns a row then writes conflict-free float4s
__global__ void __launch_bounds__(128) k_tc_gemm_test(
  int layout, int N, int K, const float* __restrict__ A, const float* __restrict__ B,
  float* __restrict__ D)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lbo_a = layout ? 128u * 16u : 128u;
  const uint32_t sbo_a = layout ? 128u : (uint32_t)(K / 4) * 128u;
  const uint32_t lbo_b = layout ? (uint32_t)N * 16u : 128u;
  const uint32_t sbo_b = sbo_a;
  const uint32_t a_bytes = 128u * K * 4u, b_bytes = (uint32_t)N * K * 4u;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;
  // stage operands (generic proxy stores)
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, k = e % K;
    float hi, lo;
    b2tc::split_tf32(A[e], hi, lo);
    const uint32_t off = b2tc::kmajor_offset(r, k, lbo_a, sbo_a);
    *reinterpret_cast<float*>(a_hi + off) = hi;
    *reinterpret_cast<float*>(a_lo + off) = lo;
  }
  for (int e = tid; e < N * K; e += 128) {
    const int r = e / K, k = e % K;
    float hi, lo;
    b2tc::split_tf32(B[e], hi, lo);
    const uint32_t off = b2tc::kmajor_offset(r, k, lbo_b, sbo_b);
    *reinterpret_cast<float*>(b_hi + off) = hi;
    *reinterpret_cast<float*>(b_lo + off) = lo;
  }
  uint32_t ncols = 32;
  while ((int)ncols < N)
    ncols <<= 1;
  if (warp == 0)
    b2tc::tmem_alloc(&tmem_slot, ncols);
  if (tid == 0)
    b2tc::mbar_init(&bar, 1);
  b2tc::fence_async_smem();
  b2tc::fence_before_sync();
  __syncthreads();
  b2tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = b2tc::make_idesc_tf32(128, N);
    const uint32_t ah = b2tc::smem_u32(a_hi), al = b2tc::smem_u32(a_lo);
    const uint32_t bh = b2tc::smem_u32(b_hi), bl = b2tc::smem_u32(b_lo);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t adv_a = (uint32_t)ks * 2u * lbo_a; // two 16-byte chunks per K = 8
      const uint32_t adv_b = (uint32_t)ks * 2u * lbo_b;
      const uint64_t dah = b2tc::make_desc(ah + adv_a, lbo_a, sbo_a);
      const uint64_t dal = b2tc::make_desc(al + adv_a, lbo_a, sbo_a);
      const uint64_t dbh = b2tc::make_desc(bh + adv_b, lbo_b, sbo_b);
      const uint64_t dbl = b2tc::make_desc(bl + adv_b, lbo_b, sbo_b);
      b2tc::mma_tf32(tmem, dah, dbh, idesc, ks > 0 ? 1u : 0u);
      b2tc::mma_tf32(tmem, dah, dbl, idesc, 1u);
      b2tc::mma_tf32(tmem, dal, dbh, idesc, 1u);
    }
    b2tc::mma_commit(&bar);
  }
  b2tc::mbar_wait(&bar, 0);
  b2tc::fence_after_sync();
  // epilogue: thread t owns row t
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    b2tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    for (int c = 0; c < 32 && c0 + c < N; ++c)
      D[(size_t)tid * N + c0 + c] = __uint_as_float(v[c]);
  }
  b2tc::fence_before_sync();
  __syncthreads();
  if (warp == 0)
    b2tc::tmem_dealloc(tmem, ncols);
}

} // namespace
} // namespace b2

using namespace b2;

extern "C" int b200md_tc_selftest(
  int layout, int N, int K, const float* d_A, const float* d_B, float* d_D, void* stream)
{
  if (layout < 0 || layout > 1) {
    set_error("b200md_tc_selftest: layout must be 0 or 1");
    return B200MD_ERR_ARG;
  }
  if (N < 16 || N > 256 || N % 16 != 0 || K < 8 || K % 8 != 0) {
    set_error("b200md_tc_selftest: need 16 <= N <= 256 (multiple of 16) and K a multiple of 8");
    return B200MD_ERR_ARG;
  }
  const size_t bytes = 2 * (size_t)(128 + N) * K * sizeof(float);
  if (bytes > 200 * 1024) {
    set_error("b200md_tc_selftest: operands do not fit in shared memory");
    return B200MD_ERR_ARG;
  }
  if (bytes > 48 * 1024)
    B2_CUDA(cudaFuncSetAttribute(k_tc_gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  k_tc_gemm_test<<<1, 128, bytes, (cudaStream_t)stream>>>(layout, N, K, d_A, d_B, d_D);
  B2_LAUNCHED();
  return B200MD_OK;
}
