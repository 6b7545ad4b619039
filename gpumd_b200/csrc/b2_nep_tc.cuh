// b2_nep_tc.cuh -- the NEP hidden layer on the 5th-generation tensor cores (device only).
//
// apply_ann_one_layer (src/utilities/nep_utilities.cuh:166-193 of the reference) for 128 atoms of
// ONE type at a time, as two GEMMs with a pointwise stage between them:
//     Z  [128 x HN] = Q [128 x DK] . W0^T              (tcgen05.mma kind::tf32, 3xTF32 split)
//     x  = tanh(Z - b0),  E = sum_j w1_j x_j - bias,   C_j = w1_j (1 - x_j^2)
//     Fp [128 x DN] = C [128 x HN] . W0
// Rows come from the type tiles the neighbour rebuild maintains (B2NeighborView::tile_*), so
// every row of a tile shares W0.  The weights arrive as a ready-made shared-memory image
// (NepModel::tc_img) by one TMA bulk copy; accumulators live in TMEM (HN + DN columns); thread r
// of the 128-thread CTA owns row r = TMEM lane r in both epilogues.
#pragma once
#include "b2_nep.cuh"
#include "b2_tc.cuh"

namespace b2 {

// N3 = 0 when the U-table GEMM is not part of the kernel
__host__ __device__ inline int b2_tc_tmem_cols(int HN, int DN, int N3)
{
  int c = 32;
  while (c < HN + DN || c < N3)
    c <<= 1;
  return c;
}
__host__ __device__ inline size_t b2_tc_img_bytes(int img_floats)
{
  return ((size_t)img_floats * 4 + 127) / 128 * 128;
}
__host__ __device__ inline int b2_tc_ka(int HN, int DK, int K3)
{
  int ka = DK > HN ? DK : HN;
  return K3 > ka ? K3 : ka;
}
// weight image + the two A tiles (hi / lo) + the landing buffer of the TMA copy of one q tile
__host__ __device__ inline size_t b2_tc_smem_bytes(int img_floats, int HN, int DK, int K3)
{
  return b2_tc_img_bytes(img_floats) + 2 * 128 * (size_t)b2_tc_ka(HN, DK, K3) * 4 + 128 * (size_t)DK * 4;
}

// Input: the tile-major copy of q (B2NepView::qt) -- ONE cp.async.bulk per 128-atom tile into a landing
// buffer that already has the K-major order of the A operand; the copy of tile k+1 is issued as soon as
// tile k has been split into its TF32 hi / lo operands, so it flies under the three GEMMs and epilogues
// of tile k.  Output: dU/dq goes back tile-major (P.fpt) with coalesced float4 stores.
// Persistent: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... so that the TMEM
// allocation, barrier set-up and (tiles being ordered by type) almost every weight fetch are paid
// once per CTA instead of once per tile.
__global__ void __launch_bounds__(128) k_mlp_tc(B2NepView P)
{
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t bar_w, bar_mma, bar_q;
  __shared__ uint32_t tmem_slot;
  const int ntile = P.tile_meta[0];
  if ((int)blockIdx.x >= ntile)
    return;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t N = (size_t)P.n;
  const int HN = P.HN, DK = P.DK, DN = P.DN;
  const int K3 = P.K3, N3 = P.N3; // N3 = 0: the U table is left to k_utable
  const int KA = b2_tc_ka(HN, DK, N3 ? K3 : 0);
  float* img = reinterpret_cast<float*>(tc_smem);
  const float* sb0 = img + 2 * HN * DK + 2 * DN * HN;
  const float* sw1 = sb0 + HN;
  unsigned char* a_hi = tc_smem + b2_tc_img_bytes(P.tc_img_floats);
  unsigned char* a_lo = a_hi + (size_t)128 * KA * 4;
  float4* stage = reinterpret_cast<float4*>(a_lo + (size_t)128 * KA * 4); // [DK/4][128] float4
  const uint32_t q_bytes = (uint32_t)DK * 128u * 4u;
  const uint32_t ncols = (uint32_t)b2_tc_tmem_cols(HN, DN, N3);

  if (warp == 0)
    b2tc::tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) {
    b2tc::mbar_init(&bar_w, 1);
    b2tc::mbar_init(&bar_mma, 1);
    b2tc::mbar_init(&bar_q, 1);
  }
  b2tc::fence_before_sync();
  __syncthreads();
  b2tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t ah = b2tc::smem_u32(a_hi), al = b2tc::smem_u32(a_lo);
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t w_phase = 0, mma_phase = 0, q_phase = 0;
  int cur_type = -1;
  if (tid == 0) { // first q tile of this CTA
    b2tc::mbar_expect_tx(&bar_q, q_bytes);
    b2tc::bulk_g2s(stage, P.qt + (size_t)blockIdx.x * DK * 128, q_bytes, &bar_q);
  }

  int i_next = P.tile_atom[blockIdx.x * 128 + tid];
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int t = P.tile_type[tile];
    const int i = i_next; // -1 = padding row
    if (tile + (int)gridDim.x < ntile)
      i_next = P.tile_atom[(tile + gridDim.x) * 128 + tid];
    const bool new_weights = t != cur_type;      // uniform over the CTA
    if (new_weights && tid == 0) {
      // every thread is past the previous tile's last read of the image (the barrier after
      // epilogue 1) and its MMAs have completed (bar_mma), so the image may be replaced
      const uint32_t bytes = (uint32_t)P.tc_img_floats * 4u;
      b2tc::mbar_expect_tx(&bar_w, bytes);
      b2tc::bulk_g2s(img, P.tc_img + (size_t)t * P.tc_img_floats, bytes, &bar_w);
    }
    cur_type = t;
    // ---- Q tile (landed by TMA, K-major like the A operand): scale, split into TF32 hi / lo ----
    b2tc::mbar_wait(&bar_q, q_phase);
    q_phase ^= 1u;
    for (int kc = 0; kc < DK / 4; ++kc) {
      float4 v = stage[kc * 128 + tid];
      const float4 sc = __ldg(reinterpret_cast<const float4*>(P.q_scaler) + kc); // zero beyond dim
      if (i < 0)
        v = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // padding row
      const float x[4] = {v.x * sc.x, v.y * sc.y, v.z * sc.z, v.w * sc.w};
      float hi[4], lo[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        b2tc::split_tf32(x[c], hi[c], lo[c]);
      const uint32_t off = (uint32_t)tid * 16u + (uint32_t)kc * 2048u;
      *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    b2tc::fence_async_smem();
    b2tc::fence_before_sync();
    __syncthreads();
    b2tc::fence_after_sync();
    if (tid == 0 && tile + (int)gridDim.x < ntile) {
      // every thread has read the landing buffer (barrier above): fetch the next tile under this one
      b2tc::mbar_expect_tx(&bar_q, q_bytes);
      b2tc::bulk_g2s(stage, P.qt + (size_t)(tile + gridDim.x) * DK * 128, q_bytes, &bar_q);
    }
    if (new_weights) {
      b2tc::mbar_wait(&bar_w, w_phase); // weights, b0, w1 have landed
      w_phase ^= 1u;
    }
    if (tid == 0) {
      // GEMM 1: Z = Q . W0^T   (B1: [HN rows x DK], K chunks HN*16 bytes apart)
      const uint32_t idesc = b2tc::make_idesc_tf32(128, HN);
      const uint32_t bh = b2tc::smem_u32(img), bl = bh + (uint32_t)HN * DK * 4u;
      const uint32_t lbo_b = (uint32_t)HN * 16u;
      for (int ks = 0; ks < DK / 8; ++ks) {
        const uint64_t dah = b2tc::make_desc(ah + ks * 4096u, 2048u, 128u);
        const uint64_t dal = b2tc::make_desc(al + ks * 4096u, 2048u, 128u);
        const uint64_t dbh = b2tc::make_desc(bh + ks * 2u * lbo_b, lbo_b, 128u);
        const uint64_t dbl = b2tc::make_desc(bl + ks * 2u * lbo_b, lbo_b, 128u);
        b2tc::mma_tf32(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
        b2tc::mma_tf32(tmem, dah, dbl, idesc, 1u);
        b2tc::mma_tf32(tmem, dah, dbh, idesc, 1u);
      }
      b2tc::mma_commit(&bar_mma);
    }
    b2tc::mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    b2tc::fence_after_sync();
    // ---- epilogue 1: activation, site energy, and C as the A operand of GEMM 2 ----
    float F = 0.0f;
    for (int c0 = 0; c0 < HN; c0 += 16) {
      uint32_t v[16];
      b2tc::tmem_ld16(tmem + lane_base + (uint32_t)c0, v);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float hi[4], lo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = c0 + 4 * g + c;
          const float x1 = tanhf(__uint_as_float(v[4 * g + c]) - sb0[j]);
          const float w1j = sw1[j];
          F = fmaf(w1j, x1, F);
          b2tc::split_tf32(w1j * (1.0f - x1 * x1), hi[c], lo[c]);
        }
        const uint32_t off = (uint32_t)tid * 16u + (uint32_t)(c0 / 4 + g) * 2048u;
        *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    b2tc::fence_async_smem();
    b2tc::fence_before_sync();
    __syncthreads();
    b2tc::fence_after_sync();
    if (tid == 0) {
      // GEMM 2: Fp = C . W0   (B2: [DN rows x HN], K chunks DN*16 bytes apart), D at column HN
      const uint32_t idesc = b2tc::make_idesc_tf32(128, DN);
      const uint32_t bh = b2tc::smem_u32(img) + 2u * HN * DK * 4u, bl = bh + (uint32_t)DN * HN * 4u;
      const uint32_t lbo_b = (uint32_t)DN * 16u;
      for (int ks = 0; ks < HN / 8; ++ks) {
        const uint64_t dah = b2tc::make_desc(ah + ks * 4096u, 2048u, 128u);
        const uint64_t dal = b2tc::make_desc(al + ks * 4096u, 2048u, 128u);
        const uint64_t dbh = b2tc::make_desc(bh + ks * 2u * lbo_b, lbo_b, 128u);
        const uint64_t dbl = b2tc::make_desc(bl + ks * 2u * lbo_b, lbo_b, 128u);
        b2tc::mma_tf32(tmem + (uint32_t)HN, dal, dbh, idesc, ks > 0 ? 1u : 0u);
        b2tc::mma_tf32(tmem + (uint32_t)HN, dah, dbl, idesc, 1u);
        b2tc::mma_tf32(tmem + (uint32_t)HN, dah, dbh, idesc, 1u);
      }
      b2tc::mma_commit(&bar_mma);
    }
    b2tc::mbar_wait(&bar_mma, mma_phase);
    mma_phase ^= 1u;
    b2tc::fence_after_sync();
    // ---- epilogue 2: dU/dq (times q_scaler).  Angular part -> FpA for k_force_angular; radial part
    //      -> the A operand of GEMM 3 (or FpR for k_utable when the table GEMM is not fused) ----
    for (int c0 = 0; c0 < DN; c0 += 16) {
      uint32_t v[16];
      b2tc::tmem_ld16(tmem + lane_base + (uint32_t)(HN + c0), v);
      float f[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int d = c0 + c;
        f[c] = (i >= 0 && d < P.dim) ? __uint_as_float(v[c]) * __ldg(&P.q_scaler[d]) : 0.0f;
      }
      // tile-major dU/dq: row tid, chunk (c0 + 4g)/4 -- consecutive threads, consecutive 16 bytes
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (c0 + 4 * g < DK)
          reinterpret_cast<float4*>(P.fpt)[((size_t)tile * (DK / 4) + (c0 / 4 + g)) * 128 + tid] =
            make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
      if (i >= 0 && N3 == 0) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c0 + c < P.nr1)
            P.FpR[(size_t)(c0 + c) * N + i] = f[c];
      }
      if (N3 && c0 < K3) { // nr1 <= K3 <= 16: the radial part sits in the first chunk
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (4 * g < K3) {
            float hi[4], lo[4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
              b2tc::split_tf32((4 * g + c < P.nr1) ? f[4 * g + c] : 0.0f, hi[c], lo[c]);
            const uint32_t off = (uint32_t)tid * 16u + (uint32_t)g * 2048u;
            *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
    }
    if (i >= 0)
      P.acc[i] = (double)(F - __ldg(&P.bias[t]));
    if (N3) {
      // ---- GEMM 3: U[128 x N3] = FpR[128 x K3] . B3^T, accumulator at column 0 (Z is dead; every
      //      thread has finished its epilogue-2 loads at the barrier below) ----
      b2tc::fence_async_smem();
      b2tc::fence_before_sync();
      __syncthreads();
      b2tc::fence_after_sync();
      if (tid == 0) {
        const uint32_t idesc = b2tc::make_idesc_tf32(128, N3);
        const uint32_t bh = b2tc::smem_u32(img) + (2u * HN * DK + 2u * DN * HN + 2u * HN) * 4u;
        const uint32_t bl = bh + (uint32_t)N3 * K3 * 4u;
        const uint32_t lbo_b = (uint32_t)N3 * 16u;
        for (int ks = 0; ks < K3 / 8; ++ks) {
          const uint64_t dah = b2tc::make_desc(ah + ks * 4096u, 2048u, 128u);
          const uint64_t dal = b2tc::make_desc(al + ks * 4096u, 2048u, 128u);
          const uint64_t dbh = b2tc::make_desc(bh + ks * 2u * lbo_b, lbo_b, 128u);
          const uint64_t dbl = b2tc::make_desc(bl + ks * 2u * lbo_b, lbo_b, 128u);
          b2tc::mma_tf32(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
          b2tc::mma_tf32(tmem, dah, dbl, idesc, 1u);
          b2tc::mma_tf32(tmem, dah, dbh, idesc, 1u);
        }
        b2tc::mma_commit(&bar_mma);
      }
      b2tc::mbar_wait(&bar_mma, mma_phase);
      mma_phase ^= 1u;
      b2tc::fence_after_sync();
      // ---- epilogue 3: this thread's row of the U table: AoS (UST floats per atom) or, for the
      //      few-type kernels, float4 planes [(t2*KP4 + q)*n + i] (row offset 4*(t2*KP4+q)) ----
      float* Urow = P.U + (size_t)(i >= 0 ? i : 0) * P.UST;
      float4* Upl = reinterpret_cast<float4*>(P.U) + (i >= 0 ? i : 0);
      for (int c0 = 0; c0 < N3; c0 += 16) {
        uint32_t v[16];
        b2tc::tmem_ld16(tmem + lane_base + (uint32_t)c0, v);
        if (i >= 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c0 + 4 * g < P.UST) { // UST = nt*KP is a multiple of 4
              const float4 val =
                make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                            __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
              if (P.u_planes == 2) {
                // compact planes: row offset 4*(t2*KP4 + q); q < KP4-1 -> float4 plane t2*KQ + q, the last
                // chunk of a type holds k = K1-1 in .x -> float plane t2
                const int chunk = c0 / 4 + g, KP4 = P.KP / 4, KQ = KP4 - 1;
                const int t2 = chunk / KP4, q = chunk - t2 * KP4;
                if (q < KQ)
                  Upl[(size_t)(t2 * KQ + q) * N] = val;
                else
                  (P.U + (size_t)P.nt * KQ * 4 * N)[(size_t)t2 * N + (i >= 0 ? i : 0)] = val.x;
              } else if (P.u_planes)
                Upl[(size_t)(c0 / 4 + g) * N] = val;
              else
                *reinterpret_cast<float4*>(Urow + c0 + 4 * g) = val;
            }
          }
        }
      }
    }
    // the next tile's GEMM 1 overwrites Z (columns [0, HN)) only after the barrier in its staging
    // phase, i.e. after every thread's epilogue-1 loads above; its GEMM 2 after the next barrier
  }
  b2tc::fence_before_sync();
  __syncthreads();
  if (warp == 0)
    b2tc::tmem_dealloc(tmem, ncols);
}

} // namespace b2
