// Cluster split-K tcgen05 GEMM for products with a FEW ROW TILES (one clip: B*Nq = 256 rows), sm_100a.
//
//   D[M,N] = A_hi[M,K] W[N,K]^T + A_lo[M,K] W[N,K]^T          (the [hi|lo] activation split of gemm_tc.cuh, bf16-exact weights)
//
// With two row tiles a linear layer of the decoder cannot fill the machine by tiling the output alone: gemm_tc.cuh's narrow
// 128 x 64 tiles give 32 CTAs for N = 1024, and every one of them still streams its whole 128 x 2K A tile (512 KB) plus
// 256 KB of weights through ONE TMA unit -- ncu (warm L2): 28 k cycles per launch with the tensor pipe 20 % active, i.e. the
// kernel is the operand feed of a single SM (~38 B/clk for 128-byte rows of a pitched matrix), not the MMAs.  Here a
// CLUSTER OF 4 CTAs owns one 128 x 128 output tile and splits K: CTA r loads and multiplies k-blocks [r K/4, (r+1) K/4)
// (at most four 48 KB stages, all requested up front: no ring, no empty barriers), so four TMA units feed a tile and the
// k-loop is a quarter as long.  The partial accumulators are then reduce-SCATTERED over distributed shared memory: CTA r
// keeps the 32 output columns [32 r, 32 r + 32) of the tile, receives the other three CTAs' partials for them (stored
// column-major, conflict free, into the -- by then idle -- first operand stage) and runs the common epilogue
// (gemm_store_chunk of gemm_tc.cuh: bias / ReLU / fp32, bf16 / fp16, [hi|lo] outputs, GroupNorm tile sums) on its chunk.
// Two cluster barriers per launch (all MMAs retired -> partials exchanged) replace three quarters of the k-loop.
//
// Used for every activation x weight GEMM of the un-chained launch path whose shape fits (K / 64 a multiple of 4 and at
// most 16, N a multiple of 128, all clusters resident in one wave -- beyond that the output tiles alone fill the machine):
// reference transformer_parq.py:176-180, 365-386, generic_mlp.py:94-110 at one clip.  launch_gemm (parq_api.cu) selects it.
#pragma once
#include "chain_tc.cuh"
#include "gemm_tc.cuh"

namespace parq {

namespace gemmsk {
constexpr int S = 4;                      // CTAs per cluster = K split
constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 64;
constexpr int MAX_STEPS = 4;              // k-blocks per CTA
constexpr int A_BYTES = BM * BK * 2;      // 16 KB
constexpr int B_BYTES = BN * BK * 2;      // 16 KB
constexpr int SLOT_BYTES = 2 * A_BYTES + B_BYTES;   // [A_hi | A_lo | W] of one k-block
constexpr int THREADS = 192;              // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (TMEM quadrant = warp % 4; warp 2 allocates TMEM)
constexpr int SMEM_BYTES = 1024 /*align slack*/ + MAX_STEPS * SLOT_BYTES + 256 /*barriers*/ + 32 * 4 /*bias*/ + 64 /*GroupNorm sums*/ +
                           4 * 32 * 33 * 4 /*per-warp store staging*/;
static_assert((S - 1) * 32 * BM * 4 <= SLOT_BYTES, "the received partials fit into the first operand stage");
}  // namespace gemmsk

__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// grid = 4 * (M / 128 rounded up) * (N / 128) CTAs in clusters of 4 along x
__global__ void __launch_bounds__(gemmsk::THREADS, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using namespace gemmsk;
  extern __shared__ uint8_t smem_raw_sk[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_sk) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + MAX_STEPS * SLOT_BYTES);   // [MAX_STEPS]
  uint64_t* tfull_bar = full_bar + MAX_STEPS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  float* sbias = reinterpret_cast<float*>(smem + MAX_STEPS * SLOT_BYTES + 256);      // [32]
  double* sgn = reinterpret_cast<double*>(sbias + 32);                               // [4][2]
  uint32_t* sstage = reinterpret_cast<uint32_t*>(sgn + 8);                           // [4][32 * 33]
  float* recv = reinterpret_cast<float*>(smem);                                      // [S - 1][32 columns][128 rows], after all MMAs

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x / S;
  const int tiles_n = p.N / BN;
  const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
  const int steps = (p.K / BK) / S;                    // k-blocks of this CTA
  const int kb0 = static_cast<int>(rank) * steps;
  const int asplit = (p.a_split_n > 0 && n0 >= p.a_split_n) ? p.a_split_off : 0;

  if (warp == 0) {                         // (every lane: no divergence in front of the block-wide barrier below)
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAX_STEPS; ++i) mbar_init(&full_bar[i], 1);
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncwarp();
  asm volatile("bar.sync 2, 192;" ::: "memory");      // == gemmsk::THREADS
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto a_ptr = [&](int i, int which) { return smem + i * SLOT_BYTES + which * A_BYTES; };
  auto b_ptr = [&](int i) { return smem + i * SLOT_BYTES + 2 * A_BYTES; };

  if (warp == 0) {
    if (lane == 0) {                       // ---------------- TMA producer: weights before the dependency wait, activations after
      const bool w_const = p.const_operand == 2;          // the B operand holds constants: fetched while the predecessor drains
      for (int i = 0; i < steps; ++i) {
        mbar_expect_tx(&full_bar[i], SLOT_BYTES);
        if (w_const) tma_load_2d(b_ptr(i), &tmB, &full_bar[i], p.b_koff[0] + (kb0 + i) * BK, n0);
      }
      pdl_wait();
      pdl_launch_dependents();
      for (int i = 0; i < steps; ++i) {
        if (!w_const) tma_load_2d(b_ptr(i), &tmB, &full_bar[i], p.b_koff[0] + (kb0 + i) * BK, n0);
        tma_load_2d(a_ptr(i, 0), &tmA, &full_bar[i], p.a_koff[0] + asplit + (kb0 + i) * BK, m0);
        tma_load_2d(a_ptr(i, 1), &tmA, &full_bar[i], p.a_koff[1] + asplit + (kb0 + i) * BK, m0);
      }
    } else {
      pdl_wait();
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    pdl_wait();
    pdl_launch_dependents();
    if (lane == 0) {                       // ---------------- MMA issuer
      const uint32_t idesc = umma_idesc(BM, BN, 1);
      for (int i = 0; i < steps; ++i) {
        mbar_wait(&full_bar[i], 0);
        tc_fence_after();
        const uint64_t adesc0 = umma_desc_sw128(smem_u32(a_ptr(i, 0)));
        const uint64_t adesc1 = umma_desc_sw128(smem_u32(a_ptr(i, 1)));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(b_ptr(i)));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_base, adesc0 + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_base, adesc1 + 2 * k, bdesc + 2 * k, idesc, 1u);
      }
      umma_commit(tfull_bar);
    }
  } else {                                 // ---------------- epilogue warps: bias of this CTA's 32 columns, then the accumulator
    pdl_wait();
    pdl_launch_dependents();
    const int et = threadIdx.x - 64;       // 0..127
    if (et < 32) sbias[et] = p.ep.bias != nullptr ? __ldg(p.ep.bias + n0 + static_cast<int>(rank) * 32 + et) : 0.f;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
  }
  __syncwarp();
  cluster_sync_all();                      // every CTA's MMAs have retired: operand stage 0 of all four may be overwritten

  const int q = warp & 3;                  // TMEM lane quadrant of an epilogue warp
  const int rowin = q * 32 + lane;
  float mine[32];
  if (warp >= 2) {
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < S; ++c) {
      tmem_ld32(taddr + c * 32, r);
      tmem_wait_ld();
      if (c == static_cast<int>(rank)) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mine[i] = __uint_as_float(r[i]);
      } else {
        // columns [32 c, 32 c + 32) belong to CTA c: slot (this rank, skipping c itself), column-major -> every store of a warp
        // is 32 consecutive rows = one 128-byte line of the peer's shared memory
        const int src = static_cast<int>(rank) < c ? static_cast<int>(rank) : static_cast<int>(rank) - 1;
        const uint32_t dst = mapa_u32(smem_u32(recv + (src * 32) * BM + rowin), static_cast<uint32_t>(c));
#pragma unroll
        for (int i = 0; i < 32; ++i) st_cluster_f32(dst + i * BM * 4, __uint_as_float(r[i]));
      }
    }
  }
  __syncwarp();
  cluster_sync_all();                      // partials exchanged (release / acquire at cluster scope); no remote access after this

  if (warp >= 2) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float v = mine[i];
#pragma unroll
      for (int s = 0; s < S - 1; ++s) v += recv[(s * 32 + i) * BM + rowin];       // fixed order: deterministic
      r[i] = __float_as_uint(v);
    }
    float gsum = 0.f, gsq = 0.f;
    const long long row0 = m0 + q * 32;
    const int col0 = n0 + static_cast<int>(rank) * 32;
    asm volatile("bar.sync 1, 128;" ::: "memory");      // sbias written by the first epilogue warp
    gemm_store_chunk<false>(p.ep, r, p.ep.bias != nullptr ? sbias : nullptr, 0.f, row0, lane, col0, p.M, p.N, sstage + q * (32 * 33), gsum, gsq);
    if (p.ep.gn_out != nullptr) {
      double ds = gsum, dq = gsq;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
      }
      if (lane == 0) { sgn[2 * q] = ds; sgn[2 * q + 1] = dq; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64)
        p.ep.gn_out[static_cast<long long>(m0 / BM) * p.ep.gn_stride + col0 / GN_SLOT_COLS] =
            make_double2((sgn[0] + sgn[2]) + (sgn[4] + sgn[6]), (sgn[1] + sgn[3]) + (sgn[5] + sgn[7]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace parq
