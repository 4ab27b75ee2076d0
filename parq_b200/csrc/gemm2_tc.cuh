// CTA-pair (cta_group::2) variant of the persistent tcgen05 GEMM of gemm_tc.cuh:
//
//   D[M,N] = sum_t  A_t[M,K] * B_t[N,K]^T          (bf16 operands, fp32 accumulate in TMEM)
//
// Two CTAs of a cluster (the two SMs of a TPC) compute one 256 x 256 output tile: each CTA holds 128 rows of A and
// 128 of the 256 rows of B per k-block, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) and every CTA
// receives its own 128 x 256 accumulator half in its own TMEM.  Per k-block a CTA therefore moves and re-reads
// 32 KB of operands instead of 48 KB: the single-CTA kernel needs 96 KB of shared-memory traffic (TMA writes + MMA
// reads) per 512 tensor cycles -- 187 B/clk against the 128 B/clk of an SM, i.e. it is shared-memory bound at ~68 % of
// the tensor peak -- the pair needs 64 KB (125 B/clk).  L2 -> SM operand traffic drops by the same third.
//
// Protocol (both CTAs run the same code; r = %cluster_ctarank, leader = rank 0):
//   TMA producer (each CTA)  waits its own empty[stage]; the leader arms its full[stage] with the bytes of BOTH CTAs;
//                            both CTAs issue cta_group::2 TMA loads that complete on the LEADER's full[stage]
//   MMA issuer (leader only) waits full[stage], issues the M = 256 MMAs, commits with a 2-CTA multicast: empty[stage]
//                            of both CTAs, and tfull[acc] of both CTAs at the end of a tile
//   epilogue (each CTA)      waits its own tfull[acc], drains its TMEM half (shared code with gemm_tc.cuh), arrives on
//                            the LEADER's tempty[acc] (count 256)
// Epilogue features, the term / dual-A machinery, the constant-operand prefetch before the PDL wait and the TMA-staged
// channels-first addend are those of gemm_tc.cuh.
#pragma once
#include "gemm_tc.cuh"

namespace parq {

namespace gemm2 {
constexpr int BM = 128;                 // rows of A per CTA (256 per pair)
constexpr int BN = 256;                 // columns of the tile; each CTA stages BN/2 rows of B
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;    // 16 KB
constexpr int BH_BYTES = (BN / 2) * BK * 2;   // 16 KB: this CTA's half of the B tile
constexpr int RING_BYTES = 192 * 1024;
constexpr int MAX_STAGES = 6;
constexpr int THREADS = 256;
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256 + 2 * BN * 4 + 64 + 4 * 32 * 33 * 4;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> the leader's copy
}  // namespace gemm2

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* desc, uint64_t* leader_bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(leader_bar) & gemm2::PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once all previously issued MMAs of the pair retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & gemm2::PEER_MASK) : "memory");
}

template <bool kNchw>
__global__ void __launch_bounds__(gemm2::THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using namespace gemm2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES);   // [8] (leader's copy is the live one)
  uint64_t* empty_bar = full_bar + 8;                                    // [8]
  uint64_t* tfull_bar = empty_bar + 8;                                   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                                  // [2] (leader's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* add_full = tempty_bar + 3;                                   // [3]
  uint64_t* add_empty = add_full + 3;                                    // [3]
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);   // [2][BN]
  uint32_t* sstage = reinterpret_cast<uint32_t*>(sbias + 2 * BN) + 16;                   // [4 warps][32][33], after the GN sums

  const bool dual = p.dual_a != 0;
  const bool add_tma = kNchw && p.ep.add_tma != 0 && p.ep.nchw_add != nullptr && !dual;
  // ring: plain stage = [A | B half] 32 KB (6 stages; 4 when three 16 KB addend buffers share the ring),
  //       dual-A stage = [A_hi | A_lo | B half] 48 KB (4 stages)
  const int stage_bytes = dual ? 2 * A_BYTES + BH_BYTES : A_BYTES + BH_BYTES;
  const int nst = dual ? 4 : (add_tma ? 4 : MAX_STAGES);
  auto a_ptr = [&](int st, int which) { return smem + st * stage_bytes + which * A_BYTES; };
  auto b_ptr = [&](int st) { return smem + st * stage_bytes + (dual ? 2 : 1) * A_BYTES; };
  auto add_buf = [&](int i) { return reinterpret_cast<float*>(smem + 4 * (A_BYTES + BH_BYTES) + i * 16384); };
  auto tile_fast = [&](int m0) { return add_tma && (m0 % p.ep.nchw_HW) + BM <= p.ep.nchw_HW; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 256);     // the epilogue threads of both CTAs
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&add_full[i], 1);
      mbar_init(&add_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                         // CTA-level ordering of the TMEM-address write (also what racecheck models)
  cluster_sync_all();                      // barriers of both CTAs initialised, TMEM of both allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pair tiles: 256 rows x 256 columns; the dimension with fewer tiles varies fastest (operand sharing through L2)
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_tiles = tiles_m * tiles_n;
  const bool m_fastest = tiles_m <= tiles_n;
  const int pair = blockIdx.x >> 1, npair = gridDim.x >> 1;
  auto tile_origin = [&](int tile, int& m0, int& n0) {      // origin of THIS CTA's 128-row half
    int tm, tn;
    if (m_fastest) { tm = tile % tiles_m; tn = tile / tiles_m; }
    else           { tm = tile / tiles_n; tn = tile % tiles_n; }
    m0 = tm * 2 * BM + static_cast<int>(rank) * BM;
    n0 = tn * BN;
  };
  const int kb_per_term = p.K / BK;
  const int nterm_loops = dual ? 1 : p.nterms;
  const int num_kb = kb_per_term * nterm_loops;

  if (warp == 0) {
    if (lane == 0) {                       // ---------------- TMA producer (both CTAs)
      auto a_off = [&](int t) { return t == 0 ? p.a_koff[0] : (t == 1 ? p.a_koff[1] : p.a_koff[2]); };
      auto b_off = [&](int t) { return t == 0 ? p.b_koff[0] : (t == 1 ? p.b_koff[1] : p.b_koff[2]); };
      const uint32_t pair_tx = 2u * static_cast<uint32_t>(stage_bytes);
      int pre = 0;
      if (p.const_operand != 0 && pair < num_tiles) {
        int m0, n0;
        tile_origin(pair, m0, n0);
        pre = num_kb < nst ? num_kb : nst;
        for (int i = 0; i < pre; ++i) {
          const int t = i / kb_per_term, kb = i % kb_per_term;
          if (leader) mbar_expect_tx(&full_bar[i], pair_tx);
          if (p.const_operand == 1) {
            tma_load_2d_pair(a_ptr(i, 0), &tmA, &full_bar[i], a_off(t) + kb * BK, m0);
            if (dual) tma_load_2d_pair(a_ptr(i, 1), &tmA, &full_bar[i], a_off(1) + kb * BK, m0);
          } else {
            tma_load_2d_pair(b_ptr(i), &tmB, &full_bar[i], b_off(t) + kb * BK, n0 + static_cast<int>(rank) * (BN / 2));
          }
        }
      }
      pdl_wait();
      pdl_launch_dependents();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npair) {
        int m0, n0;
        tile_origin(tile, m0, n0);
        const int asplit = (p.a_split_n > 0 && n0 >= p.a_split_n) ? p.a_split_off : 0;
        for (int t = 0; t < nterm_loops; ++t) {
          const int ak = a_off(t) + asplit, bk = b_off(t);
          for (int kb = 0; kb < kb_per_term; ++kb) {
            const bool prefetched = pre > 0;
            if (!prefetched) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (leader) mbar_expect_tx(&full_bar[stage], pair_tx);
            } else {
              --pre;
            }
            if (!(prefetched && p.const_operand == 1)) {
              tma_load_2d_pair(a_ptr(stage, 0), &tmA, &full_bar[stage], ak + kb * BK, m0);
              if (dual) tma_load_2d_pair(a_ptr(stage, 1), &tmA, &full_bar[stage], a_off(1) + asplit + kb * BK, m0);
            }
            if (!(prefetched && p.const_operand == 2))
              tma_load_2d_pair(b_ptr(stage), &tmB, &full_bar[stage], bk + kb * BK, n0 + static_cast<int>(rank) * (BN / 2));
            if (++stage == nst) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else {
      pdl_wait();
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    pdl_wait();
    pdl_launch_dependents();
    if (lane == 0 && leader) {             // ---------------- MMA issuer (leader CTA only)
      constexpr uint32_t idesc = umma_idesc(2 * BM, BN, 1);
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      for (int tile = pair; tile < num_tiles; tile += npair, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tempty_bar[acc], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(a_ptr(stage, 0)));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(b_ptr(stage)));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          if (dual) {
            const uint64_t adesc1 = umma_desc_sw128(smem_u32(a_ptr(stage, 1)));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss_pair(d_tmem, adesc1 + 2 * k, bdesc + 2 * k, idesc, 1u);
          }
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {                  // ---------------- epilogue warps (both CTAs)
    pdl_wait();
    pdl_launch_dependents();
    const int q = warp - 4;
    const int et = threadIdx.x - 128;
    const bool col_bias = (p.ep.bias != nullptr) && !p.ep.bias_per_row;
    int lt = 0;
    int add_cc = 0;
    for (int tile = pair; tile < num_tiles; tile += npair, ++lt) {
      int m0, n0;
      tile_origin(tile, m0, n0);
      const int acc = lt & 1;
      const long long row0 = m0 + q * 32;
      const long long row = row0 + lane;
      uint32_t* stage = sstage + q * (32 * 33);
      float row_bias = 0.f;
      if (col_bias) {
        float* sb = sbias + acc * BN;
        sb[et] = (n0 + et < p.N) ? __ldg(p.ep.bias + n0 + et) : 0.f;
        sb[et + 128] = (n0 + et + 128 < p.N) ? __ldg(p.ep.bias + n0 + et + 128) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      } else if (p.ep.bias != nullptr && row < p.M) {
        row_bias = __ldg(p.ep.bias + row);
      }
      mbar_wait(&tfull_bar[acc], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      uint32_t r0[32], r1[32];
      float gsum = 0.f, gsq = 0.f;
      const bool fast = tile_fast(m0);
      tmem_ld32(taddr, r0);
#pragma unroll 1
      for (int c = 0; c < BN / 32; c += 2) {
        tmem_wait_ld();
        tmem_ld32(taddr + (c + 1) * 32, r1);
        int col0 = n0 + c * 32;
        if (col0 < p.N) {
          const float* sadd = nullptr;
          if (fast) { mbar_wait(&add_full[add_cc % 3], (add_cc / 3) & 1); sadd = add_buf(add_cc % 3); }
          gemm_store_chunk<kNchw>(p.ep, r0, col_bias ? sbias + acc * BN + c * 32 : nullptr, row_bias, row0, lane, col0, p.M, p.N, stage, gsum, gsq, sadd, et);
          if (fast) { mbar_arrive(&add_empty[add_cc % 3]); ++add_cc; }
        }
        tmem_wait_ld();
        if (c + 2 < BN / 32) tmem_ld32(taddr + (c + 2) * 32, r0);
        col0 += 32;
        if (col0 < p.N) {
          const float* sadd = nullptr;
          if (fast) { mbar_wait(&add_full[add_cc % 3], (add_cc / 3) & 1); sadd = add_buf(add_cc % 3); }
          gemm_store_chunk<kNchw>(p.ep, r1, col_bias ? sbias + acc * BN + (c + 1) * 32 : nullptr, row_bias, row0, lane, col0, p.M, p.N, stage, gsum, gsq, sadd, et);
          if (fast) { mbar_arrive(&add_empty[add_cc % 3]); ++add_cc; }
        }
      }
      tc_fence_before();
      mbar_arrive_leader(&tempty_bar[acc]);       // this CTA's accumulator half is drained
      if (p.ep.gn_out != nullptr) {
        double ds = gsum, dq = gsq;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ds += __shfl_xor_sync(0xffffffffu, ds, o);
          dq += __shfl_xor_sync(0xffffffffu, dq, o);
        }
        double* sg = reinterpret_cast<double*>(sbias + 2 * BN);      // [4][2]
        if (lane == 0) { sg[2 * q] = ds; sg[2 * q + 1] = dq; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0)
          p.ep.gn_out[static_cast<long long>(m0 / BM) * p.ep.gn_stride + n0 / BN] =
              make_double2((sg[0] + sg[2]) + (sg[4] + sg[6]), (sg[1] + sg[3]) + (sg[5] + sg[7]));
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }

  if (warp == 2 || warp == 3) {
    pdl_wait();
    pdl_launch_dependents();
    if (add_tma && warp == 2 && lane == 0) {
      int cc = 0;
      for (int tile = pair; tile < num_tiles; tile += npair) {
        int m0, n0;
        tile_origin(tile, m0, n0);
        if (!tile_fast(m0)) continue;
        const int bt = m0 / p.ep.nchw_HW, pix0 = m0 - bt * p.ep.nchw_HW;
        for (int c = 0; c < BN / 32 && n0 + c * 32 < p.N; ++c, ++cc) {
          const int buf = cc % 3;
          mbar_wait(&add_empty[buf], ((cc / 3) & 1) ^ 1);
          mbar_expect_tx(&add_full[buf], p.ep.nchw_add_bf16 ? 32 * 128 * 2 : 32 * 128 * 4);
          tma_load_2d(add_buf(buf), &tmC, &add_full[buf], pix0, bt * p.N + n0 + c * 32);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                      // neither CTA leaves while its peer may still signal its barriers / read its TMEM
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace parq
