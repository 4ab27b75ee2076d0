// CTA-pair (cta_group::2) variant of the flash attention kernel of attn_tc.cuh.
//
// The two 128-query tiles of a (clip, head, key split) form a cluster of two CTAs: the leader issues M = 256 MMAs
// (S = Q K^T with N = 128 keys, O += P V with N = 256 channels), each CTA keeps its own 128 rows of S / P / O in its
// own TMEM and runs its own softmax warps.  Every operand tile is split across the pair -- a CTA stages 64 of the 128
// keys of K and 128 of the 256 channels of V^T -- so per key tile a CTA moves and re-reads 192 KB of shared memory
// instead of 320 KB (the single-CTA kernel needs 156 B/clk against the 128 B/clk an SM has: it is shared-memory bound
// at ~80 % of the tensor pipe) and K / V^T cross the L2 -> SM fabric once per 256 queries instead of twice.
//
// Barriers: kv_full / q_full / p_full live in the leader (armed with the bytes / arrivals of both CTAs);
// kv_empty, s_full, pv_done are signalled in both CTAs by multicast tcgen05.commit.  Everything else (online softmax,
// lazy rescale, split-KV partials, direct output, PDL prologue) is the single-CTA kernel's code.
#pragma once
#include "attn_tc.cuh"
#include "gemm2_tc.cuh"

namespace parq {

// D[tmem] (+)= A[tmem] * B[smem]^T over the CTA pair
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool kFp16>
__global__ void __launch_bounds__(attn::THREADS, 1)
attn2_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using namespace attn;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* ring = smem + Q_BYTES;
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(ring + NS * STAGE_BYTES);
  uint64_t* kv_empty = kv_full + NS;
  uint64_t* q_full = kv_empty + NS;
  uint64_t* s_full = q_full + 1;     // [2]
  uint64_t* p_full = s_full + 2;     // [2]
  uint64_t* pv_done = p_full + 2;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // == blockIdx.x & 1: which 128-query half of the pair's 256 queries
  const bool leader = rank == 0;
  const int split = blockIdx.y;                 // grid = (query tiles, key splits, clips x heads), cluster = (2, 1, 1)
  const int qt = blockIdx.x;
  const int bh = blockIdx.z;
  const int b = bh / p.H, h = bh % p.H;

  const int ntiles = (p.Nk + BKEY - 1) / BKEY;
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(ntiles, t0 + p.tiles_per_split);
  const int n = t1 - t0;                 // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);     // the softmax threads of both CTAs (the leader's copy is the live one)
    }
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();                     // CTA-level ordering of the TMEM-address write (also what racecheck models)
  cluster_sync_all();                  // barriers of both CTAs initialised, TMEM of both allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0 && n > 0) {             // ---------------- TMA producer
      const int ch0 = h * DH;
      int stage = 0;
      uint32_t phase = 0;
      // Each CTA stages HALF of every operand tile of the pair's M = 256 MMAs, one 32 KB ring stage each:
      //   K : this CTA's 64 of the tile's 128 keys, all 256 channels  (4 boxes of 64 keys x 64 channels)
      //   V^T: this CTA's 128 of the head's 256 channels, all 128 keys (2 boxes of 128 channels x 64 keys)
      // All loads complete on the LEADER's kv_full[stage], which its producer arms with the bytes of both CTAs.
      auto load_k = [&](int tile) {
        const int row = (p.kv_tiled ? ((b * p.ntile + tile) * p.H + h) * BKEY : b * p.Nk + tile * BKEY) + static_cast<int>(rank) * (BKEY / 2);
        const int col = p.kv_tiled ? 0 : ch0;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&kv_full[stage], 2 * STAGE_BYTES);
        uint8_t* dst = ring + stage * STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d_pair(dst + c * 8192, &tmK, &kv_full[stage], col + c * 64, row);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      };
      auto load_v = [&](int tile) {
        const int col = p.kv_tiled ? 0 : b * p.Nk + tile * BKEY;
        const int row = (p.kv_tiled ? ((b * p.ntile + tile) * p.H + h) * DH : ch0) + static_cast<int>(rank) * (DH / 2);
        mbar_wait(&kv_empty[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&kv_full[stage], 2 * STAGE_BYTES);
        uint8_t* dst = ring + stage * STAGE_BYTES;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) tma_load_2d_pair(dst + kc * 16384, &tmV, &kv_full[stage], col + kc * 64, row);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      };
      // ring order: K(0), [K(j+1), V(j)] ...  With an old (cached) K the first two key tiles are requested
      // before the programmatic-dependency wait; Q comes from the previous kernel and follows the wait.
      const bool early = p.kv_const != 0;
      if (early) {
        load_k(t0);
        if (n > 1) load_k(t0 + 1);
      }
      pdl_wait();
      pdl_launch_dependents();
      if (leader) mbar_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_2d_pair(sQ + c * (BQ * 128), &tmQ, q_full, ch0 + c * 64, b * p.Nq + qt * BQ);
      if (!early) load_k(t0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n && !(early && j == 0)) load_k(t0 + j + 1);
        load_v(t0 + j);
      }
    } else {
      pdl_wait();
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    pdl_wait();
    pdl_launch_dependents();
    if (lane == 0 && n > 0 && leader) {   // ---------------- MMA issuer (leader CTA only, M = 256 over the pair)
      constexpr uint32_t fmt = kFp16 ? 0u : 1u;
      constexpr uint32_t idesc_s = umma_idesc(2 * BQ, BKEY, fmt);
      constexpr uint32_t idesc_pv = umma_idesc(2 * BQ, DH, fmt);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](int buf) {
        const uint32_t d_tmem = tmem_base + buf * 128;
        mbar_wait(&kv_full[stage], phase);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(ring + stage * STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) {            // 64-channel chunks: Q box c against K box c (64 keys of this CTA)
          const uint64_t qd = umma_desc_sw128(q_addr + c * (BQ * 128));
          const uint64_t kd = umma_desc_sw128(k_addr + c * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_pair(d_tmem, qd + 2 * k, kd + 2 * k, idesc_s, (c | k) != 0);
        }
        umma_commit_pair(&kv_empty[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
        umma_commit_pair(&s_full[buf]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) issue_s((j + 1) & 1);
        const int buf = j & 1;
        mbar_wait(&p_full[buf], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t p_tmem = tmem_base + buf * 128;
        mbar_wait(&kv_full[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {          // 64-key chunks of this CTA's 128 channels
          const uint64_t vd = umma_desc_sw128(smem_u32(ring + stage * STAGE_BYTES + kc * 16384));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts_pair(tmem_O, p_tmem + kc * 32 + k * 8, vd + 2 * k, idesc_pv, (j | kc | k) != 0);
        }
        umma_commit_pair(&kv_empty[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
        umma_commit_pair(pv_done);
      }
    }
  } else if (warp >= 4 && n > 0) {        // ---------------- softmax / correction / epilogue
    pdl_wait();
    pdl_launch_dependents();
    const int q = warp - 4;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n; ++j) {
      const int buf = j & 1;
      mbar_wait(&s_full[buf], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_tmem = tmem_base + lane_base + buf * 128;
      attn_softmax_tile<kFp16>(s_tmem, tmem_O + lane_base, p.Nk - (t0 + j) * BKEY, j == 0, pv_done, (j - 1) & 1, m_run, l_run,
                               [&] { mbar_arrive_leader(&p_full[buf]); });
    }
    // epilogue
    mbar_wait(pv_done, (n - 1) & 1);
    tc_fence_after();
    if (p.out_direct != nullptr) {
      // single split: normalise here and emit the [hi|lo] operand of the out-projection directly
      const int C = p.H * DH;
      attn_store_normalised(tmem_O + lane_base, 1.f / l_run, p.out_direct + (static_cast<long long>(b) * p.Nq + qt * BQ + q * 32 + lane) * (2 * C) + h * DH, C);
    } else {
      // un-normalised O, m, l of this split for attn_combine_kernel
      const long long part = (static_cast<long long>(bh) * p.nsplit + split) * p.Nq + qt * BQ + q * 32 + lane;
      attn_store_partial(tmem_O + lane_base, p.o_part + part * DH);
      p.ml_part[part] = make_float2(m_run, l_run);
    }
  }

  if (warp == 2 || warp == 3 || n <= 0) {
    pdl_wait();
    pdl_launch_dependents();
  }
  tc_fence_before();
  cluster_sync_all();                  // neither CTA leaves while its peer may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace parq
