// Stream-K scheduling of the CTA-pair flash attention (attn2_tc.cuh) for the cross-attention.
//
// attn2 launches (key splits x query pairs x clips x heads) CTAs -- 1024 at config 2, 6.9 waves of 148 -- and every CTA
// pays its own prologue (barrier init, 2-CTA TMEM allocation, two cluster barriers, Q load) and its own epilogue (128 KB
// of partial O), and the merge kernel then re-reads 134 MB of partials.  Here the grid is ONE CTA pair per SM pair.
// The work is the flat list of (item, key tile) units, item = (clip, head, 256-query pair); pair p owns the contiguous
// unit range [U p / P, U (p+1) / P): perfectly balanced, one prologue per SM, and an item is cut only where a range
// ends, so there are at most P + items partial outputs (148 x 256 rows instead of 1024 x 128) -- items that fall
// completely inside one range are normalised and written directly.
//
// A pair walks its range as SEGMENTS (one per item it touches).  The K / V^T ring, the S / P double buffer and the
// pv_done phases simply continue across segments (a global tile counter g supplies buffer indices and parities); per
// segment the producer reloads Q once the previous segment's MMAs have retired (q_empty), the first P.V of a segment
// overwrites O only after the previous segment's epilogue has drained it (o_free), and the softmax warps restart
// their running max / sum.  Everything else is attn2_tc.cuh.
#pragma once
#include "attn2_tc.cuh"

namespace parq {

struct Attn3Params {
  int B, H, Nq, Nk;
  int qpairs;              // Nq / 256
  int ntiles;              // key tiles per item
  long long units;         // items * ntiles
  int npairs;              // CTA pairs in the grid (== gridDim.y)
  int slots;               // partial slots per pair
  float* o_part;           // [(pair*slots + seg)*256 + row][256]
  float2* ml_part;         // [(pair*slots + seg)*256 + row] = (m, l)
  __nv_bfloat16* out;      // (B*Nq, 2*H*256) [hi|lo]: direct output of whole-item segments
  int kv_const, kv_tiled;
  uint32_t* flags;         // fused merge (items cut into <= 3 pieces), or null: [(pair*2 + rank)*4 + softmax warp], all zero before and
                           // after a launch.  A piece that does not start its item is always the FIRST segment of its pair; its four
                           // softmax warps publish their rows with flag = 1.  The pair that holds the item's first piece finishes it at
                           // the very end of its range: it waits for those flags, folds the other pieces into its own O (still in
                           // TMEM), writes the final operand and clears the flags -- no attn3_combine_kernel launch.
};

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* ptr) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t* ptr, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

__device__ __forceinline__ long long sk_unit_begin(long long units, int npairs, int pair) {
  return units * pair / npairs;
}

template <bool kFp16>
__global__ void __launch_bounds__(attn::THREADS, 1)
attn3_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const Attn3Params p) {
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
  uint64_t* q_empty = pv_done + 1;   // [1] all MMAs of a segment retired: Q may be overwritten
  uint64_t* o_free = q_empty + 1;    // [1] (leader's copy is live) the segment's O has been read by both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // == blockIdx.x: which 128-query half of the item's 256 queries
  const bool leader = rank == 0;
  // ranges are handed out in REVERSE block order: the pair that merges an item waits for pairs with higher range indices, i.e.
  // for blocks with lower indices -- blocks that are scheduled first even if the grid were ever not fully resident
  const int pair = p.npairs - 1 - static_cast<int>(blockIdx.y);
  const long long u0 = sk_unit_begin(p.units, p.npairs, pair), u1 = sk_unit_begin(p.units, p.npairs, pair + 1);

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
      mbar_init(&p_full[i], 256);
    }
    mbar_init(pv_done, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, 256);
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  // segment walk shared by all roles: item, first tile, number of tiles
  struct Seg { int item, ta, n; };
  auto seg_at = [&](long long u) {
    Seg s;
    s.item = static_cast<int>(u / p.ntiles);
    s.ta = static_cast<int>(u - static_cast<long long>(s.item) * p.ntiles);
    const long long left = u1 - u;
    s.n = static_cast<int>(left < p.ntiles - s.ta ? left : p.ntiles - s.ta);
    return s;
  };
  auto item_bh = [&](int item, int& b, int& h, int& qt) {
    const int qp = item % p.qpairs, bh = item / p.qpairs;
    b = bh / p.H;
    h = bh - b * p.H;
    qt = qp * 2 + static_cast<int>(rank);
  };

  if (warp == 0) {
    if (lane == 0) {                      // ---------------- TMA producer (both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      auto load_k = [&](int b, int h, int tile) {
        const int row = (p.kv_tiled ? ((b * p.ntiles + tile) * p.H + h) * BKEY : b * p.Nk + tile * BKEY) + static_cast<int>(rank) * (BKEY / 2);
        const int col = p.kv_tiled ? 0 : h * DH;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&kv_full[stage], 2 * STAGE_BYTES);
        uint8_t* dst = ring + stage * STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d_pair(dst + c * 8192, &tmK, &kv_full[stage], col + c * 64, row);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      };
      auto load_v = [&](int b, int h, int tile) {
        const int col = p.kv_tiled ? 0 : b * p.Nk + tile * BKEY;
        const int row = (p.kv_tiled ? ((b * p.ntiles + tile) * p.H + h) * DH : h * DH) + static_cast<int>(rank) * (DH / 2);
        mbar_wait(&kv_empty[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&kv_full[stage], 2 * STAGE_BYTES);
        uint8_t* dst = ring + stage * STAGE_BYTES;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) tma_load_2d_pair(dst + kc * 16384, &tmV, &kv_full[stage], col + kc * 64, row);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      };
      int segi = 0;
      for (long long u = u0; u < u1; ++segi) {
        const Seg s = seg_at(u);
        int b, h, qt;
        item_bh(s.item, b, h, qt);
        // ring order per segment: K(0), [K(j+1), V(j)] ...  The first two key tiles of the FIRST segment are
        // requested before the programmatic-dependency wait when K / V^T are an old cache; Q follows the wait.
        const bool early = segi == 0 && p.kv_const != 0;
        if (early) {
          load_k(b, h, s.ta);
          if (s.n > 1) load_k(b, h, s.ta + 1);
        }
        if (segi == 0) {
          pdl_wait();
          pdl_launch_dependents();
        } else {
          mbar_wait(q_empty, (segi - 1) & 1);           // the previous segment's MMAs no longer read Q
        }
        if (leader) mbar_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_2d_pair(sQ + c * (BQ * 128), &tmQ, q_full, h * DH + c * 64, b * p.Nq + qt * BQ);
        if (!early) load_k(b, h, s.ta);
        for (int j = 0; j < s.n; ++j) {
          if (j + 1 < s.n && !(early && j == 0)) load_k(b, h, s.ta + j + 1);
          load_v(b, h, s.ta + j);
        }
        u += s.n;
      }
    } else {
      pdl_wait();
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    pdl_wait();
    pdl_launch_dependents();
    if (lane == 0 && leader) {            // ---------------- MMA issuer (leader CTA only, M = 256 over the pair)
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
        for (int c = 0; c < 4; ++c) {
          const uint64_t qd = umma_desc_sw128(q_addr + c * (BQ * 128));
          const uint64_t kd = umma_desc_sw128(k_addr + c * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_pair(d_tmem, qd + 2 * k, kd + 2 * k, idesc_s, (c | k) != 0);
        }
        umma_commit_pair(&kv_empty[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
        umma_commit_pair(&s_full[buf]);
      };
      int g = 0;                          // tiles issued so far by this pair (buffer index / parities)
      int segi = 0;
      for (long long u = u0; u < u1; ++segi) {
        const Seg s = seg_at(u);
        mbar_wait(q_full, segi & 1);
        tc_fence_after();
        issue_s(g & 1);
        for (int j = 0; j < s.n; ++j, ++g) {
          if (j + 1 < s.n) issue_s((g + 1) & 1);
          const int buf = g & 1;
          mbar_wait(&p_full[buf], (g >> 1) & 1);
          if (j == 0 && segi > 0) mbar_wait(o_free, (segi - 1) & 1);    // the previous segment's O has been drained
          tc_fence_after();
          const uint32_t p_tmem = tmem_base + buf * 128;
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int kc = 0; kc < 2; ++kc) {
            const uint64_t vd = umma_desc_sw128(smem_u32(ring + stage * STAGE_BYTES + kc * 16384));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ts_pair(tmem_O, p_tmem + kc * 32 + k * 8, vd + 2 * k, idesc_pv, (j | kc | k) != 0);
          }
          umma_commit_pair(&kv_empty[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
          umma_commit_pair(pv_done);
        }
        u += s.n;
        // every MMA of the segment retired: Q may be reloaded.  Not after the last segment: nobody would wait for it, and
        // a multicast arrival must not be in flight towards a CTA that is about to leave.
        if (u < u1) umma_commit_pair(q_empty);
      }
    }
  } else if (warp >= 4) {                 // ---------------- softmax / correction / epilogue (both CTAs)
    pdl_wait();
    pdl_launch_dependents();
    const int q = warp - 4;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int g = 0;
    int segi = 0;
    for (long long u = u0; u < u1; ++segi) {
      const Seg s = seg_at(u);
      int b, h, qt;
      item_bh(s.item, b, h, qt);
      float m_run = -INFINITY, l_run = 0.f;
      const bool merger = p.flags != nullptr && s.ta == 0 && s.n < p.ntiles;
      for (int j = 0; j < s.n; ++j, ++g) {
        const int buf = g & 1;
        if (merger && j == (s.n > 8 ? s.n - 8 : 0)) {
          // a few tiles before the merge: pull this row of the other pieces (written long ago, evicted by the K / V^T stream) back
          // into L2 -- the merge at the end of the kernel is a chain of dependent round trips on every SM at once
          const long long ub = static_cast<long long>(s.item) * p.ntiles + p.ntiles - 1;
          const int my_row = static_cast<int>(rank) * BQ + q * 32 + lane;
          for (int k = 1; k <= 2; ++k) {
            if (pair + k >= p.npairs || sk_unit_begin(p.units, p.npairs, pair + k) > ub) break;
            const float* src = p.o_part + ((static_cast<long long>(pair + k) * p.slots) * 256 + my_row) * DH;
#pragma unroll
            for (int i = 0; i < DH * 4 / 128; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + i * 32));
          }
        }
        mbar_wait(&s_full[buf], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t s_tmem = tmem_base + lane_base + buf * 128;
        // Every completion of pv_done is observed exactly once and in order (global tile counter g): tile g consumes the
        // completion of P(g-1)V(g-1) before touching O or at its end; for the first tile of a segment that completion was
        // consumed by the previous segment's epilogue.
        attn_softmax_tile<kFp16>(s_tmem, tmem_O + lane_base, p.Nk - (s.ta + j) * BKEY, j == 0, pv_done, (g - 1) & 1, m_run, l_run,
                                 [&] { mbar_arrive_leader(&p_full[buf]); });
      }
      // segment epilogue: g tiles issued so far, the last one is g-1
      mbar_wait(pv_done, (g - 1) & 1);
      tc_fence_after();
      if (s.ta == 0 && s.n == p.ntiles) {
        // the whole item was processed here: normalise and emit the [hi|lo] operand of the out-projection directly
        const int C = p.H * DH;
        attn_store_normalised(tmem_O + lane_base, 1.f / l_run, p.out + (static_cast<long long>(b) * p.Nq + qt * BQ + q * 32 + lane) * (2 * C) + h * DH, C);
      } else if (p.flags != nullptr && s.ta == 0) {
        // fused merge: this pair holds the FIRST piece of the item and finishes it last; the other pieces are the first
        // segments of the following pairs (at most two of them)
        const long long ub = static_cast<long long>(s.item) * p.ntiles + p.ntiles - 1;
        const int my_row = static_cast<int>(rank) * BQ + q * 32 + lane;
        // piece k lives in pair + 1 + k; the second one exists when that pair's range still starts inside the item
        const bool two = pair + 2 < p.npairs && sk_unit_begin(p.units, p.npairs, pair + 2) <= ub;
        const int nk = two ? 2 : 1;
        if (lane == 0) {
          for (int k = 0; k < nk; ++k)
            while (ld_acquire_gpu_u32(p.flags + ((pair + 1 + k) * 2 + static_cast<int>(rank)) * 4 + q) == 0u) __nanosleep(100);
        }
        __syncwarp();
        const long long part0 = (static_cast<long long>(pair + 1) * p.slots) * 256 + my_row;
        const long long part1 = two ? (static_cast<long long>(pair + 2) * p.slots) * 256 + my_row : part0;
        const float2 ml0 = __ldcg(&p.ml_part[part0]);
        const float2 ml1 = two ? __ldcg(&p.ml_part[part1]) : make_float2(-INFINITY, 0.f);
        const float* ok[2] = {p.o_part + part0 * DH, p.o_part + part1 * DH};
        const float M = fmaxf(fmaxf(m_run, ml0.x), ml1.x);
        const float w_me = fast_exp2(m_run - M);
        const float wk[2] = {fast_exp2(ml0.x - M), two ? fast_exp2(ml1.x - M) : 0.f};
        const float L = l_run * w_me + wk[0] * ml0.y + wk[1] * ml1.y;
        const float inv = 1.f / L;
        const float s_me = w_me * inv, s0 = wk[0] * inv, s1 = wk[1] * inv;
        const int C = p.H * DH;
        __nv_bfloat16* dst = p.out + (static_cast<long long>(b) * p.Nq + qt * BQ + q * 32 + lane) * (2 * C) + h * DH;
        float4 a0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a0[i] = __ldcg(reinterpret_cast<const float4*>(ok[0]) + i);
#pragma unroll 1
        for (int c = 0; c < DH / 32; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_O + lane_base + c * 32, o);
          tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i] = __uint_as_float(o[4 * i]) * s_me + a0[i].x * s0;
            v[4 * i + 1] = __uint_as_float(o[4 * i + 1]) * s_me + a0[i].y * s0;
            v[4 * i + 2] = __uint_as_float(o[4 * i + 2]) * s_me + a0[i].z * s0;
            v[4 * i + 3] = __uint_as_float(o[4 * i + 3]) * s_me + a0[i].w * s0;
          }
          if (c + 1 < DH / 32) {                              // the next chunk of the first piece, in flight during the stores below
#pragma unroll
            for (int i = 0; i < 8; ++i) a0[i] = __ldcg(reinterpret_cast<const float4*>(ok[0] + (c + 1) * 32) + i);
          }
          if (two) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a1 = __ldcg(reinterpret_cast<const float4*>(ok[1] + c * 32) + i);
              v[4 * i] += a1.x * s1; v[4 * i + 1] += a1.y * s1; v[4 * i + 2] += a1.z * s1; v[4 * i + 3] += a1.w * s1;
            }
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            hi[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            lo[i] = pack_bf16x2(v[2 * i] - __uint_as_float(hi[i] << 16), v[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            reinterpret_cast<uint4*>(dst + c * 32)[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            reinterpret_cast<uint4*>(dst + C + c * 32)[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
          }
        }
        // the pieces are consumed: flags back to zero for the next launch
        __syncwarp();
        if (lane == 0) {
          for (int k = 0; k < nk; ++k) p.flags[((pair + 1 + k) * 2 + static_cast<int>(rank)) * 4 + q] = 0u;
        }
      } else {
        // un-normalised O, m, l of this segment for the merge (attn3_combine_kernel, or the pair that holds the item's first piece)
        const long long part = (static_cast<long long>(pair) * p.slots + segi) * 256 + static_cast<int>(rank) * BQ + q * 32 + lane;
        attn_store_partial(tmem_O + lane_base, p.o_part + part * DH);
        p.ml_part[part] = make_float2(m_run, l_run);
        if (p.flags != nullptr) {
          __threadfence();
          __syncwarp();
          if (lane == 0) st_release_gpu_u32(p.flags + (pair * 2 + static_cast<int>(rank)) * 4 + q, 1u);
        }
      }
      tc_fence_before();
      mbar_arrive_leader(o_free);         // this CTA's O rows are read: the next segment may overwrite them
      u += s.n;
    }
  }

  if (warp == 2 || warp == 3) {
    pdl_wait();
    pdl_launch_dependents();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// Merge the segments of every item that was cut at a range boundary (items processed by a single pair were written
// directly and are skipped): out = sum_s 2^(m_s-M) O_s / sum_s 2^(m_s-M) l_s as the exact bf16 split [hi | lo].
// One block per (item, 32 of its 256 query rows): the 64-bit range arithmetic that locates the item's pieces runs once
// per block; 64 threads per row, a thread owns 4 consecutive channels.
constexpr int SK_COMBINE_ROWS = 8;       // rows of an item per block (a thread walks rows / 4 of them): 8 -> 2 048 blocks at config 2
constexpr int SK_FLAG_WORDS = 1024;       // Attn3Params::flags: (pair, rank, softmax warp) words, >= 2 * 4 * SK_MAX_PAIRS
constexpr int SK_MAX_PAIRS = 128;        // CTA pairs the schedule may use (74 on a B200)

// kWide (few items cut into many pieces each: one clip on the whole machine): the four 64-thread groups of a block share ONE row
// and take every fourth piece of it, their partial merges meet in shared memory -- a quarter of the dependent loads per thread.
template <bool kWide>
__global__ void __launch_bounds__(256)
attn3_combine_kernel(const Attn3Params p, const int rows) {
  const int item = blockIdx.x / (256 / rows);
  const int r0 = (blockIdx.x % (256 / rows)) * rows;      // first row of this block inside the item
  __shared__ int s_np;
  __shared__ long long s_slot[SK_MAX_PAIRS];                // first row of every piece of the item in o_part / ml_part (<= npairs pieces)
  __shared__ float s_merge[kWide ? 3 * 64 * 6 : 1];         // groups 1..3: (M, L, acc[4]) per thread of the group
  if (threadIdx.x == 0) {
    const long long ua = static_cast<long long>(item) * p.ntiles, ub = ua + p.ntiles - 1;      // first / last unit of the item
    auto pair_of = [&](long long u) {
      int pr = static_cast<int>(u * p.npairs / p.units);
      while (sk_unit_begin(p.units, p.npairs, pr + 1) <= u) ++pr;
      while (sk_unit_begin(p.units, p.npairs, pr) > u) --pr;
      return pr;
    };
    const int pa = pair_of(ua), pb = pair_of(ub);
    s_np = pb - pa + 1;
    for (int pr = pa; pr <= pb; ++pr) {
      const int seg = item - static_cast<int>(sk_unit_begin(p.units, p.npairs, pr) / p.ntiles);
      s_slot[pr - pa] = (static_cast<long long>(pr) * p.slots + seg) * 256;
    }
  }
  // the piece table above depends on the launch parameters only: it is built while the attention kernel drains
  pdl_wait();
  pdl_launch_dependents();
  __syncthreads();
  const int np = s_np;
  if (np == 1) return;                                       // whole item in one range: written directly
  const int qp = item % p.qpairs, bh = item / p.qpairs;
  const int b = bh / p.H, h = bh - b * p.H;
  const int C = p.H * 256;
  const int d = (threadIdx.x & 63) * 4;
  const int grp = threadIdx.x >> 6;
  const int r_first = kWide ? r0 : r0 + grp, r_step = kWide ? 1 : 4;
  const int i_first = kWide ? grp : 0, i_step = kWide ? 4 : 1;
  for (int r = r_first; r < r0 + rows; r += r_step) {
    float M = -INFINITY;
    for (int i = i_first; i < np; i += i_step) M = fmaxf(M, __ldg(&p.ml_part[s_slot[i] + r].x));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float L = 0.f;
    for (int i = i_first; i < np; i += i_step) {
      const long long idx = s_slot[i] + r;
      const float2 ml = __ldg(&p.ml_part[idx]);
      const float4 o = __ldg(reinterpret_cast<const float4*>(p.o_part + idx * 256 + d));
      const float w = exp2f(ml.x - M);
      L += w * ml.y;
      acc.x += w * o.x; acc.y += w * o.y; acc.z += w * o.z; acc.w += w * o.w;
    }
    if (kWide) {
      // merge the four groups' (M, L, acc) of this row in group 0 (a group without pieces carries M = -inf, L = 0)
      float* mine = s_merge + ((grp - 1) * 64 + (threadIdx.x & 63)) * 6;
      if (grp > 0) { mine[0] = M; mine[1] = L; mine[2] = acc.x; mine[3] = acc.y; mine[4] = acc.z; mine[5] = acc.w; }
      __syncthreads();
      if (grp == 0) {
        float Mt = M;
#pragma unroll
        for (int g = 0; g < 3; ++g) Mt = fmaxf(Mt, s_merge[(g * 64 + threadIdx.x) * 6]);
        const float w0 = exp2f(M - Mt);                    // group 0 always has a piece (np >= 2 > 0)
        L *= w0; acc.x *= w0; acc.y *= w0; acc.z *= w0; acc.w *= w0;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const float* o = s_merge + (g * 64 + threadIdx.x) * 6;
          const float w = o[0] == -INFINITY ? 0.f : exp2f(o[0] - Mt);
          L += w * o[1];
          acc.x += w * o[2]; acc.y += w * o[3]; acc.z += w * o[4]; acc.w += w * o[5];
        }
      }
      __syncthreads();
      if (grp != 0) continue;
    }
    const float inv = 1.f / L;
    const float v[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
    const uint32_t h0 = pack_bf16x2(v[0], v[1]), h1 = pack_bf16x2(v[2], v[3]);
    const uint32_t l0 = pack_bf16x2(v[0] - __uint_as_float(h0 << 16), v[1] - __uint_as_float(h0 & 0xFFFF0000u));
    const uint32_t l1 = pack_bf16x2(v[2] - __uint_as_float(h1 << 16), v[3] - __uint_as_float(h1 & 0xFFFF0000u));
    const long long row = static_cast<long long>(b) * p.Nq + qp * 256 + r;
    __nv_bfloat16* dst = p.out + row * (2 * C) + h * 256 + d;
    *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(dst + C) = make_uint2(l0, l1);
  }
}

}  // namespace parq
