// Flash-style attention for head_dim 256 on tcgen05/TMEM (sm_100a), split over keys -- the single-CTA kernel.
// (attn2_tc.cuh is the same algorithm on a CTA pair with cta_group::2, attn3_tc.cuh adds the stream-K schedule;
//  this one serves query counts that are not a multiple of 256 and defines the shared constants.)
//
// Replaces nn.MultiheadAttention's softmax(QK^T)V of the PARQ decoder layer
// (reference transformer_parq.py:372-382; torch multi_head_attention_forward),
// without ever materialising the (B*heads, Nq, Nk) score tensor:
//   * cross-attention of the 256 queries of a clip over all T*H*W image tokens
//     (bf16 operands; K and V^T are projected once per clip by gemm_tc.cuh),
//   * self-attention among the queries (same kernel, fp16 operands, Nk = Nq).
//
// One CTA = (clip b, head h, 128-query tile, key split s).  Roles:
//   warp 0   TMA producer   Q tile once, then K half-tiles / V^T chunks through a
//                           5 x 32 KB ring (order K0, [K(j+1), V(j)] ...)
//   warp 1   MMA issuer     S_j = Q K_j^T  (4x4 tcgen05.mma 128x128x16, SS) into one of two
//                           S buffers;  O += P_j V_j (2x4 tcgen05.mma 128x256x16, A = P from TMEM)
//   warp 2   TMEM allocator (512 columns: S0 | S1 | O)
//   warps 4-7 softmax       thread = query row = TMEM lane: online softmax in the exp2 domain with
//                           lazy rescaling of O (only when the running max grows by > 2^8),
//                           P_j written back over S_j as packed 16-bit pairs.
// S(j+1) is issued before waiting for P(j), so the tensor pipe runs QK^T of the next
// tile while the softmax warps work on the current one.
//
// Layouts (all K-major for the MMA, 128B-swizzled by TMA):
//   Q   [B*Nq , H*256]   rows = queries, pre-scaled by 1/sqrt(256)
//   K   [B*Nk , H*256]   rows = keys
//   V^T [H*256, ldv  ]   rows = channels, columns = keys of all clips (b*Nk + key)
// or, for the decoder's cross-attention (kv_tiled), the tile-contiguous caches written by the projection GEMM:
//   K   [clip*tile][head][128 keys][256 ch],  V^T [clip*tile][head][256 ch][128 keys]
// Output: un-normalised partial O (fp32), running max m (log2 units) and sum l per
// (clip, head, split, query); attn_combine_kernel merges the splits.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace parq {

struct AttnParams {
  int B, H, Nq, Nk;
  int nsplit;            // number of key splits (grid.x); every split is non-empty
  int tiles_per_split;   // key tiles (128 keys) per split
  float* o_part;         // [((b*H+h)*nsplit + s)*Nq + q][256]
  float2* ml_part;       // [((b*H+h)*nsplit + s)*Nq + q] = (m, l)
  __nv_bfloat16* out_direct;   // nsplit == 1 only: normalised output (B*Nq, 2*H*256) [hi|lo], no combine pass
  int kv_const;                // K / V^T were written >= 2 launches ago: their first tiles are fetched before the PDL wait
  int kv_tiled;                // K / V^T are tile-contiguous caches (gemm_tc.cuh GemmEpilogue::kv_tiled); else plain matrices
  int ntile;                   // key tiles per clip in the tiled cache
};

namespace attn {
constexpr int DH = 256;
constexpr int BQ = 128;
constexpr int BKEY = 128;
constexpr int NS = 5;                      // ring stages
constexpr int STAGE_BYTES = 32 * 1024;
constexpr int Q_BYTES = BQ * DH * 2;       // 64 KB
constexpr int THREADS = 256;
constexpr int SMEM_BYTES = Q_BYTES + NS * STAGE_BYTES + 1024 + 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units
}  // namespace attn

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- pieces shared by the three flash-attention kernels (attn_tc / attn2_tc / attn3_tc); a softmax warp owns 32 query rows,
// ---- thread = row (TMEM lane), s_tmem / o_tmem already carry the warp's lane offset ----------------------------------------
// One key tile of the online softmax: S (128 fp32 columns of this row) -> masked row maximum -> lazy rescale of O -> P = 2^(S - m)
// packed to 16 bits over S in TMEM -> arrive_p() tells the MMA issuer that P is there.
// mbarrier waits are by phase parity, so every completion of pv_done has to be observed exactly once and in order (skipping one
// lets a later wait alias an older phase): a tile consumes the completion of the previous P.V either before it touches O (the
// rescale) or at its end; `first` = first tile of a (segment of an) item, whose predecessor was consumed by the epilogue before.
template <bool kFp16, typename ArriveP>
__device__ __forceinline__ void attn_softmax_tile(uint32_t s_tmem, uint32_t o_tmem, int nvalid, bool first, uint64_t* pv_done, uint32_t pv_parity,
                                                  float& m_run, float& l_run, ArriveP arrive_p) {
  using namespace attn;
  uint32_t su[128];                                // raw fp32 bits of this row of S
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_ld32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&su[c * 32]));
  tmem_wait_ld();
  if (nvalid < BKEY) {                             // keys of this tile that do not belong to the clip
#pragma unroll
    for (int i = 0; i < 128; ++i)
      if (i >= nvalid) su[i] = 0xff800000u;        // -inf
  }
  float tmax = __uint_as_float(su[0]);
#pragma unroll
  for (int i = 1; i < 128; ++i) tmax = fmaxf(tmax, __uint_as_float(su[i]));
  tmax *= LOG2E;
  bool pv_seen = first;
  if (first) {
    m_run = tmax;
  } else if (__any_sync(0xffffffffu, tmax > m_run + RESCALE_THRESHOLD)) {
    // O must be quiescent: wait for the previous P.V to retire, then rescale this warp's 32 rows.
    mbar_wait(pv_done, pv_parity);
    pv_seen = true;
    tc_fence_after();
    const float m_new = fmaxf(m_run, tmax);
    const float alpha = fast_exp2(m_run - m_new);
#pragma unroll 1
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t o[32];
      tmem_ld32(o_tmem + c * 32, o);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st32(o_tmem + c * 32, o);
    }
    l_run *= alpha;
    m_run = m_new;
  }
  float lsum = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t pk[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float p0 = fast_exp2(fmaf(__uint_as_float(su[half * 64 + 2 * i]), LOG2E, -m_run));
      const float p1 = fast_exp2(fmaf(__uint_as_float(su[half * 64 + 2 * i + 1]), LOG2E, -m_run));
      lsum += p0 + p1;
      pk[i] = kFp16 ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
    }
    tmem_st32(s_tmem + half * 32, pk);
  }
  l_run += lsum;
  tmem_wait_st();
  tc_fence_before();
  arrive_p();
  if (!pv_seen) mbar_wait(pv_done, pv_parity);
}

// O of this row, normalised, as the [hi|lo] operand of the out-projection: dst = &out[row][head * 256], C = channels of the layer
__device__ __forceinline__ void attn_store_normalised(uint32_t o_tmem, float inv, __nv_bfloat16* dst, int C) {
  using namespace attn;
#pragma unroll 1
  for (int c = 0; c < DH / 32; ++c) {
    uint32_t o[32];
    tmem_ld32(o_tmem + c * 32, o);
    tmem_wait_ld();
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v0 = __uint_as_float(o[2 * i]) * inv, v1 = __uint_as_float(o[2 * i + 1]) * inv;
      hi[i] = pack_bf16x2(v0, v1);
      lo[i] = pack_bf16x2(v0 - __uint_as_float(hi[i] << 16), v1 - __uint_as_float(hi[i] & 0xFFFF0000u));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      reinterpret_cast<uint4*>(dst + c * 32)[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
      reinterpret_cast<uint4*>(dst + C + c * 32)[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
    }
  }
}

// un-normalised O of this row for a merge kernel: orow = &o_part[partial row][0]
__device__ __forceinline__ void attn_store_partial(uint32_t o_tmem, float* orow) {
  using namespace attn;
#pragma unroll 1
  for (int c = 0; c < DH / 32; ++c) {
    uint32_t o[32];
    tmem_ld32(o_tmem + c * 32, o);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      reinterpret_cast<float4*>(orow + c * 32)[i] =
          make_float4(__uint_as_float(o[4 * i]), __uint_as_float(o[4 * i + 1]), __uint_as_float(o[4 * i + 2]), __uint_as_float(o[4 * i + 3]));
  }
}

template <bool kFp16>
__global__ void __launch_bounds__(attn::THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
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
  const int split = blockIdx.x;
  const int qt = blockIdx.y;
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
      mbar_init(&p_full[i], 128);
    }
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0 && n > 0) {             // ---------------- TMA producer
      const int ch0 = h * DH;
      int stage = 0;
      uint32_t phase = 0;
      // plain matrices: K rows = keys of all clips, V^T columns = keys of all clips; tiled caches: one contiguous
      // [128 keys][256 ch] (K) / [256 ch][128 keys] (V^T) block per (key tile, head)
      auto load_k = [&](int tile) {
        const int row = p.kv_tiled ? ((b * p.ntile + tile) * p.H + h) * BKEY : b * p.Nk + tile * BKEY;
        const int col = p.kv_tiled ? 0 : ch0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_expect_tx(&kv_full[stage], STAGE_BYTES);
          uint8_t* dst = ring + stage * STAGE_BYTES;
          tma_load_2d(dst, &tmK, &kv_full[stage], col + (half * 2) * 64, row);
          tma_load_2d(dst + 16384, &tmK, &kv_full[stage], col + (half * 2 + 1) * 64, row);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      };
      auto load_v = [&](int tile) {
        const int col = p.kv_tiled ? 0 : b * p.Nk + tile * BKEY;
        const int row = p.kv_tiled ? ((b * p.ntile + tile) * p.H + h) * DH : ch0;
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          mbar_expect_tx(&kv_full[stage], STAGE_BYTES);
          tma_load_2d(ring + stage * STAGE_BYTES, &tmV, &kv_full[stage], col + kc * 64, row);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
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
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tma_load_2d(sQ + c * (BQ * 128), &tmQ, q_full, ch0 + c * 64, b * p.Nq + qt * BQ);
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
    if (lane == 0 && n > 0) {             // ---------------- MMA issuer
      constexpr uint32_t fmt = kFp16 ? 0u : 1u;
      constexpr uint32_t idesc_s = umma_idesc(BQ, BKEY, fmt);
      constexpr uint32_t idesc_pv = umma_idesc(BQ, DH, fmt);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_s = [&](int buf) {
        const uint32_t d_tmem = tmem_base + buf * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(ring + stage * STAGE_BYTES);
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const uint64_t qd = umma_desc_sw128(q_addr + (half * 2 + c2) * (BQ * 128));
            const uint64_t kd = umma_desc_sw128(k_addr + c2 * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(d_tmem, qd + 2 * k, kd + 2 * k, idesc_s, (half | c2 | k) != 0);
          }
          umma_commit(&kv_empty[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit(&s_full[buf]);
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
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after();
          const uint64_t vd = umma_desc_sw128(smem_u32(ring + stage * STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts(tmem_O, p_tmem + kc * 32 + k * 8, vd + 2 * k, idesc_pv, (j | kc | k) != 0);
          umma_commit(&kv_empty[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit(pv_done);
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
                               [&] { mbar_arrive(&p_full[buf]); });
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
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Merge the key splits: out = sum_s 2^(m_s-M) O_s / sum_s 2^(m_s-M) l_s, heads concatenated along
// channels (the layout nn.MultiheadAttention feeds to out_proj), emitted as the exact bf16 split
// [hi | lo] that the out-projection GEMM consumes.  One block per query row; 64 threads per head, a
// thread owns 4 consecutive channels (float4 loads, independent over the splits).
__global__ void __launch_bounds__(256)
attn_combine_kernel(const float* __restrict__ o_part, const float2* __restrict__ ml_part, __nv_bfloat16* __restrict__ out,
                    int H, int Nq, int nsplit) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;            // b*Nq + q
  const int b = row / Nq, q = row % Nq;
  const int C = H * 256;
  for (int item = threadIdx.x; item < H * 64; item += blockDim.x) {
    const int h = item >> 6, d = (item & 63) * 4;
    const long long base = (static_cast<long long>(b * H + h) * nsplit) * Nq + q;
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldg(&ml_part[base + static_cast<long long>(s) * Nq].x));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float L = 0.f;
#pragma unroll 4
    for (int s = 0; s < nsplit; ++s) {
      const long long idx = base + static_cast<long long>(s) * Nq;
      const float2 ml = __ldg(&ml_part[idx]);
      const float4 o = __ldg(reinterpret_cast<const float4*>(o_part + idx * 256 + d));
      const float w = exp2f(ml.x - M);
      L += w * ml.y;
      acc.x += w * o.x; acc.y += w * o.y; acc.z += w * o.z; acc.w += w * o.w;
    }
    const float inv = 1.f / L;
    const float v[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
    const uint32_t h0 = pack_bf16x2(v[0], v[1]), h1 = pack_bf16x2(v[2], v[3]);
    const uint32_t l0 = pack_bf16x2(v[0] - __uint_as_float(h0 << 16), v[1] - __uint_as_float(h0 & 0xFFFF0000u));
    const uint32_t l1 = pack_bf16x2(v[2] - __uint_as_float(h1 << 16), v[3] - __uint_as_float(h1 & 0xFFFF0000u));
    __nv_bfloat16* dst = out + static_cast<long long>(row) * (2 * C) + h * 256 + d;
    *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(dst + C) = make_uint2(l0, l1);
  }
}

}  // namespace parq
