// GEMM with the residual add + LayerNorm of the decoder layer fused into its epilogue
// (reference transformer_parq.py:375-376, 381-385:  x = LN(x + Linear(...)), post-norm, eps 1e-5, biased variance).
//
//   v[M, 1024]  = sum_t A_t[M,K] * W_t[1024,K]^T + bias + resid          (tensor cores, fp32 accumulate in TMEM)
//   x_out       = (v - mean_row) * rstd_row * gamma + beta               (fp32)
//   a_x, a_xpe  = exact bf16 [hi|lo] splits of x_out and of x_out + pe   (operands of the following GEMMs)
//
// A row of 1024 outputs spans four 128 x 256 tiles, so the four CTAs that own them form a CLUSTER and exchange their
// per-row partial statistics through distributed shared memory: every epilogue thread (= one row) stores its partial
// into the same slot of all four CTAs and arrives (release.cluster) on their mbarriers; after the acquire-wait all four
// partials are local and are summed in a fixed order (identical mean / rstd in the four CTAs, deterministic).
// Two exchanges (sum, then sum of squared deviations: the two-pass variance of the stand-alone kernel); between the
// passes v lives in TMEM (written back in place with tcgen05.st).  Saves, per LayerNorm, a kernel launch and the
// 16 MB write + 16 MB read of the intermediate y.
//
// STATUS: correct (parity-tested) but NOT the default.  Measured at config 2 (R = 4096 rows = one tile per CTA, so the
// epilogue is fully exposed and runs on 4 warps per SM): the row-wise kernels drop by 0.46 ms/step, the fused GEMMs cost
// 0.59 ms/step more (three TMEM passes, two cluster exchanges, per-thread residual reads) -> +0.13 ms/step.  It pays only
// when a CTA owns several tiles (M >= 2 x 37 x 128 rows) so that the epilogue overlaps the next main loop; selected
// with PARQ_FLAG_LN_FUSION.
#pragma once
#include "gemm2_tc.cuh"

namespace parq {

struct GemmLnParams {
  int M, K;                // N is 1024 (4 tiles of 256 = the cluster)
  int nterms;
  int a_koff[3], b_koff[3];
  int dual_a;
  const float* bias;       // (1024)
  const float* resid;      // (M, 1024) fp32
  const float* gamma;      // (1024)
  const float* beta;       // (1024)
  const float* pe;         // (M, 1024) fp32 or nullptr
  float* x_out;            // (M, 1024) fp32
  __nv_bfloat16* a_x;      // (M, 2048) [hi|lo] or nullptr
  __nv_bfloat16* a_xpe;    // (M, 2048) [hi|lo] of x_out + pe, or nullptr
};

namespace gemmln {
constexpr int N = 1024;
constexpr int CLUSTER = 4;
constexpr int SMEM_BYTES = gemm::STAGES * (gemm::A_BYTES + gemm::B_BYTES) + 1024 /*align*/ + 256 /*barriers*/ + 2 * 4 * 128 * 4 /*row statistics*/ +
                           4 * 32 * 33 * 4 /*per-warp store staging*/;
}  // namespace gemmln

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

__global__ void __launch_bounds__(gemm::THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmLnParams p) {
  using namespace gemm;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;          // [1]
  uint64_t* tempty_bar = tfull_bar + 1;              // [1]
  uint64_t* xbar = tempty_bar + 1;                   // [2] statistics exchanges (512 arrivals: 128 rows x 4 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 2);
  float* s_stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);      // [2][4][128]
  uint32_t* sstage = reinterpret_cast<uint32_t*>(s_stat + 2 * 4 * 128);                      // [4 warps][32][33]
  const bool dual = p.dual_a != 0;
  const int nst = dual ? 3 : STAGES;
  const uint32_t stage_tx = dual ? 2 * A_BYTES + B_BYTES : A_BYTES + B_BYTES;
  auto a_ptr = [&](int st, int which) { return dual ? smem + st * (2 * A_BYTES + B_BYTES) + which * A_BYTES : smem + st * A_BYTES; };
  auto b_ptr = [&](int st) { return dual ? smem + st * (2 * A_BYTES + B_BYTES) + 2 * A_BYTES : smem + STAGES * A_BYTES + st * B_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();             // the N tile of this CTA
  const int n0 = static_cast<int>(rank) * BN;
  const int cluster_id = blockIdx.x / gemmln::CLUSTER, nclusters = gridDim.x / gemmln::CLUSTER;
  const int tiles_m = (p.M + BM - 1) / BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 128);
    mbar_init(&xbar[0], 128 * gemmln::CLUSTER);
    mbar_init(&xbar[1], 128 * gemmln::CLUSTER);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peers' exchange barriers exist before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kb_per_term = p.K / BK;
  const int nterm_loops = dual ? 1 : p.nterms;
  const int num_kb = kb_per_term * nterm_loops;

  if (warp == 0) {
    if (lane == 0) {                       // ---------------- TMA producer
      auto a_off = [&](int t) { return t == 0 ? p.a_koff[0] : (t == 1 ? p.a_koff[1] : p.a_koff[2]); };
      auto b_off = [&](int t) { return t == 0 ? p.b_koff[0] : (t == 1 ? p.b_koff[1] : p.b_koff[2]); };
      // the weights (B) do not depend on the previous kernel: first ring stages before the PDL wait
      int pre = 0;
      if (cluster_id < tiles_m) {
        pre = num_kb < nst ? num_kb : nst;
        for (int i = 0; i < pre; ++i) {
          const int t = i / kb_per_term, kb = i % kb_per_term;
          mbar_expect_tx(&full_bar[i], stage_tx);
          tma_load_2d(b_ptr(i), &tmB, &full_bar[i], b_off(t) + kb * BK, n0);
        }
      }
      pdl_wait();
      pdl_launch_dependents();
      int stage = 0;
      uint32_t phase = 0;
      for (int mt = cluster_id; mt < tiles_m; mt += nclusters) {
        const int m0 = mt * BM;
        for (int t = 0; t < nterm_loops; ++t) {
          const int ak = a_off(t), bk = b_off(t);
          for (int kb = 0; kb < kb_per_term; ++kb) {
            const bool prefetched = pre > 0;
            if (!prefetched) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], stage_tx);
            } else {
              --pre;
            }
            tma_load_2d(a_ptr(stage, 0), &tmA, &full_bar[stage], ak + kb * BK, m0);
            if (dual) tma_load_2d(a_ptr(stage, 1), &tmA, &full_bar[stage], a_off(1) + kb * BK, m0);
            if (!prefetched) tma_load_2d(b_ptr(stage), &tmB, &full_bar[stage], bk + kb * BK, n0);
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
    if (lane == 0) {                       // ---------------- MMA issuer
      constexpr uint32_t idesc = umma_idesc(BM, BN, 1);
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      for (int mt = cluster_id; mt < tiles_m; mt += nclusters, ++lt) {
        mbar_wait(tempty_bar, (lt & 1) ^ 1);            // the epilogue has drained the accumulator of the previous tile
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(a_ptr(stage, 0)));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(b_ptr(stage)));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          if (dual) {
            const uint64_t adesc1 = umma_desc_sw128(smem_u32(a_ptr(stage, 1)));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss(tmem_base, adesc1 + 2 * k, bdesc + 2 * k, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar);
      }
    }
  } else if (warp >= 4) {                  // ---------------- epilogue: residual + LayerNorm over the cluster
    pdl_wait();
    pdl_launch_dependents();
    const int q = warp - 4;
    const int r = q * 32 + lane;           // row of the tile == TMEM lane
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t* stage = sstage + q * (32 * 33);
    // addresses of this row's statistics slot and of the exchange barriers in the four CTAs
    uint32_t slot_addr[2][gemmln::CLUSTER], bar_addr[2][gemmln::CLUSTER];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int c = 0; c < gemmln::CLUSTER; ++c) {
        slot_addr[e][c] = mapa_u32(smem_u32(s_stat + (e * 4 + static_cast<int>(rank)) * 128 + r), c);
        bar_addr[e][c] = mapa_u32(smem_u32(&xbar[e]), c);
      }
    int lt = 0;
    for (int mt = cluster_id; mt < tiles_m; mt += nclusters, ++lt) {
      const long long row = static_cast<long long>(mt) * BM + r;
      const bool row_ok = row < p.M;
      const long long rbase = (row_ok ? row : 0) * gemmln::N + n0;
      mbar_wait(tfull_bar, lt & 1);
      tc_fence_after();
      // ---- pass 1: v = acc + bias + residual, written back to TMEM; row partial sum
      float s1 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t a[32];
        tmem_ld32(taddr + c * 32, a);
        float4 rs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rs[i] = __ldg(reinterpret_cast<const float4*>(p.resid + rbase + c * 32) + i);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32) + i);
          const float v0 = __uint_as_float(a[4 * i]) + b4.x + rs[i].x, v1 = __uint_as_float(a[4 * i + 1]) + b4.y + rs[i].y;
          const float v2 = __uint_as_float(a[4 * i + 2]) + b4.z + rs[i].z, v3 = __uint_as_float(a[4 * i + 3]) + b4.w + rs[i].w;
          a[4 * i] = __float_as_uint(v0); a[4 * i + 1] = __float_as_uint(v1); a[4 * i + 2] = __float_as_uint(v2); a[4 * i + 3] = __float_as_uint(v3);
          s1 += (v0 + v1) + (v2 + v3);
        }
        tmem_st32(taddr + c * 32, a);
      }
      tmem_wait_st();
#pragma unroll
      for (int c = 0; c < gemmln::CLUSTER; ++c) st_cluster_f32(slot_addr[0][c], s1);
#pragma unroll
      for (int c = 0; c < gemmln::CLUSTER; ++c) mbar_arrive_cluster(bar_addr[0][c]);
      mbar_wait_cluster(&xbar[0], lt & 1);
      const float mean = ((s_stat[0 * 128 + r] + s_stat[1 * 128 + r]) + (s_stat[2 * 128 + r] + s_stat[3 * 128 + r])) * (1.f / gemmln::N);
      // ---- pass 1b: sum of squared deviations
      float s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t a[32];
        tmem_ld32(taddr + c * 32, a);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = __uint_as_float(a[i]) - mean; s2 = fmaf(d, d, s2); }
      }
#pragma unroll
      for (int c = 0; c < gemmln::CLUSTER; ++c) st_cluster_f32(slot_addr[1][c], s2);
#pragma unroll
      for (int c = 0; c < gemmln::CLUSTER; ++c) mbar_arrive_cluster(bar_addr[1][c]);
      mbar_wait_cluster(&xbar[1], lt & 1);
      const float var = ((s_stat[(4 + 0) * 128 + r] + s_stat[(4 + 1) * 128 + r]) + (s_stat[(4 + 2) * 128 + r] + s_stat[(4 + 3) * 128 + r])) * (1.f / gemmln::N);
      const float rstd = 1.f / sqrtf(var + 1e-5f);
      // ---- pass 2: normalise, affine, outputs (coalesced through the per-warp stage)
      const long long row0 = static_cast<long long>(mt) * BM + q * 32;
      const int rows_valid = p.M - row0 < 32 ? static_cast<int>(p.M - row0) : 32;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t a[32];
        tmem_ld32(taddr + c * 32, a);
        tmem_wait_ld();
        const int col0 = n0 + c * 32;
        float o[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0) + i), e4 = __ldg(reinterpret_cast<const float4*>(p.beta + col0) + i);
          o[4 * i] = (__uint_as_float(a[4 * i]) - mean) * rstd * g4.x + e4.x;
          o[4 * i + 1] = (__uint_as_float(a[4 * i + 1]) - mean) * rstd * g4.y + e4.y;
          o[4 * i + 2] = (__uint_as_float(a[4 * i + 2]) - mean) * rstd * g4.z + e4.z;
          o[4 * i + 3] = (__uint_as_float(a[4 * i + 3]) - mean) * rstd * g4.w + e4.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[lane * 33 + i] = __float_as_uint(o[i]);
        __syncwarp();
        gemm_flush_stage<32>(stage, p.x_out + row0 * gemmln::N + col0, gemmln::N, rows_valid, lane);
        __syncwarp();
        if (p.a_x != nullptr) {
          uint32_t hi[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { hi[i] = pack_bf16x2(o[2 * i], o[2 * i + 1]); stage[lane * 33 + i] = hi[i]; }
          __syncwarp();
          gemm_flush_stage<16>(stage, reinterpret_cast<uint16_t*>(p.a_x) + row0 * (2 * gemmln::N) + col0, 2 * gemmln::N, rows_valid, lane);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            stage[lane * 33 + i] = pack_bf16x2(o[2 * i] - __uint_as_float(hi[i] << 16), o[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
          __syncwarp();
          gemm_flush_stage<16>(stage, reinterpret_cast<uint16_t*>(p.a_x) + row0 * (2 * gemmln::N) + gemmln::N + col0, 2 * gemmln::N, rows_valid, lane);
          __syncwarp();
        }
        if (p.a_xpe != nullptr) {
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 p4 = __ldg(reinterpret_cast<const float4*>(p.pe + rbase + c * 32) + i);
              o[4 * i] += p4.x; o[4 * i + 1] += p4.y; o[4 * i + 2] += p4.z; o[4 * i + 3] += p4.w;
            }
          }
          uint32_t hi[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { hi[i] = pack_bf16x2(o[2 * i], o[2 * i + 1]); stage[lane * 33 + i] = hi[i]; }
          __syncwarp();
          gemm_flush_stage<16>(stage, reinterpret_cast<uint16_t*>(p.a_xpe) + row0 * (2 * gemmln::N) + col0, 2 * gemmln::N, rows_valid, lane);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            stage[lane * 33 + i] = pack_bf16x2(o[2 * i] - __uint_as_float(hi[i] << 16), o[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
          __syncwarp();
          gemm_flush_stage<16>(stage, reinterpret_cast<uint16_t*>(p.a_xpe) + row0 * (2 * gemmln::N) + gemmln::N + col0, 2 * gemmln::N, rows_valid, lane);
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
    }
  }

  if (warp == 2 || warp == 3) {
    pdl_wait();
    pdl_launch_dependents();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while a peer may still store into its statistics slots
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace parq
