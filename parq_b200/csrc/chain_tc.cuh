// Chained tcgen05 GEMMs for the row-local part of a decoder iteration (sm_100a).
//
// Between the attention kernels an iteration is a chain of linear layers in which every output row depends only
// on the same row of the input (reference transformer_parq.py:365-386, 176-180, 211-281):
//     P:  pe0 -> ReLU -> [self-attention V^T, independent] -> pe2 (+x -> x+pe) -> self-attention Q|K projection
//     A:  self-attention out-projection + residual + LayerNorm1 (+pe) -> cross-attention Q projection
//     B:  cross-attention out-projection + residual + LayerNorm2 -> FFN linear1 + ReLU -> linear2 + residual +
//         LayerNorm3 -> first layer of the centre / rotation heads (+ GroupNorm tile sums)
// As separate launches each link is a one-wave GEMM of 128 CTAs whose launch latency, pipeline fill and epilogue
// are fully exposed (25-37 us for ~10 us of tensor work), plus a row-wise LayerNorm kernel after three of them.
// Here a CLUSTER OF 4 CTAs owns a block of 128 rows for the whole chain: CTA r computes columns [r N/4, (r+1) N/4)
// of every stage with the same TMA -> smem ring -> tcgen05.mma (TMEM accumulator) pipeline as gemm_tc.cuh, the
// stage's output goes to global memory (it stays in L2), and the next stage streams it back as its A operand as
// soon as all four CTAs have signalled "stage done" on that stage's cluster-scope mbarrier.  Weight tiles of the next stage
// are requested while the current stage's epilogue is still running, and a stage that does not read the previous one
// (ChainStage::dep) starts its MMAs right away: they run under the previous stage's epilogue.
//   LayerNorm inside a stage: the rows of z = acc + bias + residual are spread over the four CTAs (and two
// epilogue warps per row); every thread reduces its 96-128 values to (mean, M2), writes that pair into all four
// CTAs' shared memory (DSMEM), and after a cluster-scope mbarrier combines the 8 partials of its row in a fixed
// order (Chan's parallel variance: deterministic, no cancellation).  z waits in TMEM (written back in place).
//   Activations enter the tensor cores as an exact bf16 [hi|lo] split (see gemm_tc.cuh); every stage that feeds
// another GEMM emits that split directly.
#pragma once
#include <cuda.h>

#include "gemm2_tc.cuh"
#include "ptx.cuh"

namespace parq {

enum ChainEpilogue {
  CH_EP_LN = 0,     // y = LayerNorm(acc + bias + residual) * gamma + beta -> out_cm / out_f32, a_out = split(y), a_out_pe = split(y + pe)
  CH_EP_SPLIT = 1,  // v = [relu](acc + bias) -> a_out = split(v)
  CH_EP_LP = 2,     // v = acc + bias -> out_lp (bf16 / fp16)
  CH_EP_F32 = 3,    // v = acc + bias -> out_f32 [, GroupNorm tile sums] [, out_sum_split = split(v + add_split)]
  CH_EP_LP_T = 4,   // v = acc + bias -> out_lp TRANSPOSED: out_lp[col * ld_lp + row] (V^T of the self-attention, K-major for P.V)
};

struct ChainStage {
  int N, K;                  // output columns of the stage (all four CTAs), K per term (multiple of 64)
  int tile_n, tiles;         // per CTA: `tiles` accumulator tiles of `tile_n` columns; tiles * tile_n = N / 4
  int nterms, a_koff[3], b_koff[3], dual_a;
  int hi_only;               // dual-A ring only: skip the low-order activation term (outputs that are rounded to 16 bits anyway)
  int ep, relu, lp_fp16;
  int dep;                   // index of the earlier stage of this launch whose output is this stage's A operand, or -1: the stage
                             // starts as soon as the tensor pipe is free (its MMAs overlap the previous stage's epilogue)
  const float* bias;                   // (N) or null
  // "cm" = COLUMN-MAJOR fp32 scratch [N][M]: streams that are written and later read by the SAME epilogue thread (the
  // residual stream x / x1 / x2 and the positional feature) -- a thread owns a row, so with rows contiguous every access
  // of a warp is one 128-byte line and needs no staging; every N = C stage of every chain maps (row, column) to the
  // same thread, also across launches.
  const float* resid_cm;               // LN: residual
  const float* gamma;
  const float* beta;
  const float* pe_cm;                  // LN: optional positional feature
  float* out_cm;                       // LN: y; F32: value
  float* add_cm_out;                   // F32 with add_split: the addend itself (hi + lo) as fp32
  float* out_f32;                      // row-major (M, N) copy of y / value for consumers outside the chains, or null
  __nv_bfloat16* a_out;                // (M, 2N) [hi|lo], or null
  __nv_bfloat16* a_out_pe;             // LN only, or null
  void* out_lp;                        // CH_EP_LP: (M, ld_lp) 16-bit
  long long ld_lp;
  double2* gn_out;                     // CH_EP_F32: optional (sum, sum of squares) per tile, slot m_tile * gn_stride + n0 / GN_SLOT_COLS (ptx.cuh)
  int gn_stride;
  const __nv_bfloat16* add_split;      // CH_EP_F32: optional (M, 2N) [hi|lo] addend ...
  __nv_bfloat16* out_sum_split;        // ... and the split of (v + addend)
};

namespace chain {
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int CLUSTER = 4;
constexpr int MAX_STAGES = 4;
constexpr int THREADS = 384;           // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warp 3 idle, warps 4-11 epilogue
constexpr int EPI_WARPS = 8;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = 256 * BK * 2;  // 32 KB
constexpr int RING_BYTES = 4 * (A_BYTES + B_BYTES);
constexpr int VEC_COLS = 512;          // columns of one CTA in a stage (bias / gamma / beta staging)
constexpr int LN_COLS = 256;           // columns of one CTA in a LayerNorm stage (gamma / beta staging)
constexpr int STAGE_WORDS = 32 * 20;   // per epilogue warp: 32 rows x 16 words, row stride 20 (16-byte accesses, conflict free)
constexpr int SMEM_BYTES = 1024 + RING_BYTES + 512 /*barriers*/ + 2 * CLUSTER * BM * 8 /*row statistics*/ + (VEC_COLS + 2 * LN_COLS) * 4 + 256 /*GN sums*/ +
                           EPI_WARPS * STAGE_WORDS * 4;
}  // namespace chain

struct ChainParams {
  int M, nstages;
  long long* dbg;          // optional: clock64 stamps of CTA 0 (tools/chain_timeline.py), 64 slots per launch, or null
  int trace_slot;          // >= 0 while parq_trace is on: every CTA stamps the global timer at entry / dependency resolved / exit into
                           // the upper half of the trace buffer, [trace_slot][blockIdx.x (< 160)][4]
  ChainStage st[chain::MAX_STAGES];
};
struct ChainMaps {
  CUtensorMap a[chain::MAX_STAGES], b[chain::MAX_STAGES];
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
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
// generic-proxy writes (st.global of the epilogue) <-> async-proxy reads (TMA loads of the next stage), all state spaces
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- coalesced global I/O of the epilogue ------------------------------------------------------------------------
// An epilogue thread owns one ROW of the accumulator (TMEM lane = row); written naively, every 16-byte global access of a
// warp would touch 32 different rows.  Row-major tensors therefore move through a padded shared-memory tile per warp,
// 32 rows x 16 words with a row stride of 20 words: the owner of row r reads / writes 16-byte pieces at r*20 + 4j
// (conflict free per quarter warp), towards memory 4 lanes cover 64 contiguous bytes of a row (8 rows per instruction).
// g points at [first row of the warp][first word of the 16-word block]; ldw = row pitch in 32-bit words.
constexpr int CH_STRIDE = 20;
__device__ __forceinline__ void chain_flush16(const uint32_t* stage, uint32_t* g, long long ldw, int lane) {
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), w = (lane & 3) * 4;
    *reinterpret_cast<uint4*>(g + r * ldw + w) = *reinterpret_cast<const uint4*>(stage + r * CH_STRIDE + w);
  }
  __syncwarp();
}
// 32 fp32 values of this lane's row -> row-major global (row pitch ld elements)
__device__ __forceinline__ void chain_store_f32x32(uint32_t* stage, float* g, long long ld, const float* v, int lane) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<uint4*>(stage + lane * CH_STRIDE + 4 * i) =
          make_uint4(__float_as_uint(v[hf * 16 + 4 * i]), __float_as_uint(v[hf * 16 + 4 * i + 1]), __float_as_uint(v[hf * 16 + 4 * i + 2]),
                     __float_as_uint(v[hf * 16 + 4 * i + 3]));
    chain_flush16(stage, reinterpret_cast<uint32_t*>(g) + hf * 16, ld, lane);
  }
}
// 16 packed words of this lane's row (32 bf16 / fp16 values) -> row-major global (row pitch ld 16-bit elements)
__device__ __forceinline__ void chain_store_w16(uint32_t* stage, void* g, long long ld, const uint32_t* w, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stage + lane * CH_STRIDE + 4 * i) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  chain_flush16(stage, reinterpret_cast<uint32_t*>(g), ld / 2, lane);
}
// 32 values -> bf16 "hi" at g and bf16 residuals at g + lo_off (row pitch ld bf16 elements)
__device__ __forceinline__ void chain_store_split32(uint32_t* stage, __nv_bfloat16* g, long long ld, long long lo_off, const float* v, int lane) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    hi[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    lo[i] = pack_bf16x2(v[2 * i] - __uint_as_float(hi[i] << 16), v[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
  }
  chain_store_w16(stage, g, ld, hi, lane);
  chain_store_w16(stage, g + lo_off, ld, lo, lane);
}
// the 16-byte global loads of a 32-row x 16-word block (issued early, consumed by chain_unstage16)
__device__ __forceinline__ void chain_issue16(uint4 (&t)[4], const uint32_t* g, long long ldw, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), w = (lane & 3) * 4;
    t[it] = *reinterpret_cast<const uint4*>(g + r * ldw + w);
  }
}
// ... through the tile into the 16 words of this lane's row
__device__ __forceinline__ void chain_unstage16(uint32_t* stage, const uint4 (&t)[4], uint32_t* w, int lane) {
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2), c = (lane & 3) * 4;
    *reinterpret_cast<uint4*>(stage + r * CH_STRIDE + c) = t[it];
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 v = *reinterpret_cast<const uint4*>(stage + lane * CH_STRIDE + 4 * i);
    w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
  }
  __syncwarp();
}

#define CHAIN_STAMP(slot)                                                        \
  do {                                                                           \
    if (p.dbg != nullptr && blockIdx.x == 0) p.dbg[slot] = clock64();           \
  } while (0)

// grid = 4 * (M / 128) CTAs in clusters of 4 along x; M % 128 == 0.
__global__ void __launch_bounds__(chain::THREADS, 1)
chain_tc_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  using namespace chain;
  extern __shared__ uint8_t smem_raw_c[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_c) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RING_BYTES);   // [4]
  uint64_t* empty_bar = full_bar + 4;                                    // [4]
  uint64_t* tfull_bar = empty_bar + 4;                                   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                                  // [2]
  uint64_t* xbar = tempty_bar + 2;        // row statistics of all four CTAs have arrived
  uint64_t* dbar = xbar + 1;              // [MAX_STAGES] stage s: its outputs of all four CTAs are in global memory (one barrier per
                                          // stage, single phase: an independent stage may run ahead of the previous stage's epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dbar + MAX_STAGES);
  float2* s_part = reinterpret_cast<float2*>(smem + RING_BYTES + 512);   // [2 * CLUSTER][BM] (mean, M2) partials
  float* s_vec = reinterpret_cast<float*>(s_part + 2 * CLUSTER * BM);    // bias [VEC_COLS] | gamma [LN_COLS] | beta [LN_COLS]
  double* s_gn = reinterpret_cast<double*>(s_vec + VEC_COLS + 2 * LN_COLS);   // [EPI_WARPS][2]
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_gn) + 256);   // [EPI_WARPS][STAGE_WORDS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m0 = (blockIdx.x / CLUSTER) * BM;
  unsigned long long* tr = nullptr;
  const long long tr_c0 = clock64();
  if (threadIdx.x == 0 && p.trace_slot >= 0 && g_trace_buf != nullptr && blockIdx.x < 160) {
    const unsigned long long o = g_trace_cap / 2 + (static_cast<unsigned long long>(p.trace_slot) * 160 + blockIdx.x) * 4;
    if (o + 4 <= g_trace_cap) {
      tr = g_trace_buf + o;
      tr[0] = globaltimer_ns();
    }
  }
  // every stage of a chain uses the same ring geometry: dual-A (3 slots of [A_hi | A_lo | B]) or plain (4 slots of A + B)
  const bool dual = p.st[0].dual_a != 0;
  const int nst = dual ? 3 : 4;
  auto a_ptr = [&](int st, int which) { return dual ? smem + st * (2 * A_BYTES + B_BYTES) + which * A_BYTES : smem + st * A_BYTES; };
  auto b_ptr = [&](int st) { return dual ? smem + st * (2 * A_BYTES + B_BYTES) + 2 * A_BYTES : smem + 4 * A_BYTES + st * B_BYTES; };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.b[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    mbar_init(xbar, CLUSTER * EPI_WARPS);
    for (int i = 0; i < MAX_STAGES; ++i) mbar_init(&dbar[i], CLUSTER * EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();                      // the peers' barriers exist before anyone arrives on them

  if (warp == 0) {
    if (lane == 0) {                       // ---------------------------------------------------- TMA producer
      int slot = 0;
      uint32_t phase = 0;
      for (int s = 0; s < p.nstages; ++s) {
        const ChainStage& S = p.st[s];
        const int kb_per_term = S.K / BK;
        const int nterm_loops = dual ? 1 : S.nterms;
        const int num_kb = kb_per_term * nterm_loops;
        const uint32_t b_bytes = static_cast<uint32_t>(S.tile_n) * BK * 2;
        const bool lo_term = dual && S.hi_only == 0;
        const uint32_t stage_tx = (lo_term ? 2 * A_BYTES : A_BYTES) + b_bytes;
        const int n_base = static_cast<int>(rank) * (S.N / CLUSTER);
        auto a_off = [&](int t) { return t == 0 ? S.a_koff[0] : (t == 1 ? S.a_koff[1] : S.a_koff[2]); };
        auto b_off = [&](int t) { return t == 0 ? S.b_koff[0] : (t == 1 ? S.b_koff[1] : S.b_koff[2]); };
        // weights do not depend on the previous stage (or kernel): the B tiles of the first ring slots are requested
        // before the wait, the A tiles of the same slots after it
        const int total = num_kb * S.tiles;
        const int pre = total < nst ? total : nst;
        {
          int ps = slot;
          uint32_t pp = phase;
          for (int i = 0; i < pre; ++i) {
            const int j = i / num_kb, r = i % num_kb, t = r / kb_per_term, kb = r % kb_per_term;
            mbar_wait(&empty_bar[ps], pp ^ 1);
            mbar_expect_tx(&full_bar[ps], stage_tx);
            tma_load_2d(b_ptr(ps), &maps.b[s], &full_bar[ps], b_off(t) + kb * BK, n_base + j * S.tile_n);
            if (++ps == nst) { ps = 0; pp ^= 1; }
          }
        }
        if (s == 0) {
          pdl_wait();
          pdl_launch_dependents();
          if (tr != nullptr) tr[1] = globaltimer_ns();
        } else if (S.dep >= 0) {
          mbar_wait_cluster(&dbar[S.dep], 0);
          fence_proxy_async_all();
        }
        CHAIN_STAMP(s * 8 + 0);              // the A operand of this stage may be loaded
        int idx = 0;
        for (int j = 0; j < S.tiles; ++j) {
          const int n0 = n_base + j * S.tile_n;
          for (int t = 0; t < nterm_loops; ++t) {
            for (int kb = 0; kb < kb_per_term; ++kb, ++idx) {
              if (idx >= pre) {
                mbar_wait(&empty_bar[slot], phase ^ 1);
                mbar_expect_tx(&full_bar[slot], stage_tx);
                tma_load_2d(b_ptr(slot), &maps.b[s], &full_bar[slot], b_off(t) + kb * BK, n0);
              }
              tma_load_2d(a_ptr(slot, 0), &maps.a[s], &full_bar[slot], a_off(t) + kb * BK, m0);
              if (lo_term) tma_load_2d(a_ptr(slot, 1), &maps.a[s], &full_bar[slot], a_off(1) + kb * BK, m0);
              if (++slot == nst) { slot = 0; phase ^= 1; }
            }
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
    if (lane == 0) {                       // ---------------------------------------------------- MMA issuer
      int slot = 0;
      uint32_t phase = 0;
      int cnt = 0;
      for (int s = 0; s < p.nstages; ++s) {
        const ChainStage& S = p.st[s];
        const int num_kb = (S.K / BK) * (dual ? 1 : S.nterms);
        const uint32_t idesc = umma_idesc(BM, static_cast<uint32_t>(S.tile_n), 1);
        for (int j = 0; j < S.tiles; ++j, ++cnt) {
          const int acc = cnt & 1;
          mbar_wait(&tempty_bar[acc], ((cnt >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[slot], phase);
            if (kb == 0 && j == 0) CHAIN_STAMP(s * 8 + 1);   // first operands have landed
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(a_ptr(slot, 0)));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(b_ptr(slot)));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            if (dual && S.hi_only == 0) {
              const uint64_t adesc1 = umma_desc_sw128(smem_u32(a_ptr(slot, 1)));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_ss(d_tmem, adesc1 + 2 * k, bdesc + 2 * k, idesc, 1u);
            }
            umma_commit(&empty_bar[slot]);
            if (++slot == nst) { slot = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);
          if (j == S.tiles - 1) CHAIN_STAMP(s * 8 + 2);      // last MMA of the stage issued
        }
      }
    }
  } else if (warp >= 4) {                  // ---------------------------------------------------- epilogue warps
    pdl_wait();
    pdl_launch_dependents();
    const int e = warp - 4;
    const int q = e & 3;                   // TMEM lane quadrant == warp % 4
    const int h = e >> 2;                  // which half of the tile's columns
    const int et = threadIdx.x - 128;      // 0..255
    const int rin = q * 32 + lane;         // row inside the block
    const long long row = m0 + rin;
    const long long wrow0 = m0 + q * 32;   // first row of this warp
    uint32_t* stage = s_stage + e * STAGE_WORDS;
    int cnt = 0, nln = 0;
    uint32_t xaddr[CLUSTER], daddr[CLUSTER];
#pragma unroll
    for (int c = 0; c < CLUSTER; ++c) {
      xaddr[c] = mapa_u32(smem_u32(xbar), c);
      daddr[c] = mapa_u32(smem_u32(dbar), c);          // + 8 * stage
    }
    for (int s = 0; s < p.nstages; ++s) {
      const ChainStage& S = p.st[s];
      const int ncta = S.N / CLUSTER;                        // columns of this CTA
      const int n_base = static_cast<int>(rank) * ncta;
      const int half = S.tile_n / 2, nchunks = half / 32;
      // stage the per-column vectors of this CTA's columns
      asm volatile("bar.sync 1, 256;" ::: "memory");        // the previous stage no longer reads s_vec
      for (int i = et; i < ncta; i += 256) {
        s_vec[i] = S.bias != nullptr ? __ldg(S.bias + n_base + i) : 0.f;
        if (S.ep == CH_EP_LN) {
          s_vec[VEC_COLS + i] = __ldg(S.gamma + n_base + i);
          s_vec[VEC_COLS + LN_COLS + i] = __ldg(S.beta + n_base + i);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int j = 0; j < S.tiles; ++j, ++cnt) {
        const int acc = cnt & 1;
        const int cl0 = j * S.tile_n + h * half;             // first column of this warp inside the CTA's columns
        const int col0 = n_base + cl0;                       // ... and in the stage's output
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + h * half;
        mbar_wait(&tfull_bar[acc], (cnt >> 1) & 1);
        tc_fence_after();
        if (et == 0 && j == S.tiles - 1) CHAIN_STAMP(s * 8 + 3);   // accumulator of the stage's last tile complete
        uint32_t r[32];
        if (S.ep == CH_EP_LN) {
          // ---- pass 1: z = acc + bias + residual, running (mean, M2) over this thread's columns, z back into TMEM
          float mean = 0.f, M2 = 0.f, n = 0.f;
          for (int c = 0; c < nchunks; ++c) {
            if (et == 0 && c == 1) CHAIN_STAMP(32 + s * 8 + 6);
            tmem_ld32(taddr + c * 32, r);
            float z[32];
            {
              const float* rp = S.resid_cm + static_cast<long long>(col0 + c * 32) * p.M + row;      // 32 independent, fully coalesced loads
#pragma unroll
              for (int i = 0; i < 32; ++i) z[i] = rp[static_cast<long long>(i) * p.M];
            }
            tmem_wait_ld();
            float cs = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = reinterpret_cast<const float4*>(s_vec + cl0 + c * 32)[i];
              z[4 * i] += __uint_as_float(r[4 * i]) + b4.x;
              z[4 * i + 1] += __uint_as_float(r[4 * i + 1]) + b4.y;
              z[4 * i + 2] += __uint_as_float(r[4 * i + 2]) + b4.z;
              z[4 * i + 3] += __uint_as_float(r[4 * i + 3]) + b4.w;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) cs += z[i];
            const float cm = cs * (1.f / 32.f);
            float cM2 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float d = z[i] - cm;
              cM2 = fmaf(d, d, cM2);
              r[i] = __float_as_uint(z[i]);
            }
            tmem_st32(taddr + c * 32, r);
            if (et == 0 && c == 1) CHAIN_STAMP(32 + s * 8 + 7);
            const float nn = n + 32.f, delta = cm - mean;
            mean += delta * (32.f / nn);
            M2 += cM2 + delta * delta * (n * 32.f / nn);
            n = nn;
          }
          tmem_wait_st();
          // ---- exchange: this thread's partial into slot (rank, h) of every CTA of the cluster
          {
            const uint32_t local = smem_u32(s_part + (rank * 2 + h) * BM + rin);
#pragma unroll
            for (int c = 0; c < CLUSTER; ++c) st_cluster_f32x2(mapa_u32(local, c), mean, M2);
          }
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < CLUSTER; ++c) mbar_arrive_cluster(xaddr[c]);
          }
          if (et == 0) CHAIN_STAMP(s * 8 + 4);                     // LayerNorm pass 1 done, partials sent
          mbar_wait_cluster(xbar, nln & 1);
          if (et == 0) CHAIN_STAMP(s * 8 + 5);                     // all partials here
          ++nln;
          {
            const float np = static_cast<float>(half);        // values per partial
            float mt = 0.f, Mt = 0.f, nt = 0.f;
#pragma unroll
            for (int k = 0; k < 2 * CLUSTER; ++k) {
              const float2 pk = s_part[k * BM + rin];
              const float nn = nt + np, delta = pk.x - mt;
              mt += delta * (np / nn);
              Mt += pk.y + delta * delta * (nt * np / nn);
              nt = nn;
            }
            mean = mt;
            M2 = Mt / nt;                                      // biased variance
          }
          const float rstd = 1.f / sqrtf(M2 + 1e-5f);
          // ---- pass 2: normalise, write the fp32 row and the operand splits
          for (int c = 0; c < nchunks; ++c) {
            if (et == 0 && c == 0) CHAIN_STAMP(32 + s * 8 + 0);
            tmem_ld32(taddr + c * 32, r);
            tmem_wait_ld();
            if (et == 0 && c == 0) CHAIN_STAMP(32 + s * 8 + 1);
            float y[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 g4 = reinterpret_cast<const float4*>(s_vec + VEC_COLS + cl0 + c * 32)[i];
              const float4 e4 = reinterpret_cast<const float4*>(s_vec + VEC_COLS + LN_COLS + cl0 + c * 32)[i];
              y[4 * i] = (__uint_as_float(r[4 * i]) - mean) * rstd * g4.x + e4.x;
              y[4 * i + 1] = (__uint_as_float(r[4 * i + 1]) - mean) * rstd * g4.y + e4.y;
              y[4 * i + 2] = (__uint_as_float(r[4 * i + 2]) - mean) * rstd * g4.z + e4.z;
              y[4 * i + 3] = (__uint_as_float(r[4 * i + 3]) - mean) * rstd * g4.w + e4.w;
            }
            const long long o = wrow0 * S.N + col0 + c * 32;
            const long long ocm = static_cast<long long>(col0 + c * 32) * p.M + row;
            float pv[32];
            if (S.a_out_pe != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) pv[i] = S.pe_cm[ocm + static_cast<long long>(i) * p.M];     // in flight during the stores below
            }
            if (et == 0 && c == 0) CHAIN_STAMP(32 + s * 8 + 2);
            if (S.out_cm != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) S.out_cm[ocm + static_cast<long long>(i) * p.M] = y[i];
            }
            if (et == 0 && c == 0) CHAIN_STAMP(32 + s * 8 + 3);
            if (S.out_f32 != nullptr) chain_store_f32x32(stage, S.out_f32 + o, S.N, y, lane);
            if (et == 0 && c == 0) CHAIN_STAMP(32 + s * 8 + 4);
            if (S.a_out != nullptr) chain_store_split32(stage, S.a_out + 2 * o - (col0 + c * 32), 2 * S.N, S.N, y, lane);
            if (et == 0 && c == 0) CHAIN_STAMP(32 + s * 8 + 5);
            if (S.a_out_pe != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] += pv[i];
              chain_store_split32(stage, S.a_out_pe + 2 * o - (col0 + c * 32), 2 * S.N, S.N, y, lane);
            }
          }
        } else {
          float gsum = 0.f, gsq = 0.f;
          for (int c = 0; c < nchunks; ++c) {
            tmem_ld32(taddr + c * 32, r);
            uint4 th[4], tl[4];
            if (S.add_split != nullptr) {
              const uint32_t* ag = reinterpret_cast<const uint32_t*>(S.add_split + wrow0 * 2 * S.N + col0 + c * 32);
              chain_issue16(th, ag, S.N, lane);
              chain_issue16(tl, ag + S.N / 2, S.N, lane);
            }
            tmem_wait_ld();
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = reinterpret_cast<const float4*>(s_vec + cl0 + c * 32)[i];
              v[4 * i] = __uint_as_float(r[4 * i]) + b4.x;
              v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4.y;
              v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4.z;
              v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4.w;
            }
            if (S.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            const long long cc = col0 + c * 32;
            if (S.ep == CH_EP_SPLIT) {
              chain_store_split32(stage, S.a_out + wrow0 * 2 * S.N + cc, 2 * S.N, S.N, v, lane);
            } else if (S.ep == CH_EP_LP_T) {
              // lanes = 32 consecutive rows: every store instruction writes 64 contiguous bytes of one output row (= column here)
              uint16_t* o = reinterpret_cast<uint16_t*>(S.out_lp) + cc * S.ld_lp + row;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const uint16_t hv = S.lp_fp16 ? __half_as_ushort(__float2half_rn(v[i])) : __bfloat16_as_ushort(__float2bfloat16_rn(v[i]));
                o[static_cast<long long>(i) * S.ld_lp] = hv;
              }
            } else if (S.ep == CH_EP_LP) {
              uint32_t w[16];
              if (S.lp_fp16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
              }
              chain_store_w16(stage, reinterpret_cast<uint16_t*>(S.out_lp) + wrow0 * S.ld_lp + cc, S.ld_lp, w, lane);
            } else {
              if (S.gn_out != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { gsum += v[i]; gsq = fmaf(v[i], v[i], gsq); }
              }
              const long long ocm = cc * p.M + row;
              if (S.out_cm != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i) S.out_cm[ocm + static_cast<long long>(i) * p.M] = v[i];
              }
              if (S.out_f32 != nullptr) chain_store_f32x32(stage, S.out_f32 + wrow0 * S.N + cc, S.N, v, lane);
              if (S.add_split != nullptr) {
                // addend = hi + lo of a row-major split (its loads were issued before the accumulator wait)
                uint32_t hw[16], lw[16];
                chain_unstage16(stage, th, hw, lane);
                chain_unstage16(stage, tl, lw, lane);
                float a[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  a[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
                  a[2 * i + 1] = __uint_as_float(hw[i] & 0xFFFF0000u) + __uint_as_float(lw[i] & 0xFFFF0000u);
                }
                if (S.add_cm_out != nullptr) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) S.add_cm_out[ocm + static_cast<long long>(i) * p.M] = a[i];
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] += v[i];
                chain_store_split32(stage, S.out_sum_split + wrow0 * 2 * S.N + cc, 2 * S.N, S.N, a, lane);
              }
            }
          }
          if (S.ep == CH_EP_F32 && S.gn_out != nullptr) {
            // deterministic tile statistics for the following GroupNorm: lanes -> warp (shuffles, double) -> 8 warps (smem)
            double ds = gsum, dq = gsq;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              ds += __shfl_xor_sync(0xffffffffu, ds, o);
              dq += __shfl_xor_sync(0xffffffffu, dq, o);
            }
            if (lane == 0) { s_gn[2 * e] = ds; s_gn[2 * e + 1] = dq; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) {
              double a = 0.0, b = 0.0;
              for (int k = 0; k < EPI_WARPS; ++k) { a += s_gn[2 * k]; b += s_gn[2 * k + 1]; }
              // slots of GN_SLOT_COLS columns (ptx.cuh): the tile's sums in its first slot, zeros in the others it covers
              double2* slot = S.gn_out + static_cast<long long>(m0 / BM) * S.gn_stride + (n_base + j * S.tile_n) / GN_SLOT_COLS;
              slot[0] = make_double2(a, b);
              for (int i = 1; i < S.tile_n / GN_SLOT_COLS; ++i) slot[i] = make_double2(0.0, 0.0);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if (et == 0) CHAIN_STAMP(s * 8 + 6);                         // epilogue of the stage done (this warp)
      // the stage's outputs of this warp are written: publish them to the TMA engines of the whole cluster
      if (s + 1 < p.nstages) {
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < CLUSTER; ++c) mbar_arrive_cluster(daddr[c] + 8u * s);
        }
      }
    }
  } else {
    pdl_wait();
    pdl_launch_dependents();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while a peer may still signal its barriers or write its statistics
  if (tr != nullptr) {
    tr[2] = globaltimer_ns();
    tr[3] = static_cast<unsigned long long>(clock64() - tr_c0);      // SM cycles between the two stamps: the clock under load
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace parq
