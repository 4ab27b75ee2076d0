// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//
//   D[M,N] = sum_t  A_t[M,K] * B_t[N,K]^T          (bf16 operands, fp32 accumulate in TMEM)
//
// Both operands are K-major (K contiguous) and are fetched by TMA into a
// 4-stage 128B-swizzled shared-memory ring; a single thread issues
// tcgen05.mma (128 x 256 x 16) into a double-buffered 2 x 256-column TMEM
// accumulator so that the epilogue of tile i overlaps the main loop of tile i+1.
//
// "Terms": the residual stream of the decoder is fp32, so activations enter the
// tensor cores as an exact two-way bf16 split  x = hi + lo  stored side by side
// ([hi | lo] along K).  A GEMM with `nterms` > 1 walks several (A-column-offset,
// B-column-offset) K-segments and accumulates them into the same tile:
//   2 terms: (A_hi,W_hi) + (A_lo,W_hi)                 (weights exactly bf16)
//   3 terms: ... + (A_hi,W_lo)                         (fp32 weights split too)
//
// Used for: the hoisted cross-attention K / V^T projections of all image tokens
// (reference: nn.MultiheadAttention in-proj of `memory`, transformer_parq.py:377-380),
// and every per-iteration linear layer of the decoder (in/out projections, FFN,
// reference-point MLP, head hidden layers; transformer_parq.py:176-180,365-386,
// generic_mlp.py:94-110).
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace parq {

struct GemmEpilogue {
  const float* bias;   // nullptr: none
  int bias_per_row;    // 0: bias[n] (Linear), 1: bias[m] (transposed product, e.g. V^T)
  int relu;
  float* out_f32;      // optional fp32 output, row-major, leading dimension ld_f32
  long long ld_f32;
  void* out_lp;        // optional 16-bit output (bf16, or fp16 when lp_fp16)
  long long ld_lp;
  int lp_fp16;
  long long lp_lo_off; // > 0: also store the bf16 residual (v - hi) at column offset lp_lo_off
  double2* gn_out;     // optional: per-tile (sum, sum of squares) of the outputs, slot m_tile*gn_stride + column/GN_SLOT_COLS (ptx.cuh)
  int gn_stride;
  // Tile-contiguous K / V^T cache for the cross-attention (16-bit output only; requires Nk % 32 == 0):
  //   1: rows are tokens (K = tokens Wk^T):   out[((tile*H + h)*128 + key%128)*256 + ch]
  //   2: columns are tokens (V^T = Wv tokens^T): out[((tile*H + h)*256 + ch)*128 + key%128]
  // with token = b*Nk + key, tile = b*ntile + key/128, channel = h*256 + ch: every (key tile, head) block the attention
  // kernel streams is one contiguous 64 KB piece of HBM.
  int kv_tiled;
  int kv_Nk, kv_ntile, kv_H;
  long long kv_tok_offset;   // added to the token index of a row / column: a GEMM over a sub-range of the tokens (one view of one clip)
  // Channels-first companions of a token-major product (rows = tokens (bt, pixel), columns = channels):
  //   nchw_out[(bt*N + col)*HW + pixel] = value                    (fp32; AddRayPE's (B,T,C,H,W) encoding)
  //   value += nchw_add[(bt*N + col)*HW + pixel]                    (fp32 backbone features, before out_f32 / out_lp)
  const float* nchw_add;
  int nchw_add_bf16;   // the addend is bf16 (channels-first features written by parq_fpn_concat_ex), not fp32
  float* nchw_out;
  int nchw_HW;
  int add_tma;         // the addend is staged through shared memory by TMA (third tensor map; needs HW % 4 == 0)
};

// element offset of the 32x32 chunk whose first row / column are (row0, col0) in a tiled K / V^T cache, and its row pitch
__device__ __forceinline__ long long kv_tiled_base(const GemmEpilogue& ep, long long row0, int col0, int& ld) {
  const long long tok = (ep.kv_tiled == 1 ? row0 : col0) + ep.kv_tok_offset;
  const int chan = ep.kv_tiled == 1 ? col0 : static_cast<int>(row0);
  const long long b = tok / ep.kv_Nk;
  const int key = static_cast<int>(tok - b * ep.kv_Nk);
  const long long tile = b * ep.kv_ntile + (key >> 7);
  const int h = chan >> 8, ch = chan & 255, kin = key & 127;
  if (ep.kv_tiled == 1) {
    ld = 256;
    return ((tile * ep.kv_H + h) * 128 + kin) * 256 + ch;
  }
  ld = 128;
  return ((tile * ep.kv_H + h) * 256 + ch) * 128 + kin;
}

struct GemmParams {
  int M, N, K;         // K per term, multiple of 64
  int nterms;
  int a_koff[3];       // element offset of each term along A's K axis
  int b_koff[3];
  int const_operand;   // 1: A holds constants (weights), 2: B does -- its first tiles are fetched before the PDL wait
  int a_split_n;       // > 0: output columns >= a_split_n read A at an extra K offset of a_split_off (two independent
  int a_split_off;     //      products that share M, e.g. the centre / rotation head layers, as ONE launch)
  int bn;              // output-tile width of the single-CTA kernel: 0 / 256 (default) or 64 -- narrow tiles spread a GEMM with few
                       // rows (one clip: 2 M-tiles) over 4x the SMs, each with a quarter of the MMA work (latency, not throughput)
  int dual_a;          // nterms == 2 with the SAME B segment (A_hi W + A_lo W): a ring stage holds both A tiles and one B
                       // tile, so W is fetched from L2 once per k-block instead of twice (3 stages of 64 KB)
  GemmEpilogue ep;
};

namespace gemm {
constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int THREADS = 256;
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BN * 4 /*bias*/ + 64 /*GroupNorm tile sums*/ +
                           4 * 32 * 33 * 4 /*per-warp store staging*/;
}  // namespace gemm

// Coalesced write of a warp's 32x32 block of 32-bit words (fp32 values, or packed 16-bit pairs when
// words_per_row == 16) that was staged in shared memory as stage[row][word] with row stride 33:
// eight (four) consecutive lanes cover one output row, so every store instruction writes whole 128-byte
// (64-byte) row segments instead of 32 scattered 16-byte pieces.
template <int kWordsPerRow, typename T>
__device__ __forceinline__ void gemm_flush_stage(const uint32_t* stage, T* out, long long ld_elems, int rows_valid, int lane) {
  constexpr int LANES_PER_ROW = kWordsPerRow / 4;
  constexpr int ROWS_PER_IT = 32 / LANES_PER_ROW;
  constexpr int ELEMS_PER_WORD = 4 / sizeof(T);
#pragma unroll
  for (int it = 0; it < 32 / ROWS_PER_IT; ++it) {
    const int r = it * ROWS_PER_IT + lane / LANES_PER_ROW;
    const int w = (lane % LANES_PER_ROW) * 4;
    const uint32_t* sp = stage + r * 33 + w;
    const uint4 v = make_uint4(sp[0], sp[1], sp[2], sp[3]);
    if (r < rows_valid) *reinterpret_cast<uint4*>(out + r * ld_elems + w * ELEMS_PER_WORD) = v;
  }
}

// Epilogue for a warp's 32 rows x 32 accumulator columns (thread = row): bias / ReLU, GroupNorm tile sums,
// then fp32 and/or 16-bit outputs (bf16 or fp16; optionally the bf16 residual "lo" of the split layout).
// Full chunks go through the per-warp shared-memory stage for coalesced stores; a ragged last chunk
// (N not a multiple of 32) is written directly with per-element guards.
// kNchw selects a lean instantiation for the token-major products with channels-first companions (AddRayPE producer):
// only bias / ReLU, nchw_add / nchw_out and a plain bf16 output exist there, which leaves the registers to keep all
// 32 addend loads of a chunk in flight; the general instantiation carries everything else and no channels-first code.
template <bool kNchw>
__device__ __forceinline__ void gemm_store_chunk(const GemmEpilogue& ep, const uint32_t (&r)[32], const float* sbias, float row_bias,
                                                 long long row0, int lane, int col0, int M, int N, uint32_t* stage, float& gsum,
                                                 float& gsq, const float* sadd = nullptr, int rowin = 0) {
  const long long row = row0 + lane;
  const bool row_ok = row < M;
  const int rows_valid = M - row0 < 32 ? static_cast<int>(M - row0) : 32;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + row_bias;
  if (sbias != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b4 = reinterpret_cast<const float4*>(sbias)[i];
      v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  const bool full = (col0 + 32 <= N);
  if constexpr (kNchw) {
   if ((ep.nchw_add != nullptr || ep.nchw_out != nullptr) && row_ok) {
    // for a fixed column the 32 lanes (consecutive tokens = consecutive pixels) touch one contiguous 128-byte run
    const long long bt = row / ep.nchw_HW;
    const long long off = (bt * N + col0) * ep.nchw_HW + (row - bt * ep.nchw_HW);
    if (ep.nchw_out != nullptr) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (full || col0 + i < N) ep.nchw_out[off + static_cast<long long>(i) * ep.nchw_HW] = v[i];
    }
    if (sadd != nullptr) {
      // addend chunk staged by TMA as [32 channels][128 tile rows] (fp32 or bf16): conflict-free column reads
      if (ep.nchw_add_bf16) {
        const __nv_bfloat16* s16 = reinterpret_cast<const __nv_bfloat16*>(sadd);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __bfloat162float(s16[i * 128 + rowin]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += sadd[i * 128 + rowin];
      }
    } else if (ep.nchw_add != nullptr) {
      // tiles that straddle two images: straight from global memory (a loop of its own, so the 32 loads are
      // independent of the stores above)
      if (ep.nchw_add_bf16) {
        const __nv_bfloat16* __restrict__ add = reinterpret_cast<const __nv_bfloat16*>(ep.nchw_add) + off;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (full || col0 + i < N) v[i] += __bfloat162float(add[static_cast<long long>(i) * ep.nchw_HW]);
      } else {
        const float* __restrict__ add = ep.nchw_add + off;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (full || col0 + i < N) v[i] += __ldg(add + static_cast<long long>(i) * ep.nchw_HW);
      }
    }
   }
    if (ep.out_lp != nullptr) {      // plain bf16 rows (the channels-last tokens)
      uint16_t* obase = reinterpret_cast<uint16_t*>(ep.out_lp) + row0 * ep.ld_lp + col0;
      if (full) {
#pragma unroll
        for (int i = 0; i < 16; ++i) stage[lane * 33 + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        __syncwarp();
        gemm_flush_stage<16>(stage, obase, ep.ld_lp, rows_valid, lane);
        __syncwarp();
      } else if (row_ok) {
        for (int i = 0; i < 32; ++i)
          if (col0 + i < N) obase[lane * ep.ld_lp + i] = __bfloat16_as_ushort(__float2bfloat16_rn(v[i]));
      }
    }
    return;
  }
  if (ep.gn_out != nullptr && row_ok) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (full || col0 + i < N) { gsum += v[i]; gsq = fmaf(v[i], v[i], gsq); }
  }
  if (ep.out_f32 != nullptr) {
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; ++i) stage[lane * 33 + i] = __float_as_uint(v[i]);
      __syncwarp();
      gemm_flush_stage<32>(stage, ep.out_f32 + row0 * ep.ld_f32 + col0, ep.ld_f32, rows_valid, lane);
      __syncwarp();
    } else if (row_ok) {
      float* o = ep.out_f32 + row * ep.ld_f32 + col0;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) o[i] = v[i];
    }
  }
  if (ep.out_lp != nullptr && ep.kv_tiled != 0 && ep.kv_Nk % 32 == 0) {
    // tile-contiguous K / V^T cache: the chunk is a dense 32 x 32 block with its own origin and pitch
    // (Nk % 32 == 0 and N % 32 == 0: chunks are full and never straddle a key tile or a clip)
    int ld;
    uint16_t* dst = reinterpret_cast<uint16_t*>(ep.out_lp) + kv_tiled_base(ep, row0, col0, ld);
#pragma unroll
    for (int i = 0; i < 16; ++i) stage[lane * 33 + i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    __syncwarp();
    gemm_flush_stage<16>(stage, dst, ld, rows_valid, lane);
    __syncwarp();
  } else if (ep.out_lp != nullptr && ep.kv_tiled != 0) {
    // ragged clips (Nk % 32 != 0): a chunk may straddle a key tile or a clip, every element finds its own place
    if (row_ok) {
      uint16_t* base = reinterpret_cast<uint16_t*>(ep.out_lp);
      if (ep.kv_tiled == 1) {               // this thread's token row: 32 consecutive channels of one head
        int ld;
        uint16_t* dst = base + kv_tiled_base(ep, row, col0, ld);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (full || col0 + i < N) dst[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v[i]));
      } else {                               // this thread's channel row: 32 consecutive tokens
#pragma unroll                               // (full unroll: a dynamic index would push v[] into local memory)
        for (int i = 0; i < 32; ++i) {
          if (full || col0 + i < N) {
            int ld;
            base[kv_tiled_base(ep, row, col0 + i, ld)] = __bfloat16_as_ushort(__float2bfloat16_rn(v[i]));
          }
        }
      }
    }
  } else if (ep.out_lp != nullptr) {
    uint16_t* obase = reinterpret_cast<uint16_t*>(ep.out_lp) + row0 * ep.ld_lp + col0;
    uint32_t w[16];
    if (ep.lp_fp16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    }
    if (full) {
#pragma unroll
      for (int i = 0; i < 16; ++i) stage[lane * 33 + i] = w[i];
      __syncwarp();
      gemm_flush_stage<16>(stage, obase, ep.ld_lp, rows_valid, lane);
      __syncwarp();
    } else if (row_ok) {
      uint16_t* o = obase + lane * ep.ld_lp;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < N) o[i] = static_cast<uint16_t>((i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xFFFFu));
    }
    if (ep.lp_lo_off > 0) {   // residual of the bf16 split (only meaningful for bf16 outputs)
      uint32_t l[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float h0 = __uint_as_float(w[i] << 16), h1 = __uint_as_float(w[i] & 0xFFFF0000u);
        l[i] = pack_bf16x2(v[2 * i] - h0, v[2 * i + 1] - h1);
      }
      if (full) {
#pragma unroll
        for (int i = 0; i < 16; ++i) stage[lane * 33 + i] = l[i];
        __syncwarp();
        gemm_flush_stage<16>(stage, obase + ep.lp_lo_off, ep.ld_lp, rows_valid, lane);
        __syncwarp();
      } else if (row_ok) {
        uint16_t* ol = obase + lane * ep.ld_lp + ep.lp_lo_off;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < N) ol[i] = static_cast<uint16_t>((i & 1) ? (l[i >> 1] >> 16) : (l[i >> 1] & 0xFFFFu));
      }
    }
  }
}

template <bool kNchw>
__global__ void __launch_bounds__(gemm::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using namespace gemm;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  // ring layouts: plain = 4 stages, A tiles then B tiles; dual-A = 3 stages of [A_hi | A_lo | B] (same 192 KB)
  const bool dual = p.dual_a != 0;
  // Channels-first addend through TMA (kNchw only): the plain ring shrinks to 3 stages and the freed A stage
  // (16 KB) and B stage (32 KB) become three 16 KB addend buffers [32 channels][128 rows] fp32.
  const bool add_tma = kNchw && p.ep.add_tma != 0 && p.ep.nchw_add != nullptr && !dual;
  const int nst = (dual || add_tma) ? 3 : STAGES;
  uint64_t* add_full = full_bar + 16;      // [3], byte offset 128 of the barrier block
  uint64_t* add_empty = add_full + 3;      // [3]
  auto add_buf = [&](int i) { return reinterpret_cast<float*>(i == 0 ? smem + 3 * A_BYTES : smem + STAGES * A_BYTES + 3 * B_BYTES + (i - 1) * 16384); };
  auto tile_fast = [&](int m0) { return add_tma && (m0 % p.ep.nchw_HW) + BM <= p.ep.nchw_HW; };
  const int bn = p.bn > 0 ? p.bn : BN;                       // tile width (columns of B rows per stage: bn x 64)
  const uint32_t b_bytes = static_cast<uint32_t>(bn) * BK * 2;
  const uint32_t stage_tx = dual ? 2 * A_BYTES + b_bytes : A_BYTES + b_bytes;
  auto a_ptr = [&](int st, int which) { return dual ? smem + st * (2 * A_BYTES + B_BYTES) + which * A_BYTES : smem + st * A_BYTES; };
  auto b_ptr = [&](int st) { return dual ? smem + st * (2 * A_BYTES + B_BYTES) + 2 * A_BYTES : smem + STAGES * A_BYTES + st * B_BYTES; };
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);   // [2][BN]
  uint32_t* sstage = reinterpret_cast<uint32_t*>(sbias + 2 * BN) + 16;                   // [4 warps][32][33], after the GN sums

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&add_full[i], 1);
      mbar_init(&add_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_n = (p.N + bn - 1) / bn;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int num_tiles = tiles_m * tiles_n;
  // Tile order: the dimension with FEWER tiles varies fastest, so that the CTAs running at the same
  // time share the tile of the large (streamed) operand through L2 and it is read from HBM once.
  const bool m_fastest = tiles_m <= tiles_n;
  auto tile_origin = [&](int tile, int& m0, int& n0) {
    if (m_fastest) { m0 = (tile % tiles_m) * BM; n0 = (tile / tiles_m) * bn; }
    else           { m0 = (tile / tiles_n) * BM; n0 = (tile % tiles_n) * bn; }
  };
  const int kb_per_term = p.K / BK;
  const int nterm_loops = dual ? 1 : p.nterms;
  const int num_kb = kb_per_term * nterm_loops;     // ring stages consumed per tile

  if (warp == 0) {
    if (lane == 0) {                       // ---------------- TMA producer
      // static selects (not p.a_koff[t]): a dynamically indexed parameter array would force the whole
      // parameter block into local memory
      auto a_off = [&](int t) { return t == 0 ? p.a_koff[0] : (t == 1 ? p.a_koff[1] : p.a_koff[2]); };
      auto b_off = [&](int t) { return t == 0 ? p.b_koff[0] : (t == 1 ? p.b_koff[1] : p.b_koff[2]); };
      // The operand that holds weights does not depend on the previous kernel: its first ring stages are
      // requested before the programmatic-dependency wait, so the fetch overlaps the predecessor's tail.
      int pre = 0;
      if (p.const_operand != 0 && static_cast<int>(blockIdx.x) < num_tiles) {
        int m0, n0;
        tile_origin(blockIdx.x, m0, n0);
        pre = num_kb < nst ? num_kb : nst;
        for (int i = 0; i < pre; ++i) {
          const int t = i / kb_per_term, kb = i % kb_per_term;
          mbar_expect_tx(&full_bar[i], stage_tx);
          if (p.const_operand == 1) {
            tma_load_2d(a_ptr(i, 0), &tmA, &full_bar[i], a_off(t) + kb * BK, m0);
            if (dual) tma_load_2d(a_ptr(i, 1), &tmA, &full_bar[i], a_off(1) + kb * BK, m0);
          } else {
            tma_load_2d(b_ptr(i), &tmB, &full_bar[i], b_off(t) + kb * BK, n0);
          }
        }
      }
      pdl_wait();
      pdl_launch_dependents();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m0, n0;
        tile_origin(tile, m0, n0);
        const int asplit = (p.a_split_n > 0 && n0 >= p.a_split_n) ? p.a_split_off : 0;
        for (int t = 0; t < nterm_loops; ++t) {
          const int ak = a_off(t) + asplit, bk = b_off(t);
          for (int kb = 0; kb < kb_per_term; ++kb) {
            const bool prefetched = pre > 0;       // first k-blocks of this CTA's first tile
            if (!prefetched) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], stage_tx);
            } else {
              --pre;
            }
            if (!(prefetched && p.const_operand == 1)) {
              tma_load_2d(a_ptr(stage, 0), &tmA, &full_bar[stage], ak + kb * BK, m0);
              if (dual) tma_load_2d(a_ptr(stage, 1), &tmA, &full_bar[stage], a_off(1) + asplit + kb * BK, m0);
            }
            if (!(prefetched && p.const_operand == 2)) tma_load_2d(b_ptr(stage), &tmB, &full_bar[stage], bk + kb * BK, n0);
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
      const uint32_t idesc = umma_idesc(BM, static_cast<uint32_t>(bn), 1);
      int stage = 0;
      uint32_t phase = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
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
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          if (dual) {
            const uint64_t adesc1 = umma_desc_sw128(smem_u32(a_ptr(stage, 1)));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss(d_tmem, adesc1 + 2 * k, bdesc + 2 * k, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {                  // ---------------- epilogue warps
    pdl_wait();
    pdl_launch_dependents();
    const int q = warp - 4;                // TMEM lane quadrant == warp % 4
    const int et = threadIdx.x - 128;      // 0..127
    const bool col_bias = (p.ep.bias != nullptr) && !p.ep.bias_per_row;
    int lt = 0;
    int add_cc = 0;                        // addend chunks consumed (ring of 3 buffers)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      int m0, n0;
      tile_origin(tile, m0, n0);
      const int acc = lt & 1;
      const long long row0 = m0 + q * 32;
      const long long row = row0 + lane;
      uint32_t* stage = sstage + q * (32 * 33);
      // stage this tile's bias while the MMA of the tile is still running
      float row_bias = 0.f;
      if (col_bias) {
        float* sb = sbias + acc * BN;
        sb[et] = (et < bn && n0 + et < p.N) ? __ldg(p.ep.bias + n0 + et) : 0.f;
        sb[et + 128] = (et + 128 < bn && n0 + et + 128 < p.N) ? __ldg(p.ep.bias + n0 + et + 128) : 0.f;
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
      for (int c = 0; c < bn / 32; c += 2) {
        tmem_wait_ld();
        tmem_ld32(taddr + (c + 1) * 32, r1);                 // next chunk in flight while this one is stored
        int col0 = n0 + c * 32;
        if (col0 < p.N) {
          const float* sadd = nullptr;
          if (fast) { mbar_wait(&add_full[add_cc % 3], (add_cc / 3) & 1); sadd = add_buf(add_cc % 3); }
          gemm_store_chunk<kNchw>(p.ep, r0, col_bias ? sbias + acc * BN + c * 32 : nullptr, row_bias, row0, lane, col0, p.M, p.N, stage, gsum, gsq, sadd, et);
          if (fast) { mbar_arrive(&add_empty[add_cc % 3]); ++add_cc; }
        }
        tmem_wait_ld();
        if (c + 2 < bn / 32) tmem_ld32(taddr + (c + 2) * 32, r0);
        col0 += 32;
        if (col0 < p.N) {
          const float* sadd = nullptr;
          if (fast) { mbar_wait(&add_full[add_cc % 3], (add_cc / 3) & 1); sadd = add_buf(add_cc % 3); }
          gemm_store_chunk<kNchw>(p.ep, r1, col_bias ? sbias + acc * BN + (c + 1) * 32 : nullptr, row_bias, row0, lane, col0, p.M, p.N, stage, gsum, gsq, sadd, et);
          if (fast) { mbar_arrive(&add_empty[add_cc % 3]); ++add_cc; }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (p.ep.gn_out != nullptr) {
        // deterministic tile statistics for the following GroupNorm: lanes -> warp (shuffles, double) -> 4 warps (smem)
        double ds = gsum, dq = gsq;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ds += __shfl_xor_sync(0xffffffffu, ds, o);
          dq += __shfl_xor_sync(0xffffffffu, dq, o);
        }
        double* sg = reinterpret_cast<double*>(sbias + 2 * BN);      // [4][2]
        if (lane == 0) { sg[2 * q] = ds; sg[2 * q + 1] = dq; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          // slots of GN_SLOT_COLS columns (ptx.cuh): the tile's sums in its first slot, zeros in the others it covers
          double2* slot = p.ep.gn_out + static_cast<long long>(m0 / BM) * p.ep.gn_stride + n0 / GN_SLOT_COLS;
          slot[0] = make_double2((sg[0] + sg[2]) + (sg[4] + sg[6]), (sg[1] + sg[3]) + (sg[5] + sg[7]));
          for (int i = 1; i < bn / GN_SLOT_COLS; ++i) slot[i] = make_double2(0.0, 0.0);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }

  if (warp == 2 || warp == 3) {
    pdl_wait();
    pdl_launch_dependents();
    if (add_tma && warp == 2 && lane == 0) {
      // addend producer: one TMA box {128 pixels, 32 channels} per epilogue chunk, up to three chunks ahead of the
      // epilogue warps (also across tiles: the features do not depend on the MMA)
      int cc = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m0, n0;
        tile_origin(tile, m0, n0);
        if (!tile_fast(m0)) continue;
        const int bt = m0 / p.ep.nchw_HW, pix0 = m0 - bt * p.ep.nchw_HW;
        for (int c = 0; c < bn / 32 && n0 + c * 32 < p.N; ++c, ++cc) {
          const int buf = cc % 3;
          mbar_wait(&add_empty[buf], ((cc / 3) & 1) ^ 1);
          mbar_expect_tx(&add_full[buf], p.ep.nchw_add_bf16 ? 32 * 128 * 2 : 32 * 128 * 4);
          tma_load_2d(add_buf(buf), &tmC, &add_full[buf], pix0, bt * p.N + n0 + c * 32);
        }
      }
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
