// Row-wise (per query) kernels between the tensor-core GEMMs of one decoder iteration.
// All statistics, the residual stream and the box epilogue stay in fp32; GEMM inputs are
// emitted as an exact two-way bf16 split [hi | lo] (see gemm_tc.cuh).
#pragma once
#include "ptx.cuh"

namespace parq {

__device__ __forceinline__ void store_split(__nv_bfloat16* hi_ptr, long long lo_off, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi_ptr[0] = h;
  hi_ptr[lo_off] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// 8 consecutive values -> 8 bf16 "hi" (16 bytes) at dst and 8 bf16 residuals at dst + lo_off
__device__ __forceinline__ void store_split8(__nv_bfloat16* dst, long long lo_off, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hi[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    lo[i] = pack_bf16x2(v[2 * i] - __uint_as_float(hi[i] << 16), v[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst + lo_off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// pos2posemb3d (reference transformer_parq.py:45-64): 128 sin/cos features per axis, axes
// concatenated in the order (y, x, z).  out: (R, 2*384) [hi|lo].
// ---------------------------------------------------------------------------------------------
// feature j (0..383) of a reference point (rx, ry, rz): axes in the order (y, x, z), sin on even / cos on odd indices
// (not inlined: sinf / cosf carry a large slow path for big arguments; one copy instead of one per call site keeps the heads
// kernel, which calls this 12 times per row, inside the instruction cache)
__device__ __noinline__ float posemb_value(float rx, float ry, float rz, int j, const float* __restrict__ dim_t) {
  const int seg = j >> 7, i = j & 127;
  const float r = (seg == 0) ? ry : (seg == 1 ? rx : rz);
  const float a = __fdiv_rn(__fmul_rn(r, 6.283185307179586f), dim_t[i]);
  return (i & 1) ? cosf(a) : sinf(a);
}

__global__ void posemb_kernel(const float* __restrict__ ref, const float* __restrict__ dim_t, __nv_bfloat16* __restrict__ out,
                              int R) {
  pdl_wait();
  pdl_launch_dependents();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * 384) return;
  const int row = idx / 384, j = idx % 384;
  const float v = posemb_value(ref[row * 3 + 0], ref[row * 3 + 1], ref[row * 3 + 2], j, dim_t);
  store_split(out + static_cast<long long>(row) * 768 + j, 384, v);
}

// ---------------------------------------------------------------------------------------------
// a_sum = split(x + pe), x given as its [hi|lo] split (the sampled features), pe fp32: the query / key input of the
// self-attention (reference transformer_parq.py:372) on the un-chained launch path.  One thread per 8 channels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_sum_kernel(const __nv_bfloat16* __restrict__ a_x, const float* __restrict__ pe, __nv_bfloat16* __restrict__ a_sum, int R, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;       // 8-channel group
  if (i >= static_cast<long long>(R) * C / 8) return;
  const long long row = i / (C / 8);
  const int c = static_cast<int>(i - row * (C / 8)) * 8;
  const uint4 h = *reinterpret_cast<const uint4*>(a_x + row * 2 * C + c), l = *reinterpret_cast<const uint4*>(a_x + row * 2 * C + C + c);
  const float4 p0 = *reinterpret_cast<const float4*>(pe + row * C + c), p1 = *reinterpret_cast<const float4*>(pe + row * C + c + 4);
  const uint32_t hs[4] = {h.x, h.y, h.z, h.w}, ls[4] = {l.x, l.y, l.z, l.w};
  const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
  float v[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = __uint_as_float(hs[k] << 16) + __uint_as_float(ls[k] << 16) + pv[2 * k];
    v[2 * k + 1] = __uint_as_float(hs[k] & 0xFFFF0000u) + __uint_as_float(ls[k] & 0xFFFF0000u) + pv[2 * k + 1];
  }
  store_split8(a_sum + row * 2 * C + c, C, v);
}

// ---------------------------------------------------------------------------------------------
// x_out = LayerNorm(x_in + y) (post-norm residual, reference transformer_parq.py:375-376, 381-385),
// eps 1e-5, biased variance.  One warp per row.  Also emits the bf16 split of x_out and, when pe is
// given, of x_out + pe (the cross-attention query input, transformer_parq.py:377).
// ---------------------------------------------------------------------------------------------
// The residual input is either fp32 (x_in) or, when x_in is null, the [hi|lo] bf16 split x_in_split (R, 2C) -- the form in
// which the sampled features leave the sampling kernel (hi + lo carries 16 mantissa bits, exactly what the GEMMs saw).
template <int C>
__global__ void __launch_bounds__(256)
add_ln_kernel(const float* __restrict__ x_in, const __nv_bfloat16* __restrict__ x_in_split, const float* __restrict__ y,
              const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ pe,
              float* __restrict__ x_out, __nv_bfloat16* __restrict__ a_x, __nv_bfloat16* __restrict__ a_xpe, int R) {
  constexpr int PASSES = C / 256;          // a lane owns 8 consecutive channels per pass
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const long long base = static_cast<long long>(row) * C;
  float v[PASSES][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PASSES; ++i) {
    const int c = (i * 32 + lane) * 8;
    float4 a0, a1;
    if (x_in != nullptr) {
      a0 = *reinterpret_cast<const float4*>(x_in + base + c);
      a1 = *reinterpret_cast<const float4*>(x_in + base + c + 4);
    } else {
      const uint4 h = *reinterpret_cast<const uint4*>(x_in_split + 2 * base + c), l = *reinterpret_cast<const uint4*>(x_in_split + 2 * base + C + c);
      a0 = make_float4(__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16), __uint_as_float(h.x & 0xFFFF0000u) + __uint_as_float(l.x & 0xFFFF0000u),
                       __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16), __uint_as_float(h.y & 0xFFFF0000u) + __uint_as_float(l.y & 0xFFFF0000u));
      a1 = make_float4(__uint_as_float(h.z << 16) + __uint_as_float(l.z << 16), __uint_as_float(h.z & 0xFFFF0000u) + __uint_as_float(l.z & 0xFFFF0000u),
                       __uint_as_float(h.w << 16) + __uint_as_float(l.w << 16), __uint_as_float(h.w & 0xFFFF0000u) + __uint_as_float(l.w & 0xFFFF0000u));
    }
    const float4 b0 = *reinterpret_cast<const float4*>(y + base + c), b1 = *reinterpret_cast<const float4*>(y + base + c + 4);
    v[i][0] = a0.x + b0.x; v[i][1] = a0.y + b0.y; v[i][2] = a0.z + b0.z; v[i][3] = a0.w + b0.w;
    v[i][4] = a1.x + b1.x; v[i][5] = a1.y + b1.y; v[i][6] = a1.z + b1.z; v[i][7] = a1.w + b1.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[i][k];
  }
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < PASSES; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; ss += d * d; }
  const float rstd = 1.f / sqrtf(warp_sum(ss) / C + 1e-5f);
#pragma unroll
  for (int i = 0; i < PASSES; ++i) {
    const int c = (i * 32 + lane) * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 e0 = *reinterpret_cast<const float4*>(beta + c), e1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * g[k] + be[k];
    *reinterpret_cast<float4*>(x_out + base + c) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(x_out + base + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
    if (a_x != nullptr) store_split8(a_x + static_cast<long long>(row) * (2 * C) + c, C, o);
    if (a_xpe != nullptr) {
      const float4 p0 = *reinterpret_cast<const float4*>(pe + base + c), p1 = *reinterpret_cast<const float4*>(pe + base + c + 4);
      const float op[8] = {o[0] + p0.x, o[1] + p0.y, o[2] + p0.z, o[3] + p0.w, o[4] + p1.x, o[5] + p1.y, o[6] + p1.z, o[7] + p1.w};
      store_split8(a_xpe + static_cast<long long>(row) * (2 * C) + c, C, op);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm(1, C) of the 3-layer heads (reference generic_mlp.py:85-86): statistics over the whole
// (C x Nq) block of one clip.  Two phases, deterministic:
//   GEMM epilogue   : per-tile (sum, sumsq) partials in double (fixed slots -> deterministic)
//   gn_apply_kernel : reduces the partials in fixed order, applies affine + ReLU, emits bf16 split
// h: (B*Nq, ldh) fp32, group g occupies columns [g*C, (g+1)*C).
// ---------------------------------------------------------------------------------------------
// Statistics come from the GEMM epilogues (gemm_tc.cuh, gemm_sk.cuh, chain_tc.cuh) as (sum, sumsq) double2 slots of
// GN_SLOT_COLS output columns per 128-row tile (ptx.cuh); group g (centre / rotation head) occupies the slots of columns
// [g*C, (g+1)*C).  A clip owns Nq/128 consecutive m-tiles.

// One WARP reduces the slots of (clip b, group g): lanes stride over the slots, then a butterfly in double (fixed order:
// deterministic); every lane returns the result.
__device__ __forceinline__ void gn_mean_rstd(const double2* partial, int b, int g, int C, int Nq, int lane, float& mean, float& rstd) {
  double s = 0.0, ss = 0.0;
  const int mt = Nq / 128, nt = C / GN_SLOT_COLS;
  for (int i = lane; i < mt * nt; i += 32) {
    const int m = i / nt, n = i - m * nt;
    const double2 v = partial[static_cast<long long>(b * mt + m) * GN_SLOTS_PER_MTILE + g * nt + n];
    s += v.x;
    ss += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const double cnt = static_cast<double>(C) * Nq;
  const double m = s / cnt;
  const double var = fmax(ss / cnt - m * m, 0.0);
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + 1e-5));
}

// out: (B*Nq, groups * 2C): group g's [hi | lo] at columns [g*2C, (g+1)*2C).  One block per GN_ROWS consecutive rows (they
// belong to one clip: Nq % GN_ROWS == 0), so the double-precision reduction of the tile sums runs once per block instead of once
// per row (it is a serial prologue of ~2-3 k cycles); a thread owns 8 consecutive channels of every row of the block.
template <int GN_ROWS>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ h, int ldh, int C, int Nq, int groups, const double2* __restrict__ partial,
                const float* __restrict__ gamma0, const float* __restrict__ beta0, const float* __restrict__ gamma1,
                const float* __restrict__ beta1, __nv_bfloat16* __restrict__ out, int R) {
  pdl_wait();
  pdl_launch_dependents();
  const int row0 = blockIdx.x * GN_ROWS;
  const int b = row0 / Nq;
  const int per_group = C / 8;
  __shared__ float s_stat[2][2];
  if ((threadIdx.x >> 5) < groups) {
    float mean, rstd;
    gn_mean_rstd(partial, b, threadIdx.x >> 5, C, Nq, threadIdx.x & 31, mean, rstd);
    if ((threadIdx.x & 31) == 0) { s_stat[threadIdx.x >> 5][0] = mean; s_stat[threadIdx.x >> 5][1] = rstd; }
  }
  __syncthreads();
  for (int item = threadIdx.x; item < groups * per_group; item += blockDim.x) {
    const int g = item / per_group, c = (item % per_group) * 8;
    const float mean = s_stat[g][0], rstd = s_stat[g][1];
    const float* gamma = g == 0 ? gamma0 : gamma1;
    const float* beta = g == 0 ? beta0 : beta1;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 e0 = *reinterpret_cast<const float4*>(beta + c), e1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float4 v0[GN_ROWS], v1[GN_ROWS];
#pragma unroll
    for (int r = 0; r < GN_ROWS; ++r) {                    // all loads of the block's rows in flight together
      const float* src = h + static_cast<long long>(min(row0 + r, R - 1)) * ldh + g * C + c;
      v0[r] = *reinterpret_cast<const float4*>(src);
      v1[r] = *reinterpret_cast<const float4*>(src + 4);
    }
#pragma unroll
    for (int r = 0; r < GN_ROWS; ++r) {
      if (row0 + r >= R) break;
      const float v[8] = {v0[r].x, v0[r].y, v0[r].z, v0[r].w, v1[r].x, v1[r].y, v1[r].z, v1[r].w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaxf((v[k] - mean) * rstd * gg[k] + be[k], 0.f);
      store_split8(out + static_cast<long long>(row0 + r) * (groups * 2 * C) + g * 2 * C + c, C, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Final layer of the four heads + box update, one warp per query
// (reference transformer_parq.py:236-279, utils/parq_utils.py:90-105):
//   cls    = Wcls x + b             (10)      size_s = Wsz x + b   (3)
//   c_off  = Wc3 relu(GN(h2c)) + b  (3)       ortho6d = Wr3 relu(GN(h2r)) + b (6)
//   coord_pos = denorm(ref);  center = denorm(sigmoid(c_off + inverse_sigmoid(ref)))
//   prob = softmax(cls);  size = exp(size_s) * mean_size[argmax prob];  ref_next = norm(center)
// ---------------------------------------------------------------------------------------------
struct HeadsParams {
  const float* x;          // (R, C)   decoder-layer output (after LN3)
  const float* h2;         // (R, 2C)  pre-GroupNorm hidden of layer 2: center | rotation
  const double2* partial;  // GroupNorm tile sums of h2 written by the GEMM epilogue
  const float *gamma_c, *beta_c, *gamma_r, *beta_r;
  const float *w_cls, *b_cls, *w_size, *b_size, *w_c3, *b_c3, *w_r3, *b_r3;   // (n, C) row-major, fp32
  const float* ref;        // (R, 3) normalised reference points of this iteration
  const float* mean_size;  // (num_cls, 3)
  float *logits, *center, *size, *ortho6d, *prob, *coord_pos;   // outputs of this iteration
  float* ref_next;         // (R, 3)
  __nv_bfloat16* posemb_next;   // optional (R, 768) [hi|lo]: pos2posemb3d of ref_next for the next iteration (saves its launch)
  const float* dim_t;      // (128) with posemb_next
  float* rot;              // optional (R, 9): rotation matrix of ortho6d (utils/ortho6d_transforms.py:53-66)
  int R, Nq, C, num_cls;
  float span[3], lo[3];
};

constexpr int HEADS_CLS_SLOTS = 16;                              // class logits padded to 16 output slots
constexpr int HEADS_SLOTS = HEADS_CLS_SLOTS + 3 + 3 + 6;         // + size, centre offset, ortho6d = 28 (<= 32 lanes)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Epilogue of one row: lane j holds output slot j (bias not yet added).
template <bool kPosemb>
__device__ __forceinline__ void heads_row_epilogue(const HeadsParams& p, int row, int lane, float mine) {
  const bool is_cls = lane < p.num_cls;
  float ref_n = 0.f;                         // next reference coordinate (centre lanes)
  const int k3 = lane >= HEADS_CLS_SLOTS + 3 ? lane - HEADS_CLS_SLOTS - 3 : lane - HEADS_CLS_SLOTS;   // axis for size / centre lanes
  if (is_cls) mine += p.b_cls[lane];
  else if (lane >= HEADS_CLS_SLOTS && lane < HEADS_CLS_SLOTS + 3) mine += p.b_size[lane - HEADS_CLS_SLOTS];
  else if (lane >= HEADS_CLS_SLOTS + 3 && lane < HEADS_CLS_SLOTS + 6) mine += p.b_c3[lane - HEADS_CLS_SLOTS - 3];
  else if (lane >= HEADS_CLS_SLOTS + 6 && lane < HEADS_SLOTS) mine += p.b_r3[lane - HEADS_CLS_SLOTS - 6];
  // softmax over the class lanes; arg-max of the probabilities with torch's first-index tie rule
  const float mx = warp_max(is_cls ? mine : -INFINITY);
  const float e = is_cls ? expf(mine - mx) : 0.f;
  const float den = warp_sum(e);
  const float pr = e / den;
  float best = is_cls ? pr : -1.f;
  int arg = is_cls ? lane : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (is_cls) {
    p.prob[static_cast<long long>(row) * p.num_cls + lane] = pr;
    p.logits[static_cast<long long>(row) * p.num_cls + lane] = mine;
  }
  if (lane >= HEADS_CLS_SLOTS && lane < HEADS_CLS_SLOTS + 3) {
    p.size[row * 3 + k3] = expf(mine) * p.mean_size[arg * 3 + k3];
  } else if (lane >= HEADS_CLS_SLOTS + 3 && lane < HEADS_CLS_SLOTS + 6) {
    const float rr = p.ref[row * 3 + k3];
    const float rc = fminf(fmaxf(rr, 0.f), 1.f);
    const float inv_sig = logf(fmaxf(rc, 1e-3f) / fmaxf(1.f - rc, 1e-3f));
    const float sg = 1.f / (1.f + expf(-(mine + inv_sig)));
    const float center = __fadd_rn(__fmul_rn(sg, p.span[k3]), p.lo[k3]);
    p.center[row * 3 + k3] = center;
    p.coord_pos[row * 3 + k3] = __fadd_rn(__fmul_rn(rr, p.span[k3]), p.lo[k3]);
    ref_n = __fdiv_rn(__fadd_rn(center, -p.lo[k3]), p.span[k3]);
    p.ref_next[row * 3 + k3] = ref_n;
  } else if (lane >= HEADS_CLS_SLOTS + 6 && lane < HEADS_SLOTS) {
    p.ortho6d[row * 6 + lane - HEADS_CLS_SLOTS - 6] = mine;
  }
  if (kPosemb && p.posemb_next != nullptr) {
    // sinusoidal embedding of the next reference point, the input of the next iteration's reference-point MLP
    const float rx = __shfl_sync(0xffffffffu, ref_n, HEADS_CLS_SLOTS + 3), ry = __shfl_sync(0xffffffffu, ref_n, HEADS_CLS_SLOTS + 4),
                rz = __shfl_sync(0xffffffffu, ref_n, HEADS_CLS_SLOTS + 5);
    __nv_bfloat16* out = p.posemb_next + static_cast<long long>(row) * 768;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int j = k * 32 + lane;
      store_split(out + j, 384, posemb_value(rx, ry, rz, j, p.dim_t));
    }
  }
  if (p.rot != nullptr) {
    // Gram-Schmidt: x = a/|a|, z = (x X b)/|x X b|, y = z X x; columns [x y z]; norms clamped at 1e-8
    float o6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) o6[k] = __shfl_sync(0xffffffffu, mine, HEADS_CLS_SLOTS + 6 + k);
    if (lane == 0) {
      const float na = fmaxf(sqrtf(o6[0] * o6[0] + o6[1] * o6[1] + o6[2] * o6[2]), 1e-8f);
      const float x0 = o6[0] / na, x1 = o6[1] / na, x2 = o6[2] / na;
      float z0 = x1 * o6[5] - x2 * o6[4], z1 = x2 * o6[3] - x0 * o6[5], z2 = x0 * o6[4] - x1 * o6[3];
      const float nz = fmaxf(sqrtf(z0 * z0 + z1 * z1 + z2 * z2), 1e-8f);
      z0 /= nz; z1 /= nz; z2 /= nz;
      const float y0 = z1 * x2 - z2 * x1, y1 = z2 * x0 - z0 * x2, y2 = z0 * x1 - z1 * x0;
      float* r = p.rot + static_cast<long long>(row) * 9;
      r[0] = x0; r[1] = y0; r[2] = z0;
      r[3] = x1; r[4] = y1; r[5] = z1;
      r[6] = x2; r[7] = y2; r[8] = z2;
    }
  }
}

// 1-D bulk copy global -> shared (TMA engine, no tensor map), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Sum over the 32 lanes of 32 per-lane values so that lane j ends up with the total of slot j: a halving exchange
// (16 + 8 + 4 + 2 + 1 = 31 shuffles) instead of 32 full butterflies (160).
__device__ __forceinline__ float reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int k = 0; k < w; ++k) {
      const float keep = up ? v[w + k] : v[k];
      const float send = up ? v[k] : v[w + k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}

// Block = 8 warps.  The 28 x C final-layer weights (zero rows for unused class slots) and the GroupNorm affine parameters
// are staged once per block in shared memory by bulk copies issued before the dependency wait (they are constants); the
// GroupNorm statistics of the (at most two) clips the block's rows belong to are reduced once per block.  A warp then takes
// FOUR rows (queries) at a time: every 16-byte weight read from shared memory feeds four rows (with two, the kernel was
// bound by the shared-memory bandwidth: 28 rows x 112 KB of weights per SM), 4 x 28 independent dot-product chains per
// lane, a halving reduction that leaves output slot j in lane j, and an epilogue with softmax / arg-max through shuffles.
// kPosemb: also write pos2posemb3d of the next reference point (HeadsParams::posemb_next).
constexpr int HEADS_THREADS = 256;
constexpr int HEADS_NR = 4;
template <int C>
constexpr size_t heads_smem_bytes() { return static_cast<size_t>(HEADS_SLOTS + 4) * C * sizeof(float) + 64; }

template <int C, bool kPosemb>
__global__ void __launch_bounds__(HEADS_THREADS)
heads_final_kernel(const HeadsParams p, int rows_per_block) {
  constexpr int NR = HEADS_NR;
  extern __shared__ __align__(128) float sw[];           // [HEADS_SLOTS][C] weights, then gamma_c | beta_c | gamma_r | beta_r
  float* s_aff = sw + HEADS_SLOTS * C;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_aff + 4 * C);
  float* s_stat = reinterpret_cast<float*>(bar + 1);      // [2 clips][2 groups][mean, rstd]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    const uint32_t row_b = C * sizeof(float);
    mbar_expect_tx(bar, static_cast<uint32_t>(p.num_cls + 3 + 3 + 6 + 4) * row_b);
    bulk_g2s(sw, p.w_cls, p.num_cls * row_b, bar);
    bulk_g2s(sw + HEADS_CLS_SLOTS * C, p.w_size, 3 * row_b, bar);
    bulk_g2s(sw + (HEADS_CLS_SLOTS + 3) * C, p.w_c3, 3 * row_b, bar);
    bulk_g2s(sw + (HEADS_CLS_SLOTS + 6) * C, p.w_r3, 6 * row_b, bar);
    bulk_g2s(s_aff, p.gamma_c, row_b, bar);
    bulk_g2s(s_aff + C, p.beta_c, row_b, bar);
    bulk_g2s(s_aff + 2 * C, p.gamma_r, row_b, bar);
    bulk_g2s(s_aff + 3 * C, p.beta_r, row_b, bar);
  }
  for (int i = p.num_cls * C + threadIdx.x * 4; i < HEADS_CLS_SLOTS * C; i += HEADS_THREADS * 4)     // unused class slots
    *reinterpret_cast<float4*>(sw + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  // everything above is constant weights: staged while the previous kernel drains
  pdl_wait();
  pdl_launch_dependents();
  const int row_begin = blockIdx.x * rows_per_block;
  const int row_end = min(p.R, row_begin + rows_per_block);
  const int clip0 = row_begin / p.Nq;
  if (warp < 4) {                                           // (clip, group) = (warp >> 1, warp & 1)
    const int clip = min(clip0 + (warp >> 1), (p.R - 1) / p.Nq);
    float mean, rstd;
    gn_mean_rstd(p.partial, clip, warp & 1, C, p.Nq, lane, mean, rstd);
    if (lane == 0) { s_stat[warp * 2] = mean; s_stat[warp * 2 + 1] = rstd; }
  }
  __syncthreads();
  mbar_wait(bar, 0);
  for (int row0 = row_begin + NR * warp; row0 < row_end; row0 += NR * (HEADS_THREADS >> 5)) {
    float mean_c[NR], rstd_c[NR], mean_r[NR], rstd_r[NR];
    long long xoff[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int row = min(row0 + r, row_end - 1);          // missing rows repeat the last one (results discarded)
      xoff[r] = static_cast<long long>(row) * C;
      const float* st = s_stat + (row / p.Nq - clip0) * 4;
      mean_c[r] = st[0]; rstd_c[r] = st[1]; mean_r[r] = st[2]; rstd_r[r] = st[3];
    }
    float acc[NR][32];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[r][j] = 0.f;
    // a lane owns 4 consecutive channels per step (16-byte global and shared loads); the 12 loads of a step are issued
    // together, the inputs are converted in place (x | relu(GN(h2 centre)) | relu(GN(h2 rotation)))
    constexpr int STEPS = C / 128;
#pragma unroll 1
    for (int i = 0; i < STEPS; ++i) {
      const int c = i * 128 + lane * 4;
      float4 vx[NR], vc[NR], vr[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        vx[r] = *reinterpret_cast<const float4*>(p.x + xoff[r] + c);
        vc[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xoff[r] + c);
        vr[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xoff[r] + C + c);
      }
#pragma unroll
      for (int j = 0; j < HEADS_CLS_SLOTS + 3; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sw + j * C + c);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          acc[r][j] = fmaf(w.x, vx[r].x, acc[r][j]);
          acc[r][j] = fmaf(w.y, vx[r].y, acc[r][j]);
          acc[r][j] = fmaf(w.z, vx[r].z, acc[r][j]);
          acc[r][j] = fmaf(w.w, vx[r].w, acc[r][j]);
        }
      }
      {
        const float4 g = *reinterpret_cast<const float4*>(s_aff + c), b = *reinterpret_cast<const float4*>(s_aff + C + c);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          vc[r].x = fmaxf((vc[r].x - mean_c[r]) * rstd_c[r] * g.x + b.x, 0.f);
          vc[r].y = fmaxf((vc[r].y - mean_c[r]) * rstd_c[r] * g.y + b.y, 0.f);
          vc[r].z = fmaxf((vc[r].z - mean_c[r]) * rstd_c[r] * g.z + b.z, 0.f);
          vc[r].w = fmaxf((vc[r].w - mean_c[r]) * rstd_c[r] * g.w + b.w, 0.f);
        }
      }
#pragma unroll
      for (int j = HEADS_CLS_SLOTS + 3; j < HEADS_CLS_SLOTS + 6; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sw + j * C + c);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          acc[r][j] = fmaf(w.x, vc[r].x, acc[r][j]);
          acc[r][j] = fmaf(w.y, vc[r].y, acc[r][j]);
          acc[r][j] = fmaf(w.z, vc[r].z, acc[r][j]);
          acc[r][j] = fmaf(w.w, vc[r].w, acc[r][j]);
        }
      }
      {
        const float4 g = *reinterpret_cast<const float4*>(s_aff + 2 * C + c), b = *reinterpret_cast<const float4*>(s_aff + 3 * C + c);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          vr[r].x = fmaxf((vr[r].x - mean_r[r]) * rstd_r[r] * g.x + b.x, 0.f);
          vr[r].y = fmaxf((vr[r].y - mean_r[r]) * rstd_r[r] * g.y + b.y, 0.f);
          vr[r].z = fmaxf((vr[r].z - mean_r[r]) * rstd_r[r] * g.z + b.z, 0.f);
          vr[r].w = fmaxf((vr[r].w - mean_r[r]) * rstd_r[r] * g.w + b.w, 0.f);
        }
      }
#pragma unroll
      for (int j = HEADS_CLS_SLOTS + 6; j < HEADS_SLOTS; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sw + j * C + c);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          acc[r][j] = fmaf(w.x, vr[r].x, acc[r][j]);
          acc[r][j] = fmaf(w.y, vr[r].y, acc[r][j]);
          acc[r][j] = fmaf(w.z, vr[r].z, acc[r][j]);
          acc[r][j] = fmaf(w.w, vr[r].w, acc[r][j]);
        }
      }
    }
    float mine_r[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) mine_r[r] = reduce_scatter32(acc[r], lane);     // lane j keeps output slot j
#pragma unroll 1
    for (int r = 0; r < NR && row0 + r < row_end; ++r) {      // one copy of the epilogue code
      const float mine = r == 0 ? mine_r[0] : (r == 1 ? mine_r[1] : (r == 2 ? mine_r[2] : mine_r[3]));
      heads_row_epilogue<kPosemb>(p, row0 + r, lane, mine);
    }
  }
}

// The same final layer for a FEW rows (one clip: 256 queries), where the kernel above would stage 128 KB of weights per block to
// serve two rows.  Block = 8 warps for HEADS_SMALL_ROWS rows; warp w owns channels [128 w, 128 w + 128) and keeps ITS slice of
// the 28 weight rows in registers (loaded before the dependency wait, every warp's 28 loads in flight at once, no shared-memory
// round trip); per row a lane forms 28 four-channel partial dot products, the halving reduction leaves slot j in lane j, the
// eight warps' partials meet in shared memory and warp r finishes row r with the common epilogue.
constexpr int HEADS_SMALL_ROWS = 2;
template <int C, bool kPosemb>
__global__ void __launch_bounds__(256)
heads_final_small_kernel(const HeadsParams p) {
  static_assert(C == 1024, "8 warps x 32 lanes x 4 channels");
  __shared__ float s_part[HEADS_SMALL_ROWS][8][32];
  __shared__ float s_stat[4];                              // mean / rstd of the centre and rotation groups of the block's clip
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = warp * 128 + lane * 4;
  float4 w[HEADS_SLOTS];
#pragma unroll
  for (int j = 0; j < HEADS_SLOTS; ++j) {
    const float* src = nullptr;
    if (j < HEADS_CLS_SLOTS) src = j < p.num_cls ? p.w_cls + j * C : nullptr;
    else if (j < HEADS_CLS_SLOTS + 3) src = p.w_size + (j - HEADS_CLS_SLOTS) * C;
    else if (j < HEADS_CLS_SLOTS + 6) src = p.w_c3 + (j - HEADS_CLS_SLOTS - 3) * C;
    else src = p.w_r3 + (j - HEADS_CLS_SLOTS - 6) * C;
    w[j] = src != nullptr ? __ldg(reinterpret_cast<const float4*>(src + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 gc = __ldg(reinterpret_cast<const float4*>(p.gamma_c + c)), bc = __ldg(reinterpret_cast<const float4*>(p.beta_c + c));
  const float4 gr = __ldg(reinterpret_cast<const float4*>(p.gamma_r + c)), br = __ldg(reinterpret_cast<const float4*>(p.beta_r + c));
  // everything above is constant weights: in flight while the previous kernel drains
  pdl_wait();
  pdl_launch_dependents();
  const int row_begin = blockIdx.x * HEADS_SMALL_ROWS;
  float4 vx[HEADS_SMALL_ROWS], vc[HEADS_SMALL_ROWS], vr[HEADS_SMALL_ROWS];
#pragma unroll
  for (int r = 0; r < HEADS_SMALL_ROWS; ++r) {
    const long long xo = static_cast<long long>(min(row_begin + r, p.R - 1)) * C;
    vx[r] = *reinterpret_cast<const float4*>(p.x + xo + c);
    vc[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xo + c);
    vr[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xo + C + c);
  }
  if (warp < 2) {                                           // the rows of a block belong to one clip (Nq % HEADS_SMALL_ROWS == 0)
    float mean, rstd;
    gn_mean_rstd(p.partial, row_begin / p.Nq, warp, C, p.Nq, lane, mean, rstd);
    if (lane == 0) { s_stat[warp * 2] = mean; s_stat[warp * 2 + 1] = rstd; }
  }
  __syncthreads();
  const float mean_c = s_stat[0], rstd_c = s_stat[1], mean_r = s_stat[2], rstd_r = s_stat[3];
#pragma unroll
  for (int r = 0; r < HEADS_SMALL_ROWS; ++r) {
    vc[r].x = fmaxf((vc[r].x - mean_c) * rstd_c * gc.x + bc.x, 0.f);
    vc[r].y = fmaxf((vc[r].y - mean_c) * rstd_c * gc.y + bc.y, 0.f);
    vc[r].z = fmaxf((vc[r].z - mean_c) * rstd_c * gc.z + bc.z, 0.f);
    vc[r].w = fmaxf((vc[r].w - mean_c) * rstd_c * gc.w + bc.w, 0.f);
    vr[r].x = fmaxf((vr[r].x - mean_r) * rstd_r * gr.x + br.x, 0.f);
    vr[r].y = fmaxf((vr[r].y - mean_r) * rstd_r * gr.y + br.y, 0.f);
    vr[r].z = fmaxf((vr[r].z - mean_r) * rstd_r * gr.z + br.z, 0.f);
    vr[r].w = fmaxf((vr[r].w - mean_r) * rstd_r * gr.w + br.w, 0.f);
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < HEADS_SLOTS) {
        const float4 v = j < HEADS_CLS_SLOTS + 3 ? vx[r] : (j < HEADS_CLS_SLOTS + 6 ? vc[r] : vr[r]);
        acc[j] = fmaf(w[j].w, v.w, fmaf(w[j].z, v.z, fmaf(w[j].y, v.y, w[j].x * v.x)));
      } else {
        acc[j] = 0.f;
      }
    }
    s_part[r][warp][lane] = reduce_scatter32(acc, lane);    // lane j: this warp's share of output slot j
  }
  __syncthreads();
  if (warp < HEADS_SMALL_ROWS && row_begin + warp < p.R) {
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) mine += s_part[warp][k][lane];          // fixed order: deterministic
    heads_row_epilogue<kPosemb>(p, row_begin + warp, lane, mine);
  }
}

}  // namespace parq
