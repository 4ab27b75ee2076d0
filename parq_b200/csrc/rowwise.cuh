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
__device__ __forceinline__ float posemb_value(float rx, float ry, float rz, int j, const float* __restrict__ dim_t) {
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
// Statistics come from the GEMM epilogue (gemm_tc.cuh): one (sum, sumsq) double2 per 128x256 output tile in
// slot m_tile*8 + n_tile, where n_tile = g*(C/256) + tile-in-group.  A clip owns Nq/128 consecutive m-tiles.
constexpr int GN_SLOTS_PER_MTILE = 8;

__device__ __forceinline__ void gn_mean_rstd(const double2* partial, int b, int g, int C, int Nq, float& mean, float& rstd) {
  double s = 0.0, ss = 0.0;
  const int mt = Nq / 128, nt = C / 256;
  for (int m = 0; m < mt; ++m)
    for (int n = 0; n < nt; ++n) {
      const double2 v = partial[static_cast<long long>(b * mt + m) * GN_SLOTS_PER_MTILE + g * nt + n];
      s += v.x;
      ss += v.y;
    }
  const double cnt = static_cast<double>(C) * Nq;
  const double m = s / cnt;
  const double var = fmax(ss / cnt - m * m, 0.0);
  mean = static_cast<float>(m);
  rstd = static_cast<float>(1.0 / sqrt(var + 1e-5));
}

// out: (B*Nq, groups * 2C): group g's [hi | lo] at columns [g*2C, (g+1)*2C).  One block per row;
// a thread owns 8 consecutive channels (C/8 threads per group).
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ h, int ldh, int C, int Nq, int groups, const double2* __restrict__ partial,
                const float* __restrict__ gamma0, const float* __restrict__ beta0, const float* __restrict__ gamma1,
                const float* __restrict__ beta1, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;           // b*Nq + q
  const int b = row / Nq;
  const int per_group = C / 8;
  __shared__ float s_stat[2][2];
  if (threadIdx.x < groups) gn_mean_rstd(partial, b, threadIdx.x, C, Nq, s_stat[threadIdx.x][0], s_stat[threadIdx.x][1]);
  __syncthreads();
  for (int item = threadIdx.x; item < groups * per_group; item += blockDim.x) {
    const int g = item / per_group, c = (item % per_group) * 8;
    const float mean = s_stat[g][0], rstd = s_stat[g][1];
    const float* gamma = g == 0 ? gamma0 : gamma1;
    const float* beta = g == 0 ? beta0 : beta1;
    const float* src = h + static_cast<long long>(row) * ldh + g * C + c;
    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 e0 = *reinterpret_cast<const float4*>(beta + c), e1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaxf((v[k] - mean) * rstd * gg[k] + be[k], 0.f);
    store_split8(out + static_cast<long long>(row) * (groups * 2 * C) + g * 2 * C + c, C, o);
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

// Block = 16 warps.  The 28 x C final-layer weights (zero rows for unused class slots) and the GroupNorm
// affine parameters are staged once per block in shared memory (before the dependency wait: they are
// constants); each warp then walks PAIRS of rows (queries) so that every weight read from shared memory
// feeds two rows: 2 x 28 independent dot-product chains per lane, a butterfly reduction, and an epilogue
// in which lane j owns output slot j (softmax/arg-max through warp shuffles).
// kPosemb: also write pos2posemb3d of the next reference point (HeadsParams::posemb_next); a separate instantiation so that
// the common one keeps its register budget (the extra code spills under the 128-register cap of 512 threads).
template <int C, bool kPosemb>
__global__ void __launch_bounds__(512)
heads_final_kernel(const HeadsParams p, int rows_per_block) {
  constexpr int NR = 2;
  extern __shared__ float sw[];           // [HEADS_SLOTS][C] weights, then gamma_c | beta_c | gamma_r | beta_r
  float* s_aff = sw + HEADS_SLOTS * C;
  for (int i = threadIdx.x * 4; i < HEADS_SLOTS * C; i += blockDim.x * 4) {
    const int j = i / C, c = i % C;
    const float* src = nullptr;
    if (j < HEADS_CLS_SLOTS) src = j < p.num_cls ? p.w_cls + j * C : nullptr;
    else if (j < HEADS_CLS_SLOTS + 3) src = p.w_size + (j - HEADS_CLS_SLOTS) * C;
    else if (j < HEADS_CLS_SLOTS + 6) src = p.w_c3 + (j - HEADS_CLS_SLOTS - 3) * C;
    else src = p.w_r3 + (j - HEADS_CLS_SLOTS - 6) * C;
    *reinterpret_cast<float4*>(sw + i) = src ? *reinterpret_cast<const float4*>(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_aff[i] = p.gamma_c[i]; s_aff[C + i] = p.beta_c[i]; s_aff[2 * C + i] = p.gamma_r[i]; s_aff[3 * C + i] = p.beta_r[i];
  }
  // everything above is constant weights: staged while the previous kernel drains
  pdl_wait();
  pdl_launch_dependents();
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row_end = min(p.R, (blockIdx.x + 1) * rows_per_block);
  for (int row0 = blockIdx.x * rows_per_block + NR * (threadIdx.x >> 5); row0 < row_end; row0 += NR * (blockDim.x >> 5)) {
    float mean_c[NR], rstd_c[NR], mean_r[NR], rstd_r[NR];
    long long xoff[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int row = min(row0 + r, row_end - 1);          // a missing second row repeats the first (result discarded)
      xoff[r] = static_cast<long long>(row) * C;
      float mean_l = 0.f, rstd_l = 0.f;
      if (lane < 2) gn_mean_rstd(p.partial, row / p.Nq, lane, C, p.Nq, mean_l, rstd_l);
      mean_c[r] = __shfl_sync(0xffffffffu, mean_l, 0); rstd_c[r] = __shfl_sync(0xffffffffu, rstd_l, 0);
      mean_r[r] = __shfl_sync(0xffffffffu, mean_l, 1); rstd_r[r] = __shfl_sync(0xffffffffu, rstd_l, 1);
    }
    float acc[NR][HEADS_SLOTS];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < HEADS_SLOTS; ++j) acc[r][j] = 0.f;
    // a lane owns 4 consecutive channels per step (16-byte global and shared loads); the loads of step i+1 are
    // issued before the arithmetic of step i
    constexpr int STEPS = C / 128;
    float4 nx[NR], na[NR], nq[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      nx[r] = *reinterpret_cast<const float4*>(p.x + xoff[r] + lane * 4);
      na[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xoff[r] + lane * 4);
      nq[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xoff[r] + C + lane * 4);
    }
#pragma unroll 1
    for (int i = 0; i < STEPS; ++i) {
      const int c = i * 128 + lane * 4;
      float4 cx[NR], ca[NR], cq[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) { cx[r] = nx[r]; ca[r] = na[r]; cq[r] = nq[r]; }
      if (i + 1 < STEPS) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          nx[r] = *reinterpret_cast<const float4*>(p.x + xoff[r] + c + 128);
          na[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xoff[r] + c + 128);
          nq[r] = *reinterpret_cast<const float4*>(p.h2 + 2 * xoff[r] + C + c + 128);
        }
      }
      const float4 gc = *reinterpret_cast<const float4*>(s_aff + c), bc = *reinterpret_cast<const float4*>(s_aff + C + c);
      const float4 gr = *reinterpret_cast<const float4*>(s_aff + 2 * C + c), br = *reinterpret_cast<const float4*>(s_aff + 3 * C + c);
      float xv[NR][4], hc[NR][4], hr[NR][4];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        xv[r][0] = cx[r].x; xv[r][1] = cx[r].y; xv[r][2] = cx[r].z; xv[r][3] = cx[r].w;
        hc[r][0] = fmaxf((ca[r].x - mean_c[r]) * rstd_c[r] * gc.x + bc.x, 0.f);
        hc[r][1] = fmaxf((ca[r].y - mean_c[r]) * rstd_c[r] * gc.y + bc.y, 0.f);
        hc[r][2] = fmaxf((ca[r].z - mean_c[r]) * rstd_c[r] * gc.z + bc.z, 0.f);
        hc[r][3] = fmaxf((ca[r].w - mean_c[r]) * rstd_c[r] * gc.w + bc.w, 0.f);
        hr[r][0] = fmaxf((cq[r].x - mean_r[r]) * rstd_r[r] * gr.x + br.x, 0.f);
        hr[r][1] = fmaxf((cq[r].y - mean_r[r]) * rstd_r[r] * gr.y + br.y, 0.f);
        hr[r][2] = fmaxf((cq[r].z - mean_r[r]) * rstd_r[r] * gr.z + br.z, 0.f);
        hr[r][3] = fmaxf((cq[r].w - mean_r[r]) * rstd_r[r] * gr.w + br.w, 0.f);
      }
#pragma unroll
      for (int j = 0; j < HEADS_SLOTS; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(sw + j * C + c);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const float* v = j < HEADS_CLS_SLOTS + 3 ? xv[r] : (j < HEADS_CLS_SLOTS + 6 ? hc[r] : hr[r]);
          acc[r][j] = fmaf(w.x, v[0], acc[r][j]);
          acc[r][j] = fmaf(w.y, v[1], acc[r][j]);
          acc[r][j] = fmaf(w.z, v[2], acc[r][j]);
          acc[r][j] = fmaf(w.w, v[3], acc[r][j]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < HEADS_SLOTS; ++j) acc[r][j] += __shfl_xor_sync(0xffffffffu, acc[r][j], o);
      // lane j keeps output slot j
      float mine = 0.f;
#pragma unroll
      for (int j = 0; j < HEADS_SLOTS; ++j)
        if (lane == j) mine = acc[r][j];
      if (row0 + r < row_end) heads_row_epilogue<kPosemb>(p, row0 + r, lane, mine);
    }
  }
}

}  // namespace parq
