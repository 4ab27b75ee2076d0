// K0  pose chain   T_camera_local = T_camera_pseudoCam o (T_world_pseudoCam^-1 o T_world_local)
//                  (reference transformer_parq.py:298-300, utils/wrappers.py:247-257)
// K1  project_sample: for every query reference point and every view
//       local -> camera transform (wrappers.py:260-267), pinhole projection and validity
//       (wrappers.py:502-522), normalised-grid round trip and zero-padded bilinear gather of the
//       C-channel texel (transformer_parq.py:148-152, ATen grid_sampler_2d), sum over ALL views
//       divided by the number of VALID views (transformer_parq.py:156-160),
//     fused with the "+ query positional feature" that forms the attention query/key input
//     (transformer_parq.py:372) and with the bf16 hi/lo split the following GEMMs consume.
//
// Rounding contract (bit-exact center_im / center_valid against the CPU oracle): every operation
// below that feeds center_im or the validity test is an explicitly rounded IEEE fp32 intrinsic in
// the order the reference executes it -- 3-term dot products of the tiny pose matmuls as
// ((a0*b0 + a1*b1) + a2*b2) without FMA, the point transform as fma(p2,r2,fma(p1,r1,p0*r0)) + t,
// then x/z (IEEE divide), *f, +c.  Never compile this file with --use_fast_math.
//
// Memory behaviour: the gather is HBM/L2 bound.  A warp owns 256 consecutive channels of one
// (clip, query); every lane moves 16-byte vectors (8 bf16 channels), so each bilinear corner is a
// fully coalesced 512-byte request and the two horizontally adjacent corners form one contiguous
// 2*C*2-byte segment.  Per-view projection parameters are computed once by lane t and broadcast with
// warp shuffles; four views (16 independent 16-byte loads per lane) are kept in flight.
#pragma once
#include "ptx.cuh"

namespace parq {

__device__ __forceinline__ float dot3_nofma(float a0, float b0, float a1, float b1, float a2, float b2) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}

// out = A o B for 12-float poses (R row-major | t): R = RA RB, t = tA + RA tB
__device__ __forceinline__ void pose_compose(const float* A, const float* B, float* out) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      out[3 * i + j] = dot3_nofma(A[3 * i], B[j], A[3 * i + 1], B[3 + j], A[3 * i + 2], B[6 + j]);
    out[9 + i] = __fadd_rn(A[9 + i], dot3_nofma(A[3 * i], B[9], A[3 * i + 1], B[10], A[3 * i + 2], B[11]));
  }
}
__device__ __forceinline__ void pose_inverse(const float* A, float* out) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) out[3 * i + j] = A[3 * j + i];
    out[9 + i] = -dot3_nofma(A[i], A[9], A[3 + i], A[10], A[6 + i], A[11]);
  }
}

__global__ void pose_chain_kernel(const float* __restrict__ T_cp, const float* __restrict__ T_wp,
                                  const float* __restrict__ T_wl, float* __restrict__ T_cl, int B, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int b = i / T;
  float cp[12], wp[12], wl[12], inv[12], tmp[12], out[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    cp[k] = T_cp[i * 12 + k];
    wp[k] = T_wp[i * 12 + k];
    wl[k] = T_wl[b * 12 + k];
  }
  pose_inverse(wp, inv);
  pose_compose(inv, wl, tmp);
  pose_compose(cp, tmp, out);
#pragma unroll
  for (int k = 0; k < 12; ++k) T_cl[i * 12 + k] = out[k];
}

struct SampleParams {
  const __nv_bfloat16* tokens;   // (B, T, H, W, C) channels-last bf16
  const float* ref;              // (B, Nq, 3) normalised reference points in (0,1)
  const float* T_cl;             // (B, T, 12)
  const float* camera;           // (B, T, 6) [w,h,fx,fy,cx,cy]
  const float* pe;               // (B*Nq, C) query positional feature, or nullptr
  float* feat;                   // (B*Nq, C) fp32 sampled features
  __nv_bfloat16* a_x;            // (B*Nq, 2C) [hi|lo] split of feat, or nullptr
  __nv_bfloat16* a_xpe;          // (B*Nq, 2C) [hi|lo] split of feat+pe, or nullptr
  float* center_im;              // (B, T, Nq, 2) or nullptr
  uint8_t* valid;                // (B, T, Nq)    or nullptr
  float* coord_pos;              // (B, Nq, 3)    or nullptr
  int B, T, H, W, C, Nq;
  float span[3], lo[3];          // denormalisation: p*span + lo
};

struct ViewTap {                 // one view's bilinear footprint
  int x0, y0;                    // floor of the sample position (may be out of range)
  float fx, fy;                  // fractional parts
  int inb;                       // bit0 nw, bit1 ne, bit2 sw, bit3 se in bounds
};

__device__ __forceinline__ uint4 ldg_nc_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ void fma_bf16x8(float (&acc)[8], const uint4& v, float w) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[2 * i] = fmaf(__uint_as_float(u[i] << 16), w, acc[2 * i]);
    acc[2 * i + 1] = fmaf(__uint_as_float(u[i] & 0xFFFF0000u), w, acc[2 * i + 1]);
  }
}

// grid = B*Nq blocks, block = (C/256) warps; warp w owns channels [256w, 256w+256).
__global__ void __launch_bounds__(128)
project_sample_kernel(const SampleParams p) {
  const int bq = blockIdx.x;
  const int b = bq / p.Nq, q = bq % p.Nq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = warp * 256 + lane * 8;

  // reference point -> metres in the local frame: p*span + lo (separately rounded mul, add)
  const float rx = p.ref[bq * 3 + 0], ry = p.ref[bq * 3 + 1], rz = p.ref[bq * 3 + 2];
  const float px = __fadd_rn(__fmul_rn(rx, p.span[0]), p.lo[0]);
  const float py = __fadd_rn(__fmul_rn(ry, p.span[1]), p.lo[1]);
  const float pz = __fadd_rn(__fmul_rn(rz, p.span[2]), p.lo[2]);
  if (threadIdx.x == 0 && p.coord_pos != nullptr) {
    p.coord_pos[bq * 3 + 0] = px;
    p.coord_pos[bq * 3 + 1] = py;
    p.coord_pos[bq * 3 + 2] = pz;
  }

  float tot[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) tot[i] = 0.f;
  int nvalid = 0;
  const float Wm1 = static_cast<float>(p.W - 1), Hm1 = static_cast<float>(p.H - 1);
  const float sx = Wm1 / 2.f, sy = Hm1 / 2.f;          // ATen CPU grid_sampler: scaling = (size-1)/2

  for (int tb = 0; tb < p.T; tb += 32) {
    // ---- lane t projects view tb+t
    const int t = tb + lane;
    ViewTap tap;
    tap.x0 = tap.y0 = 0;
    tap.fx = tap.fy = 0.f;
    tap.inb = 0;
    int is_valid = 0;
    if (t < p.T) {
      const float* Tc = p.T_cl + (static_cast<long long>(b) * p.T + t) * 12;
      const float* cam = p.camera + (static_cast<long long>(b) * p.T + t) * 6;
      float pc[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float a = __fmul_rn(px, Tc[3 * i]);
        a = __fmaf_rn(py, Tc[3 * i + 1], a);
        a = __fmaf_rn(pz, Tc[3 * i + 2], a);
        pc[i] = __fadd_rn(a, Tc[9 + i]);
      }
      const float eps = 1e-3f;
      const bool in_front = pc[2] > eps;
      const float zc = fmaxf(pc[2], eps);
      const float u = __fadd_rn(__fmul_rn(__fdiv_rn(pc[0], zc), cam[2]), cam[4]);
      const float v = __fadd_rn(__fmul_rn(__fdiv_rn(pc[1], zc), cam[3]), cam[5]);
      const float wm1 = __fadd_rn(cam[0], -1.f), hm1 = __fadd_rn(cam[1], -1.f);
      is_valid = in_front && (u >= 0.f) && (u <= wm1) && (v >= 0.f) && (v <= hm1);
      if (warp == 0) {
        const long long o = (static_cast<long long>(b) * p.T + t) * p.Nq + q;
        if (p.center_im != nullptr) {
          p.center_im[o * 2] = u;
          p.center_im[o * 2 + 1] = v;
        }
        if (p.valid != nullptr) p.valid[o] = static_cast<uint8_t>(is_valid);
      }
      // normalised grid and back (transformer_parq.py:148-150 then grid_sampler unnormalize)
      const float gx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, u), Wm1), -1.f);
      const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, v), Hm1), -1.f);
      const float ix = __fmul_rn(__fadd_rn(gx, 1.f), sx);
      const float iy = __fmul_rn(__fadd_rn(gy, 1.f), sy);
      const float x0f = floorf(ix), y0f = floorf(iy);
      tap.fx = ix - x0f;
      tap.fy = iy - y0f;
      // in-bounds tests in float (ix may be huge or NaN), then a safe int conversion
      const bool xw = (x0f >= 0.f) && (x0f <= Wm1), xe = (x0f >= -1.f) && (x0f <= Wm1 - 1.f);
      const bool yn = (y0f >= 0.f) && (y0f <= Hm1), ys = (y0f >= -1.f) && (y0f <= Hm1 - 1.f);
      tap.inb = (xw && yn ? 1 : 0) | (xe && yn ? 2 : 0) | (xw && ys ? 4 : 0) | (xe && ys ? 8 : 0);
      if (tap.inb) {
        tap.x0 = static_cast<int>(x0f);
        tap.y0 = static_cast<int>(y0f);
      }
    }
    nvalid += __popc(__ballot_sync(0xffffffffu, is_valid));

    // ---- all lanes gather.  Only views with at least one in-bounds corner contribute (zero padding):
    // their lane ids are compacted from a ballot, four of them (16 independent 16-byte loads) in flight.
    unsigned live = __ballot_sync(0xffffffffu, tap.inb != 0);
    while (live != 0u) {
      uint4 tex[4][4];
      float wgt[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool have = live != 0u;
        const int src = have ? (__ffs(live) - 1) : 0;
        if (have) live &= live - 1;
        const int x0 = __shfl_sync(0xffffffffu, tap.x0, src);
        const int y0 = __shfl_sync(0xffffffffu, tap.y0, src);
        const float fx = __shfl_sync(0xffffffffu, tap.fx, src);
        const float fy = __shfl_sync(0xffffffffu, tap.fy, src);
        int inb = __shfl_sync(0xffffffffu, tap.inb, src);
        if (!have) inb = 0;
        const float ex = 1.f - fx, sy_ = 1.f - fy;       // distances to east / south (ATen CPU form)
        wgt[k][0] = sy_ * ex;   // nw
        wgt[k][1] = sy_ * fx;   // ne
        wgt[k][2] = fy * ex;    // sw
        wgt[k][3] = fy * fx;    // se
        const long long view = static_cast<long long>(b) * p.T + tb + src;
        const __nv_bfloat16* base = p.tokens + ((view * p.H + y0) * p.W + x0) * static_cast<long long>(p.C) + ch;
        const long long rowpitch = static_cast<long long>(p.W) * p.C;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (inb & (1 << c)) {
            tex[k][c] = ldg_nc_16(base + (c >> 1) * rowpitch + (c & 1) * p.C);
          } else {
            tex[k][c] = make_uint4(0, 0, 0, 0);
            wgt[k][c] = 0.f;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) fma_bf16x8(acc, tex[k][c], wgt[k][c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) tot[i] += acc[i];
      }
    }
  }

  const float cnt = static_cast<float>(max(nvalid, 1));
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = tot[i] / cnt;
  const long long o = static_cast<long long>(bq) * p.C + ch;
  reinterpret_cast<float4*>(p.feat + o)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p.feat + o)[1] = make_float4(f[4], f[5], f[6], f[7]);
  if (p.a_x != nullptr) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      hi[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      lo[i] = pack_bf16x2(f[2 * i] - __uint_as_float(hi[i] << 16), f[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
    }
    __nv_bfloat16* dst = p.a_x + static_cast<long long>(bq) * (2 * p.C) + ch;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + p.C) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (p.a_xpe != nullptr) {
    const float4 e0 = reinterpret_cast<const float4*>(p.pe + o)[0], e1 = reinterpret_cast<const float4*>(p.pe + o)[1];
    const float g[8] = {f[0] + e0.x, f[1] + e0.y, f[2] + e0.z, f[3] + e0.w, f[4] + e1.x, f[5] + e1.y, f[6] + e1.z, f[7] + e1.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      hi[i] = pack_bf16x2(g[2 * i], g[2 * i + 1]);
      lo[i] = pack_bf16x2(g[2 * i] - __uint_as_float(hi[i] << 16), g[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
    }
    __nv_bfloat16* dst = p.a_xpe + static_cast<long long>(bq) * (2 * p.C) + ch;
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + p.C) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

}  // namespace parq
