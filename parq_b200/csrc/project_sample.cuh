// K0  pose chain   T_camera_local = T_camera_pseudoCam o (T_world_pseudoCam^-1 o T_world_local)
//                  (reference transformer_parq.py:298-300, utils/wrappers.py:247-257)
// K1  project_sample: for every query reference point and every view
//       local -> camera transform (wrappers.py:260-267), pinhole projection and validity
//       (wrappers.py:502-522), normalised-grid round trip and zero-padded bilinear gather of the
//       C-channel texel (transformer_parq.py:148-152, ATen grid_sampler_2d), sum over ALL views
//       divided by the number of VALID views (transformer_parq.py:156-160),
//     fused with the bf16 hi/lo split the following GEMMs consume: the sampled features leave the kernel ONCE, as
//     that split (2 + 2 bytes per element -- the size of the fp32 tensor the reference produces); "+ query positional
//     feature" (transformer_parq.py:372) is added by the epilogue of the GEMM that produces that feature.
//
// Rounding contract (bit-exact center_im / center_valid against the CPU oracle): every operation
// below that feeds center_im or the validity test is an explicitly rounded IEEE fp32 intrinsic in
// the order the reference executes it -- 3-term dot products of the tiny pose matmuls as
// ((a0*b0 + a1*b1) + a2*b2) without FMA, the point transform as fma(p2,r2,fma(p1,r1,p0*r0)) + t,
// then x/z (IEEE divide), *f, +c.  Never compile this file with --use_fast_math.
//
// Memory behaviour: the gather is HBM/L2 bound.  A block handles four queries: one THREAD per (query, view)
// pair projects once into shared memory (the IEEE-exact projection is instruction-heavy), then warp w owns
// channels [256w, 256w+256) of every query; every lane moves 16-byte vectors (8 bf16 channels), so each
// bilinear corner is a fully coalesced 512-byte request and the two horizontally adjacent corners form one
// contiguous 2*C*2-byte segment.  Views without an in-bounds corner cost nothing (compacted by a ballot), two
// views (8 independent 16-byte loads per lane) are in flight, and the first loads of the next query are issued
// before the epilogue of the current one.
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
  pdl_wait();
  pdl_launch_dependents();
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
  const __nv_bfloat16* tokens_lo;  // optional second plane: token value = tokens + tokens_lo (fp32 tokens as an exact bf16 pair)
  const float* ref;              // (B, Nq, 3) normalised reference points in (0,1)
  const float* T_cl;             // (B, T, 12)
  const float* camera;           // (B, T, 6) [w,h,fx,fy,cx,cy]
  float* feat;                   // (B*Nq, C) fp32 sampled features, or nullptr
  __nv_bfloat16* a_x;            // (B*Nq, 2C) [hi|lo] split of the features, or nullptr
  float* center_im;              // (B, T, Nq, 2) or nullptr
  uint8_t* valid;                // (B, T, Nq)    or nullptr
  float* coord_pos;              // (B, Nq, 3)    or nullptr
  int B, T, H, W, C, Nq;
  int ref_is_fresh;              // the reference points were written by the IMMEDIATELY preceding kernel: no projection before the PDL wait
  float span[3], lo[3];          // denormalisation: p*span + lo
};

struct ViewTap {                 // one (query, view) bilinear footprint
  int x0, y0;                    // floor of the sample position (may be out of range)
  float fx, fy;                  // fractional parts
  int inb;                       // bit0 nw, bit1 ne, bit2 sw, bit3 se in bounds; bit4: the view is VALID for the average
  float u, v;                    // projected pixel coordinates (center_im), parked here until they may be written
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

constexpr int SAMPLE_QPB = 4;                  // queries per block (Nq % 4 == 0, so a block never straddles clips)
constexpr int SAMPLE_LOADS_IN_FLIGHT = 8;      // independent 16-byte loads per lane and round: 2 views, or 1 view with its low-order plane

// grid = B*Nq/4 blocks of 4 warps.
// Phase 1: one THREAD per (query, view) pair of the block's 4 queries projects the reference point and leaves the
//          bilinear footprint in shared memory (the IEEE-exact projection is instruction-heavy: doing it once per
//          pair instead of once per warp is what makes the kernel memory- rather than issue-bound).
// Phase 2: warp w owns channels [256w, 256w+256) of every query: per query the views with at least one in-bounds
//          corner are compacted from a ballot and gathered with 16-byte loads, two views (8 loads) in flight.
// Programmatic dependent launch: reference points, poses and tokens are inputs or were produced at least two
// launches earlier (the launch order is heads -> posemb -> sampling), so phase 1 and the first round of texel loads
// are issued BEFORE the dependency wait (they overlap the tail of the previous kernel); all outputs are written after it.
// kLo: the tokens come as an exact bf16 pair (hi, lo): one view (4 + 4 loads) per round instead of two.
template <bool kLo>
__global__ void __launch_bounds__(128, 7)
project_sample_kernel(const SampleParams p) {
  constexpr int V = kLo ? 1 : 2;
  constexpr int NP = kLo ? 2 : 1;                // token planes
  extern __shared__ ViewTap s_tap[];             // [SAMPLE_QPB][T]
  const int row0 = blockIdx.x * SAMPLE_QPB;      // first (b*Nq + q) row of this block
  const int b = row0 / p.Nq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = warp * 256 + lane * 8;
  const float Wm1 = static_cast<float>(p.W - 1), Hm1 = static_cast<float>(p.H - 1);
  const float sx = Wm1 / 2.f, sy = Hm1 / 2.f;          // ATen CPU grid_sampler: scaling = (size-1)/2
  const int npairs = SAMPLE_QPB * p.T;
  const bool early = p.ref_is_fresh == 0 && npairs <= static_cast<int>(blockDim.x);   // one pair per thread: its outputs can wait in registers
  if (!early) {
    pdl_wait();
    pdl_launch_dependents();
  }

  // ---------------------------------------------------------------- phase 1: projection
  for (int pair = threadIdx.x; pair < npairs; pair += blockDim.x) {
    const int qi = pair / p.T, t = pair - qi * p.T;
    const int row = row0 + qi;
    // reference point -> metres in the local frame: p*span + lo (separately rounded mul, add)
    const float px = __fadd_rn(__fmul_rn(p.ref[row * 3 + 0], p.span[0]), p.lo[0]);
    const float py = __fadd_rn(__fmul_rn(p.ref[row * 3 + 1], p.span[1]), p.lo[1]);
    const float pz = __fadd_rn(__fmul_rn(p.ref[row * 3 + 2], p.span[2]), p.lo[2]);
    if (t == 0 && p.coord_pos != nullptr && !early) {
      p.coord_pos[row * 3 + 0] = px; p.coord_pos[row * 3 + 1] = py; p.coord_pos[row * 3 + 2] = pz;
    }
    // pose (12 floats, 48-byte rows) and camera (6 floats, 24-byte rows) as vector loads: 6 requests instead of 18
    const float4* Tc4 = reinterpret_cast<const float4*>(p.T_cl + (static_cast<long long>(b) * p.T + t) * 12);
    const float2* cam2 = reinterpret_cast<const float2*>(p.camera + (static_cast<long long>(b) * p.T + t) * 6);
    const float4 t0 = __ldg(Tc4), t1 = __ldg(Tc4 + 1), t2 = __ldg(Tc4 + 2);
    const float2 c0 = __ldg(cam2), c1 = __ldg(cam2 + 1), c2 = __ldg(cam2 + 2);
    const float Tc[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
    const float cam[6] = {c0.x, c0.y, c1.x, c1.y, c2.x, c2.y};
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
    const int is_valid = in_front && (u >= 0.f) && (u <= wm1) && (v >= 0.f) && (v <= hm1);
    // normalised grid and back (transformer_parq.py:148-150 then grid_sampler unnormalize)
    const float gx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, u), Wm1), -1.f);
    const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, v), Hm1), -1.f);
    const float ix = __fmul_rn(__fadd_rn(gx, 1.f), sx);
    const float iy = __fmul_rn(__fadd_rn(gy, 1.f), sy);
    const float x0f = floorf(ix), y0f = floorf(iy);
    ViewTap tap;
    tap.fx = ix - x0f;
    tap.fy = iy - y0f;
    // in-bounds tests in float (ix may be huge or NaN), then a safe int conversion
    const bool xw = (x0f >= 0.f) && (x0f <= Wm1), xe = (x0f >= -1.f) && (x0f <= Wm1 - 1.f);
    const bool yn = (y0f >= 0.f) && (y0f <= Hm1), ys = (y0f >= -1.f) && (y0f <= Hm1 - 1.f);
    tap.inb = (xw && yn ? 1 : 0) | (xe && yn ? 2 : 0) | (xw && ys ? 4 : 0) | (xe && ys ? 8 : 0);
    tap.x0 = tap.inb ? static_cast<int>(x0f) : 0;
    tap.y0 = tap.inb ? static_cast<int>(y0f) : 0;
    tap.inb |= is_valid ? 16 : 0;
    tap.u = u;
    tap.v = v;
    s_tap[pair] = tap;
    if (!early) {
      const long long oc = (static_cast<long long>(b) * p.T + t) * p.Nq + (row - b * p.Nq);
      if (p.center_im != nullptr) { p.center_im[oc * 2] = u; p.center_im[oc * 2 + 1] = v; }
      if (p.valid != nullptr) p.valid[oc] = static_cast<uint8_t>(is_valid);
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2: gather
  // All control flow below is warp-uniform (`live` is a ballot, taps are shared-memory broadcasts): absent views
  // and out-of-bounds corners cost no instructions.  The first round of loads of query j+1 is issued before the
  // epilogue of query j, so a warp always has texel loads in flight.
  const long long rowpitch = static_cast<long long>(p.W) * p.C;
  const ViewTap* taps = s_tap;     // taps of the query whose views are being issued
  int tb = 0, nvalid = 0;          // 32-view chunk cursor and valid-view count of that query
  unsigned live = 0u;              // views of the chunk with at least one in-bounds corner, not yet issued
  uint4 tex[V][NP][4];
  float wgt[V][4];
  int inb[V];
  auto load_chunk = [&]() __attribute__((always_inline)) {
    const int inb_l = (tb + lane < p.T) ? taps[tb + lane].inb : 0;
    nvalid += __popc(__ballot_sync(0xffffffffu, inb_l & 16));
    live = __ballot_sync(0xffffffffu, (inb_l & 15) != 0);
  };
  auto begin_query = [&](int qi) __attribute__((always_inline)) {
    taps = s_tap + qi * p.T;
    tb = 0;
    nvalid = 0;
    load_chunk();
    while (live == 0u && tb + 32 < p.T) { tb += 32; load_chunk(); }
  };
  auto issue_round = [&]() __attribute__((always_inline)) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      inb[k] = 0;
      if (live != 0u) {
        const int src = __ffs(live) - 1;
        live &= live - 1;
        const ViewTap tp = taps[tb + src];               // shared-memory broadcast
        inb[k] = tp.inb & 15;
        const float ex = 1.f - tp.fx, sy_ = 1.f - tp.fy;   // distances to east / south (ATen CPU form)
        wgt[k][0] = sy_ * ex;      // nw
        wgt[k][1] = sy_ * tp.fx;   // ne
        wgt[k][2] = tp.fy * ex;    // sw
        wgt[k][3] = tp.fy * tp.fx; // se
        const long long view = static_cast<long long>(b) * p.T + tb + src;
        const long long toff = ((view * p.H + tp.y0) * p.W + tp.x0) * static_cast<long long>(p.C) + ch;
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
          const __nv_bfloat16* base = (pl == 0 ? p.tokens : p.tokens_lo) + toff;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (inb[k] & (1 << c)) tex[k][pl][c] = ldg_nc_16(base + (c >> 1) * rowpitch + (c & 1) * p.C);
        }
        while (live == 0u && tb + 32 < p.T) { tb += 32; load_chunk(); }
      }
    }
  };

  begin_query(0);
  issue_round();
  if (early) {
    pdl_wait();
    pdl_launch_dependents();
    if (static_cast<int>(threadIdx.x) < npairs && (p.center_im != nullptr || p.valid != nullptr)) {
      const int pqi = threadIdx.x / p.T, pt = threadIdx.x - pqi * p.T;      // pair index == thread index in phase 1
      const long long oc = (static_cast<long long>(b) * p.T + pt) * p.Nq + (row0 + pqi - b * p.Nq);
      const ViewTap tp = s_tap[threadIdx.x];
      if (p.center_im != nullptr) { p.center_im[oc * 2] = tp.u; p.center_im[oc * 2 + 1] = tp.v; }
      if (p.valid != nullptr) p.valid[oc] = static_cast<uint8_t>((tp.inb >> 4) & 1);
    }
    if (p.coord_pos != nullptr && threadIdx.x < SAMPLE_QPB * 3) {
      const int r = row0 + threadIdx.x / 3, a = threadIdx.x % 3;
      // static selects: a dynamically indexed parameter array would push the whole parameter block into local memory
      const float sp_a = a == 0 ? p.span[0] : (a == 1 ? p.span[1] : p.span[2]);
      const float lo_a = a == 0 ? p.lo[0] : (a == 1 ? p.lo[1] : p.lo[2]);
      p.coord_pos[r * 3 + a] = __fadd_rn(__fmul_rn(p.ref[r * 3 + a], sp_a), lo_a);
    }
  }
#pragma unroll 1
  for (int qi = 0; qi < SAMPLE_QPB; ++qi) {
    const int row = row0 + qi;
    const long long o = static_cast<long long>(row) * p.C + ch;
    float tot[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) tot[i] = 0.f;
    while (inb[0] != 0) {                                    // the first slot is empty only when no view is left
#pragma unroll
      for (int k = 0; k < V; ++k) {
        if (inb[k] != 0) {
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {
            if (pl == 1) {                                   // the low-order plane is a second, separately accumulated view
#pragma unroll
              for (int i = 0; i < 8; ++i) { tot[i] += acc[i]; acc[i] = 0.f; }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)                      // corner order nw, ne, sw, se as in ATen
              if (inb[k] & (1 << c)) fma_bf16x8(acc, tex[k][pl][c], wgt[k][c]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) tot[i] += acc[i];
        }
      }
      issue_round();
    }
    const int nv = nvalid;                                   // every chunk of this query has been visited
    if (qi + 1 < SAMPLE_QPB) {
      begin_query(qi + 1);
      issue_round();
    }

    // sum over ALL views / max(#valid views, 1): x / n through one reciprocal and a Newton correction
    // (q = x*r; q += (x - n*q)*r), which reproduces the correctly rounded quotient of the reference's true division
    const float cnt = static_cast<float>(max(nv, 1));
    const float rc = __frcp_rn(cnt);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float q0 = tot[i] * rc;
      f[i] = nv <= 1 ? tot[i] : __fmaf_rn(__fmaf_rn(-cnt, q0, tot[i]), rc, q0);
    }
    if (p.feat != nullptr) {
      reinterpret_cast<float4*>(p.feat + o)[0] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(p.feat + o)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    if (p.a_x != nullptr) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        hi[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
        lo[i] = pack_bf16x2(f[2 * i] - __uint_as_float(hi[i] << 16), f[2 * i + 1] - __uint_as_float(hi[i] & 0xFFFF0000u));
      }
      __nv_bfloat16* dst = p.a_x + static_cast<long long>(row) * (2 * p.C) + ch;
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + p.C) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

}  // namespace parq
