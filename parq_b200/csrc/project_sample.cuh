// K0  pose chain   T_camera_local = T_camera_pseudoCam o (T_world_pseudoCam^-1 o T_world_local)
//                  (reference transformer_parq.py:298-300, utils/wrappers.py:247-257)
// K1  project_sample: for every query reference point and every view
//       local -> camera transform (wrappers.py:260-267), pinhole projection and validity
//       (wrappers.py:502-522), normalised-grid round trip and zero-padded bilinear gather of the
//       C-channel texel (transformer_parq.py:148-152, ATen grid_sampler_2d), sum over ALL views
//       divided by the number of VALID views (transformer_parq.py:156-160),
//     fused with the bf16 hi/lo split the following GEMMs consume (the sampled features leave the kernel
//     ONCE, as that split: 4 bytes per element, the size of the fp32 tensor the reference produces).
//
// Rounding contract (bit-exact center_im / center_valid against the CPU oracle): every operation
// below that feeds center_im or the validity test is an explicitly rounded IEEE fp32 intrinsic in
// the order the reference executes it -- 3-term dot products of the tiny pose matmuls as
// ((a0*b0 + a1*b1) + a2*b2) without FMA, the point transform as fma(p2,r2,fma(p1,r1,p0*r0)) + t,
// then x/z (IEEE divide), *f, +c.  Never compile this file with --use_fast_math.
//
// Memory behaviour: the gather is HBM bound and the kernel is built around the TMA engine.  Persistent CTAs (two
// per SM) each own a contiguous range of queries.  One THREAD per (query, view) pair projects once into a
// shared-memory tap table (the IEEE-exact projection is instruction-heavy).  Then one producer thread walks the
// table and, for every view with an in-bounds corner, issues bulk copies (cp.async.bulk, global -> shared,
// completion on an mbarrier) of the bilinear footprint -- the two horizontally adjacent texels of a row are ONE
// contiguous 2*C*2-byte piece (4 KB at C = 1024) -- into a ring of 8 KB slots; views without an in-bounds corner
// cost nothing.  Four consumer warps (256 channels each, 16-byte shared-memory reads per lane) weight and
// accumulate the slots in view order and write a query's row when its last view has been consumed.  The bytes in
// flight are set by the ring (2 x 12 x 8 KB per SM), not by registers or occupancy.
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
  const __nv_bfloat16* tokens;     // (B, T, H, W, C) channels-last bf16
  const __nv_bfloat16* tokens_lo;  // optional second plane: tokens = tokens + tokens_lo (fp32 tokens as an exact bf16 pair)
  const float* ref;                // (B, Nq, 3) normalised reference points in (0,1)
  const float* T_cl;               // (B, T, 12)
  const float* camera;             // (B, T, 6) [w,h,fx,fy,cx,cy]
  float* feat;                     // (B*Nq, C) fp32 sampled features, or nullptr
  __nv_bfloat16* a_x;              // (B*Nq, 2C) [hi|lo] split of the features, or nullptr
  float* center_im;                // (B, T, Nq, 2) or nullptr
  uint8_t* valid;                  // (B, T, Nq)    or nullptr
  float* coord_pos;                // (B, Nq, 3)    or nullptr
  int B, T, H, W, C, Nq;
  int rows_per_cta;                // queries per CTA (contiguous range)
  int slots;                       // ring slots of 4*C*2 bytes
  float span[3], lo[3];            // denormalisation: p*span + lo
};

struct ViewTap {                 // one (query, view) bilinear footprint
  int x0, y0;                    // floor of the sample position (may be -1)
  float fx, fy;                  // fractional parts
  int inb;                       // bit0 nw, bit1 ne, bit2 sw, bit3 se in bounds
};

__device__ __forceinline__ void fma_bf16x8(float (&acc)[8], const uint4& v, float w) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc[2 * i] = fmaf(__uint_as_float(u[i] << 16), w, acc[2 * i]);
    acc[2 * i + 1] = fmaf(__uint_as_float(u[i] & 0xFFFF0000u), w, acc[2 * i + 1]);
  }
}

// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

namespace sample {
constexpr int THREADS = 160;                   // warp 0: producer; warps 1..4: consumers (256 channels each at C = 1024)
constexpr int CONSUMERS = 4;
constexpr int MAX_PAIRS = 256;                 // (query, view) pairs per group of a CTA's range (tap table)
constexpr int MAX_SLOTS = 16;
constexpr int CTAS_PER_SM = 2;
constexpr int LO_PLANE = 1 << 30;              // slot meta: the slot holds the low-order token plane of the pair
__host__ __device__ inline size_t smem_bytes(int C, int slots) {
  return 1024 /*align slack*/ + static_cast<size_t>(slots) * 4 * C * 2 + MAX_PAIRS * sizeof(ViewTap) + MAX_PAIRS * sizeof(int) /*nvalid*/ +
         2 * MAX_SLOTS * sizeof(uint64_t) + MAX_SLOTS * sizeof(int);
}
}  // namespace sample

// grid = ceil(R / rows_per_cta) persistent CTAs of 5 warps.  Requires C == 8 * 32 * CONSUMERS = 1024 (the reference width,
// enforced by check_shape) and T <= MAX_PAIRS.
__global__ void __launch_bounds__(sample::THREADS, sample::CTAS_PER_SM)
project_sample_kernel(const SampleParams p) {
  using namespace sample;
  extern __shared__ uint8_t smem_raw_s[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_s) + 1023) & ~uintptr_t(1023));
  const int slot_bytes = 4 * p.C * 2;
  uint8_t* ring = smem;
  ViewTap* s_tap = reinterpret_cast<ViewTap*>(ring + static_cast<size_t>(p.slots) * slot_bytes);
  int* s_nvalid = reinterpret_cast<int*>(s_tap + MAX_PAIRS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_nvalid + MAX_PAIRS);
  uint64_t* empty_bar = full_bar + MAX_SLOTS;
  int* s_meta = reinterpret_cast<int*>(empty_bar + MAX_SLOTS);

  const int R = p.B * p.Nq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_begin = blockIdx.x * p.rows_per_cta;
  const int r_end = min(R, r_begin + p.rows_per_cta);
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.slots; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CONSUMERS);
    }
    fence_mbar_init();
  }
  // the reference points of iteration i > 0 are written by the previous kernel (heads): nothing is read before the wait
  pdl_wait();
  pdl_launch_dependents();

  const float Wm1 = static_cast<float>(p.W - 1), Hm1 = static_cast<float>(p.H - 1);
  const float sx = Wm1 / 2.f, sy = Hm1 / 2.f;          // ATen CPU grid_sampler: scaling = (size-1)/2
  const long long rowpitch = static_cast<long long>(p.W) * p.C;
  const int rows_per_group = max(1, MAX_PAIRS / p.T);
  int it = 0;                                           // ring cursor (producer thread and consumers advance in lock step)

  for (int g_begin = r_begin; g_begin < r_end; g_begin += rows_per_group) {
    const int g_rows = min(rows_per_group, r_end - g_begin);
    const int npairs = g_rows * p.T;
    __syncthreads();                                    // the previous group's table is no longer in use
    for (int i = threadIdx.x; i < g_rows; i += blockDim.x) s_nvalid[i] = 0;
    __syncthreads();
    // ---------------------------------------------------------------- phase 1: projection
    for (int pair = threadIdx.x; pair < npairs; pair += blockDim.x) {
      const int rl = pair / p.T, t = pair - rl * p.T;
      const int row = g_begin + rl;
      const int b = row / p.Nq;
      // reference point -> metres in the local frame: p*span + lo (separately rounded mul, add)
      const float px = __fadd_rn(__fmul_rn(p.ref[row * 3 + 0], p.span[0]), p.lo[0]);
      const float py = __fadd_rn(__fmul_rn(p.ref[row * 3 + 1], p.span[1]), p.lo[1]);
      const float pz = __fadd_rn(__fmul_rn(p.ref[row * 3 + 2], p.span[2]), p.lo[2]);
      if (t == 0 && p.coord_pos != nullptr) {
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
      s_tap[pair] = tap;
      if (is_valid) atomicAdd(&s_nvalid[rl], 1);         // integer count: order independent
      const long long oc = (static_cast<long long>(b) * p.T + t) * p.Nq + (row - b * p.Nq);
      if (p.center_im != nullptr) { p.center_im[oc * 2] = u; p.center_im[oc * 2 + 1] = v; }
      if (p.valid != nullptr) p.valid[oc] = static_cast<uint8_t>(is_valid);
    }
    __syncthreads();

    if (warp == 0) {
      // ------------------------------------------------------------ producer: one thread feeds the ring
      if (lane == 0) {
        const int nplanes = p.tokens_lo != nullptr ? 2 : 1;
        for (int pair = 0; pair < npairs; ++pair) {
          const ViewTap tp = s_tap[pair];
          if (tp.inb == 0) continue;
          const int rl = pair / p.T, t = pair - rl * p.T;
          const long long view = static_cast<long long>((g_begin + rl) / p.Nq) * p.T + t;
          const long long texel = ((view * p.H + tp.y0) * p.W + tp.x0) * static_cast<long long>(p.C);   // nw corner (may be out of range)
          const uint32_t cb = static_cast<uint32_t>(p.C) * 2;                                           // bytes per texel
          uint32_t bytes = 0;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int m = (tp.inb >> (2 * r)) & 3;
            bytes += (m == 3) ? 2 * cb : (m != 0 ? cb : 0u);
          }
          for (int plane = 0; plane < nplanes; ++plane, ++it) {
            const int slot = it % p.slots;
            mbar_wait(&empty_bar[slot], ((it / p.slots) & 1) ^ 1);
            s_meta[slot] = pair | (plane ? LO_PLANE : 0);
            mbar_expect_tx(&full_bar[slot], bytes);
            const __nv_bfloat16* src = (plane ? p.tokens_lo : p.tokens) + texel;
            uint8_t* dst = ring + static_cast<size_t>(slot) * slot_bytes;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int m = (tp.inb >> (2 * r)) & 3;
              const __nv_bfloat16* rs = src + r * rowpitch;
              uint8_t* rd = dst + r * 2 * cb;
              if (m == 3) bulk_load(rd, rs, 2 * cb, &full_bar[slot]);               // west | east: one contiguous piece
              else if (m == 1) bulk_load(rd, rs, cb, &full_bar[slot]);              // west only (x0 = W-1)
              else if (m == 2) bulk_load(rd + cb, rs + p.C, cb, &full_bar[slot]);   // east only (x0 = -1)
            }
          }
        }
        // end of group: an empty slot carrying the sentinel
        const int slot = it % p.slots;
        mbar_wait(&empty_bar[slot], ((it / p.slots) & 1) ^ 1);
        s_meta[slot] = -1;
        mbar_arrive(&full_bar[slot]);
        ++it;
      }
    } else {
      // ------------------------------------------------------------ consumers: warp w owns channels [256(w-1), 256w)
      const int ch = (warp - 1) * 256 + lane * 8;
      int cur = -1;                                                 // group-local row being accumulated
      float tot[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) tot[i] = 0.f;
      // rows are finished in order: flush() writes row `rl` from `tot`, zero_rows() the rows no view touched
      auto write_row = [&](int rl, const float (&f)[8]) {
        const int row = g_begin + rl;
        if (p.feat != nullptr) {
          float* o = p.feat + static_cast<long long>(row) * p.C + ch;
          reinterpret_cast<float4*>(o)[0] = make_float4(f[0], f[1], f[2], f[3]);
          reinterpret_cast<float4*>(o)[1] = make_float4(f[4], f[5], f[6], f[7]);
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
      };
      auto flush = [&](int rl) {
        // sum over ALL views / max(#valid views, 1): x / n through one reciprocal and a Newton correction
        // (q = x*r; q += (x - n*q)*r), which reproduces the correctly rounded quotient of the reference's true division
        const int nv = s_nvalid[rl];
        const float cnt = static_cast<float>(max(nv, 1));
        const float rc = __frcp_rn(cnt);
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float q0 = tot[i] * rc;
          f[i] = nv <= 1 ? tot[i] : __fmaf_rn(__fmaf_rn(-cnt, q0, tot[i]), rc, q0);
          tot[i] = 0.f;
        }
        write_row(rl, f);
      };
      auto zero_rows = [&](int from, int to) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int rl = from; rl < to; ++rl) write_row(rl, z);
      };
      for (;; ++it) {
        const int slot = it % p.slots;
        mbar_wait(&full_bar[slot], (it / p.slots) & 1);
        const int meta = s_meta[slot];
        if (meta < 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[slot]);
          ++it;
          break;
        }
        const int pair = meta & (LO_PLANE - 1);
        const int rl = pair / p.T;
        if (rl != cur) {
          if (cur >= 0) flush(cur);
          zero_rows(cur + 1, rl);
          cur = rl;
        }
        const ViewTap tp = s_tap[pair];                              // shared-memory broadcast
        const float ex = 1.f - tp.fx, sy_ = 1.f - tp.fy;             // distances to east / south (ATen CPU form)
        const float wgt[4] = {sy_ * ex, sy_ * tp.fx, tp.fy * ex, tp.fy * tp.fx};    // nw, ne, sw, se
        const uint8_t* sp = ring + static_cast<size_t>(slot) * slot_bytes + ch * 2;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)                                  // corner order nw, ne, sw, se as in ATen
          if (tp.inb & (1 << c)) fma_bf16x8(acc, *reinterpret_cast<const uint4*>(sp + c * p.C * 2), wgt[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) tot[i] += acc[i];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
      }
      if (cur >= 0) flush(cur);
      zero_rows(cur + 1, g_rows);
    }
  }
}

}  // namespace parq
