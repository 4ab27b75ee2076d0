// "Next" row f-1 of SURVEY.md 8: the ray positional encoding + tokeniser that feeds the decoder
//   AddRayPE.forward                  (reference model/ray_positional_encoding.py:61-139)
//   grid_2d / ray_points_snippet / ray_points (utils/encoding_utils.py:15-100), Camera.unproject (utils/wrappers.py:523-549)
//   tokens = rearrange(features + encoding, "b t c h w -> b (t h w) c")   (model/parq_lightning.py:75-85)
// Two small kernels build the 192 encoder inputs per pixel as the [hi|lo] bf16 split the tensor-core GEMM consumes;
// the two encoder layers are gemm_tc_kernel launches whose epilogue adds the channels-first backbone features and
// writes channels-last bf16 tokens directly (no fp32 encoding tensor, no transpose copy, no cast pass).
#pragma once
#include "project_sample.cuh"

namespace parq {

// per (clip, view): p_local = A (ray * depth) + t with
//   A = R_lp R_pc, t = t_lp + R_lp t_pc,  T_pc = T_camera_pseudoCam^-1,  T_lp = T_world_local^-1 o T_world_pseudoCam
__global__ void raype_affine_kernel(const float* __restrict__ T_cp, const float* __restrict__ T_wp, const float* __restrict__ T_wl,
                                    float* __restrict__ aff, int B, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int b = i / T;
  float cp[12], wp[12], wl[12], pc[12], lw[12], lp[12], out[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    cp[k] = T_cp[i * 12 + k];
    wp[k] = T_wp[i * 12 + k];
    wl[k] = T_wl[b * 12 + k];
  }
  pose_inverse(cp, pc);
  pose_inverse(wl, lw);
  pose_compose(lw, wp, lp);
  pose_compose(lp, pc, out);
#pragma unroll
  for (int k = 0; k < 12; ++k) aff[i * 12 + k] = out[k];
}

struct RayFeatParams {
  const float* camera;   // (B*T, 6)
  const float* aff;      // (B*T, 12) from raype_affine_kernel
  const float* depth;    // (n) depth planes
  __nv_bfloat16* out;    // (B*T*H*W, 2*3n): [hi | lo] of the 3n encoder inputs, sample-major (n c)
  int BT, H, W, n;
  float lo[3], span[3];  // ray_points_scale: (p - lo) / span
};

// thread = (pixel, depth sample)
__global__ void __launch_bounds__(256)
ray_features_kernel(const RayFeatParams p) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long ntok = static_cast<long long>(p.BT) * p.H * p.W;
  if (gid >= ntok * p.n) return;
  const long long tok = gid / p.n;
  const int i = static_cast<int>(gid - tok * p.n);
  const int HW = p.H * p.W;
  const int bt = static_cast<int>(tok / HW), pix = static_cast<int>(tok - static_cast<long long>(bt) * HW);
  const int y = pix / p.W, x = pix - y * p.W;
  const float* cam = p.camera + bt * 6;
  const float* A = p.aff + bt * 12;
  // unproject: ((u - cx) / fx, (v - cy) / fy, 1) with the pixel grid u = x, v = y (grid_2d over [0,W) x [0,H))
  const float rx = (static_cast<float>(x) - cam[4]) / cam[2];
  const float ry = (static_cast<float>(y) - cam[5]) / cam[3];
  const float d = p.depth[i];
  const float pr[3] = {rx * d, ry * d, d};
  const int F = 3 * p.n;
  __nv_bfloat16* o = p.out + tok * (2 * F) + i * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float pl = A[3 * c] * pr[0] + A[3 * c + 1] * pr[1] + A[3 * c + 2] * pr[2] + A[9 + c];
    const float xn = fminf(fmaxf((pl - p.lo[c]) / p.span[c], 0.f), 1.f);
    const float v = logf(fmaxf(xn, 1e-3f) / fmaxf(1.f - xn, 1e-3f));
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    o[c] = h;
    o[F + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace parq
