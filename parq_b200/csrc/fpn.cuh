// "Next" row f-3 of SURVEY.md 8: the FPN upsample + concat that builds `all_features`
//   ResnetFPN.forward (reference model/resnet_fpn.py:73-80):
//     for layer in 0..3: F.interpolate(features[layer], size of features[LAYER], mode="bilinear")  -> torch.cat(dim=1)
// One pass instead of four interpolate kernels and a concat copy: every output element (image, level*Cl + c, y, x)
// is produced directly from its pyramid level with ATen's upsample_bilinear2d arithmetic (align_corners=False:
// src = max(scale*(dst+0.5)-0.5, 0), scale = in/out in fp32; the target level itself is an exact copy).
#pragma once
#include "ptx.cuh"

namespace parq {

struct FpnParams {
  const void* level[4];    // (BT, Cl, h[l], w[l]) channels-first, fp32 or (in_bf16) bf16
  int in_bf16;
  int h[4], w[4];
  int BT, Cl, H, W;        // output (BT, 4*Cl, H, W)
  int plane0;              // first output plane (bt * 4*Cl + c) of this launch
  void* out;               // fp32, or bf16 (TOut): the channels-first addend of the AddRayPE producer at half the bytes
};

// grid = (pixel tiles, BT * 4 * Cl planes): a block works inside ONE output plane, so the level / channel / image
// decomposition is per block and the only per-thread division is pixel -> (y, x).
template <typename TIn>
__device__ __forceinline__ float fpn_ld(const TIn* p);
template <>
__device__ __forceinline__ float fpn_ld<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float fpn_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void fpn_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void fpn_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
fpn_concat_kernel(const FpnParams p) {
  const int plane = p.plane0 + blockIdx.y;            // bt * 4*Cl + c
  const int c = plane % (4 * p.Cl), bt = plane / (4 * p.Cl);
  const int l = c / p.Cl, cl = c - l * p.Cl;
  const int h = l == 0 ? p.h[0] : (l == 1 ? p.h[1] : (l == 2 ? p.h[2] : p.h[3]));
  const int w = l == 0 ? p.w[0] : (l == 1 ? p.w[1] : (l == 2 ? p.w[2] : p.w[3]));
  const TIn* __restrict__ src = static_cast<const TIn*>(l == 0 ? p.level[0] : (l == 1 ? p.level[1] : (l == 2 ? p.level[2] : p.level[3]))) +
                                (static_cast<long long>(bt) * p.Cl + cl) * h * w;
  TOut* __restrict__ dst = static_cast<TOut*>(p.out) + static_cast<long long>(plane) * p.H * p.W;
  const int HW = p.H * p.W;
  const bool copy = (h == p.H && w == p.W);
  const float sy = static_cast<float>(h) / static_cast<float>(p.H), sx = static_cast<float>(w) / static_cast<float>(p.W);
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    float v;
    if (copy) {
      v = fpn_ld<TIn>(src + pix);
    } else {
      const int y = pix / p.W, x = pix - y * p.W;
      const float fy = fmaxf(__fadd_rn(__fmul_rn(sy, static_cast<float>(y) + 0.5f), -0.5f), 0.f);
      const float fx = fmaxf(__fadd_rn(__fmul_rn(sx, static_cast<float>(x) + 0.5f), -0.5f), 0.f);
      const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float ly1 = fy - static_cast<float>(y0), lx1 = fx - static_cast<float>(x0);
      const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      const float v00 = fpn_ld<TIn>(src + y0 * w + x0), v01 = fpn_ld<TIn>(src + y0 * w + x1), v10 = fpn_ld<TIn>(src + y1 * w + x0), v11 = fpn_ld<TIn>(src + y1 * w + x1);
      // ATen: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
      v = __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01))),
                    __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11))));
    }
    fpn_st(dst + pix, v);
  }
}

}  // namespace parq
