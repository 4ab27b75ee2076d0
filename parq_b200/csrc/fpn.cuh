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

constexpr int FPN_CH = 8;     // output planes (channels of one level of one image) per block

template <typename TIn>
__device__ __forceinline__ float fpn_ld(const TIn* p);
template <>
__device__ __forceinline__ float fpn_ld<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float fpn_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void fpn_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void fpn_st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TOut>
__device__ __forceinline__ void fpn_st2(TOut* p, float a, float b);
template <>
__device__ __forceinline__ void fpn_st2<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <>
__device__ __forceinline__ void fpn_st2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b); }

// One block = FPN_CH consecutive channels of one pyramid level of one image (Cl % FPN_CH == 0, so a block never straddles a
// level or an image): the bilinear footprint of a pixel (two rows, two columns, two weights each) is computed once and reused
// for the block's channels; a thread owns PAIRS of adjacent pixels (vector stores) when the plane size is even.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
fpn_concat_kernel(const FpnParams p) {
  const int plane0 = (p.plane0 + blockIdx.x) * FPN_CH;      // bt * 4*Cl + c of the block's first plane
  const int c = plane0 % (4 * p.Cl), bt = plane0 / (4 * p.Cl);
  const int l = c / p.Cl, cl = c - l * p.Cl;
  const int h = l == 0 ? p.h[0] : (l == 1 ? p.h[1] : (l == 2 ? p.h[2] : p.h[3]));
  const int w = l == 0 ? p.w[0] : (l == 1 ? p.w[1] : (l == 2 ? p.w[2] : p.w[3]));
  const long long src_plane = static_cast<long long>(h) * w;
  const TIn* __restrict__ src = static_cast<const TIn*>(l == 0 ? p.level[0] : (l == 1 ? p.level[1] : (l == 2 ? p.level[2] : p.level[3]))) +
                                (static_cast<long long>(bt) * p.Cl + cl) * src_plane;
  const int HW = p.H * p.W;
  TOut* __restrict__ dst = static_cast<TOut*>(p.out) + static_cast<long long>(plane0) * HW;
  const bool copy = (h == p.H && w == p.W);
  const float sy = static_cast<float>(h) / static_cast<float>(p.H), sx = static_cast<float>(w) / static_cast<float>(p.W);
  const int step = (HW % 2 == 0) ? 2 : 1;                   // pixels per thread and iteration
  for (int pix = threadIdx.x * step; pix < HW; pix += blockDim.x * step) {
    int o00[2], o01[2], o10[2], o11[2];
    float ly0[2], ly1[2], lx0[2], lx1[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int q = pix + (k < step ? k : 0);
      const int y = q / p.W, x = q - y * p.W;
      // ATen upsample_bilinear2d, align_corners=False: src = max(scale*(dst+0.5)-0.5, 0)
      const float fy = fmaxf(__fadd_rn(__fmul_rn(sy, static_cast<float>(y) + 0.5f), -0.5f), 0.f);
      const float fx = fmaxf(__fadd_rn(__fmul_rn(sx, static_cast<float>(x) + 0.5f), -0.5f), 0.f);
      const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      ly1[k] = fy - static_cast<float>(y0); lx1[k] = fx - static_cast<float>(x0);
      ly0[k] = 1.f - ly1[k]; lx0[k] = 1.f - lx1[k];
      o00[k] = y0 * w + x0; o01[k] = y0 * w + x1; o10[k] = y1 * w + x0; o11[k] = y1 * w + x1;
    }
#pragma unroll
    for (int ch = 0; ch < FPN_CH; ++ch) {
      const TIn* s = src + ch * src_plane;
      float v[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (copy) {
          v[k] = fpn_ld<TIn>(s + pix + (k < step ? k : 0));
        } else {
          const float v00 = fpn_ld<TIn>(s + o00[k]), v01 = fpn_ld<TIn>(s + o01[k]), v10 = fpn_ld<TIn>(s + o10[k]), v11 = fpn_ld<TIn>(s + o11[k]);
          // ATen: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
          v[k] = __fadd_rn(__fmul_rn(ly0[k], __fadd_rn(__fmul_rn(lx0[k], v00), __fmul_rn(lx1[k], v01))),
                           __fmul_rn(ly1[k], __fadd_rn(__fmul_rn(lx0[k], v10), __fmul_rn(lx1[k], v11))));
        }
      }
      TOut* d = dst + static_cast<long long>(ch) * HW + pix;
      if (step == 2) fpn_st2<TOut>(d, v[0], v[1]);
      else fpn_st(d, v[0]);
    }
  }
}

}  // namespace parq
