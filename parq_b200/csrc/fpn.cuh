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
  const float* level[4];   // (BT, Cl, h[l], w[l]) fp32, channels-first
  int h[4], w[4];
  int BT, Cl, H, W;        // output (BT, 4*Cl, H, W)
  float* out;
};

__global__ void __launch_bounds__(256)
fpn_concat_kernel(const FpnParams p) {
  const long long total = static_cast<long long>(p.BT) * 4 * p.Cl * p.H * p.W;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % p.W);
    const int y = static_cast<int>((i / p.W) % p.H);
    const int c = static_cast<int>((i / (static_cast<long long>(p.W) * p.H)) % (4 * p.Cl));
    const int bt = static_cast<int>(i / (static_cast<long long>(p.W) * p.H * 4 * p.Cl));
    const int l = c / p.Cl, cl = c - l * p.Cl;
    const int h = l == 0 ? p.h[0] : (l == 1 ? p.h[1] : (l == 2 ? p.h[2] : p.h[3]));
    const int w = l == 0 ? p.w[0] : (l == 1 ? p.w[1] : (l == 2 ? p.w[2] : p.w[3]));
    const float* src = (l == 0 ? p.level[0] : (l == 1 ? p.level[1] : (l == 2 ? p.level[2] : p.level[3]))) +
                       (static_cast<long long>(bt) * p.Cl + cl) * h * w;
    float v;
    if (h == p.H && w == p.W) {
      v = src[y * w + x];
    } else {
      const float sy = static_cast<float>(h) / static_cast<float>(p.H), sx = static_cast<float>(w) / static_cast<float>(p.W);
      const float fy = fmaxf(__fadd_rn(__fmul_rn(sy, static_cast<float>(y) + 0.5f), -0.5f), 0.f);
      const float fx = fmaxf(__fadd_rn(__fmul_rn(sx, static_cast<float>(x) + 0.5f), -0.5f), 0.f);
      const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float ly1 = fy - static_cast<float>(y0), lx1 = fx - static_cast<float>(x0);
      const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      const float v00 = src[y0 * w + x0], v01 = src[y0 * w + x1], v10 = src[y1 * w + x0], v11 = src[y1 * w + x1];
      // ATen: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
      v = __fadd_rn(__fmul_rn(ly0, __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01))),
                    __fmul_rn(ly1, __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11))));
    }
    p.out[i] = v;
  }
}

}  // namespace parq
