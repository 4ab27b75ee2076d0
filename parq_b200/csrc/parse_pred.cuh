// parse_pred + NMS on the device ("next" row f-2 of SURVEY.md 8):
//   PARQDecoder.parse_pred          (reference model/parq_decoder.py:372-424)
//   nms / run_nms / nms_3d_faster   (reference utils/nms.py:20-70, 141-179; _samecls :182-224)
//   compute_rotation_matrix_from_ortho6d (utils/ortho6d_transforms.py:53-66)
//   Obb3D.bb3corners_object + Pose.transform (utils/wrappers.py:355-392, 260-267)
// The reference moves the last iteration's boxes to the host, builds Obb3D on the CPU, runs a numpy greedy NMS
// per clip and moves the mask back (2 D2H/H2D round trips per batch).  Here one CTA per clip does all of it:
//   1. thread = box: score/label = max/arg-max of the class probabilities (first index on ties), rotation
//      by Gram-Schmidt, the 8 corners p R^T + t in fp32 with the reference's operation order, their AABB;
//   2. bitonic sort of the non-background boxes by descending score (np.argsort ascending, consumed from the
//      end; ties between exactly equal scores are broken by the higher box index first -- numpy's unstable
//      quicksort leaves that case unspecified);
//   3. thread = sorted position: one row of the "IoU > threshold" bit matrix, IoU in float64 exactly as numpy
//      evaluates it (inter / (area_i + area_j - inter));
//   4. warp 0 walks the sorted order once: a box is kept unless an earlier kept box suppressed it;
//   5. pred_mask = kept & track-scale filter on (x, z) of the centre (parq_decoder.py:408-414).
#pragma once
#include "project_sample.cuh"

namespace parq {

struct ParsePredParams {
  const float* center;     // (B, K, 3)
  const float* size;       // (B, K, 3)
  const float* ortho6d;    // (B, K, 6)
  const float* prob;       // (B, K, num_cls)
  int B, K, num_cls;
  int background;          // label excluded from NMS (num_semcls)
  int same_class;          // 1: suppress only boxes of the same class (nms_3d_faster_samecls)
  int apply_track_scale;   // 0: FOR_VIS, every box is "valid"
  float track_scale[6];
  double threshold;
  uint8_t* pred_mask;      // (B, K)
  uint8_t* nms_mask;       // (B, K) or nullptr
  float* scores;           // (B, K) or nullptr
  int* labels;             // (B, K) or nullptr
  float* obbs;             // (B, K, 19) or nullptr: [xmin,xmax,ymin,ymax,zmin,zmax | R row-major | t | sem_id]
};

// dynamic shared memory: P = power of two >= K
//   double aabb[P][6] | double area[P] | uint32 key[P] | int idx[P] | int label[P] | uint32 sup[P][P/32] | uint32 keep[P/32]
__host__ __device__ inline size_t parse_pred_smem(int P) {
  return static_cast<size_t>(P) * (6 * 8 + 8 + 4 + 4 + 4) + static_cast<size_t>(P) * (P / 32) * 4 + (P / 32) * 4 + 64;
}

__global__ void __launch_bounds__(1024)
parse_pred_kernel(const ParsePredParams p, int P) {
  extern __shared__ double smem_d[];
  double* s_aabb = smem_d;                                  // [P][6]
  double* s_area = s_aabb + 6 * P;                          // [P]
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_area + P);
  int* s_idx = reinterpret_cast<int*>(s_key + P);
  int* s_label = s_idx + P;
  uint32_t* s_sup = reinterpret_cast<uint32_t*>(s_label + P);   // [P][P/32]
  uint32_t* s_keep = s_sup + static_cast<size_t>(P) * (P / 32);
  const int W = P / 32;
  const int b = blockIdx.x;
  const int k = threadIdx.x;
  pdl_wait();
  pdl_launch_dependents();

  // ---- 1. per-box quantities
  uint32_t key = 0u;
  int in_scope = 0;
  if (k < p.K) {
    const long long r = static_cast<long long>(b) * p.K + k;
    const float* pr = p.prob + r * p.num_cls;
    float best = pr[0];
    int lab = 0;
    for (int j = 1; j < p.num_cls; ++j)
      if (pr[j] > best) { best = pr[j]; lab = j; }
    const float* o6 = p.ortho6d + r * 6;
    const float a0 = o6[0], a1 = o6[1], a2 = o6[2], b0 = o6[3], b1 = o6[4], b2 = o6[5];
    // normalize_vector: v / max(sqrt(sum v^2), 1e-8); cross products as separately rounded mul / sub
    const float na = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2))), 1e-8f);
    const float x0 = __fdiv_rn(a0, na), x1 = __fdiv_rn(a1, na), x2 = __fdiv_rn(a2, na);
    float z0 = __fadd_rn(__fmul_rn(x1, b2), -__fmul_rn(x2, b1));
    float z1 = __fadd_rn(__fmul_rn(x2, b0), -__fmul_rn(x0, b2));
    float z2 = __fadd_rn(__fmul_rn(x0, b1), -__fmul_rn(x1, b0));
    const float nz = fmaxf(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(z0, z0), __fmul_rn(z1, z1)), __fmul_rn(z2, z2))), 1e-8f);
    z0 = __fdiv_rn(z0, nz); z1 = __fdiv_rn(z1, nz); z2 = __fdiv_rn(z2, nz);
    const float y0 = __fadd_rn(__fmul_rn(z1, x2), -__fmul_rn(z2, x1));
    const float y1 = __fadd_rn(__fmul_rn(z2, x0), -__fmul_rn(z0, x2));
    const float y2 = __fadd_rn(__fmul_rn(z0, x1), -__fmul_rn(z1, x0));
    const float R[9] = {x0, y0, z0, x1, y1, z1, x2, y2, z2};      // columns [x y z], row-major
    const float cx = p.center[r * 3], cy = p.center[r * 3 + 1], cz = p.center[r * 3 + 2];
    const float sx = p.size[r * 3], sy = p.size[r * 3 + 1], sz = p.size[r * 3 + 2];
    const float lo[3] = {__fdiv_rn(-sx, 2.f), __fdiv_rn(-sy, 2.f), __fdiv_rn(-sz, 2.f)};
    const float hi[3] = {__fdiv_rn(sx, 2.f), __fdiv_rn(sy, 2.f), __fdiv_rn(sz, 2.f)};
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      // corner order of Obb3D.bb3corners_object; only min / max matter here
      const float px = ((c & 1) ^ ((c >> 1) & 1)) ? hi[0] : lo[0];
      const float py = (c & 2) ? hi[1] : lo[1];
      const float pz = (c & 4) ? hi[2] : lo[2];
      const float t[3] = {cx, cy, cz};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float v = __fadd_rn(dot3_nofma(px, R[3 * i], py, R[3 * i + 1], pz, R[3 * i + 2]), t[i]);
        mn[i] = fminf(mn[i], v);
        mx[i] = fmaxf(mx[i], v);
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      s_aabb[k * 6 + i] = static_cast<double>(mn[i]);
      s_aabb[k * 6 + 3 + i] = static_cast<double>(mx[i]);
    }
    s_area[k] = (static_cast<double>(mx[0]) - mn[0]) * (static_cast<double>(mx[1]) - mn[1]) * (static_cast<double>(mx[2]) - mn[2]);
    s_label[k] = lab;
    if (p.scores != nullptr) p.scores[r] = best;
    if (p.labels != nullptr) p.labels[r] = lab;
    if (p.obbs != nullptr) {
      float* o = p.obbs + r * 19;
      o[0] = lo[0]; o[1] = hi[0]; o[2] = lo[1]; o[3] = hi[1]; o[4] = lo[2]; o[5] = hi[2];
#pragma unroll
      for (int i = 0; i < 9; ++i) o[6 + i] = R[i];
      o[15] = cx; o[16] = cy; o[17] = cz;
      o[18] = static_cast<float>(lab);
    }
    // probabilities are >= 0, so their bit pattern orders like the value; +1 keeps every real box above the padding
    if (lab != p.background) key = __float_as_uint(fmaxf(best, 0.f)) + 1u;
    in_scope = !p.apply_track_scale || ((cx > p.track_scale[0]) && (cx < p.track_scale[1]) && (cz > p.track_scale[4]) && (cz < p.track_scale[5]));
  }
  // ---- 2. bitonic sort, descending by (key, index)
  for (int i = k; i < P; i += blockDim.x) {
    s_key[i] = (i == k) ? key : 0u;
    s_idx[i] = i;
  }
  __syncthreads();
  for (int sz = 2; sz <= P; sz <<= 1) {
    for (int st = sz >> 1; st > 0; st >>= 1) {
      for (int i = k; i < P; i += blockDim.x) {
        const int j = i ^ st;
        if (j > i) {
          const uint32_t ki = s_key[i], kj = s_key[j];
          const int ii = s_idx[i], ij = s_idx[j];
          const bool i_first = (ki > kj) || (ki == kj && ii > ij);      // descending
          const bool desc = (i & sz) == 0;
          if (i_first != desc) {
            s_key[i] = kj; s_key[j] = ki;
            s_idx[i] = ij; s_idx[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
  // ---- 3. suppression rows: position pos suppresses later positions q
  for (int pos = k; pos < P; pos += blockDim.x) {
    const bool real = s_key[pos] != 0u;
    const int bi = s_idx[pos];
    for (int w = 0; w < W; ++w) {
      uint32_t bits = 0u;
      if (real) {
        for (int t = 0; t < 32; ++t) {
          const int q = w * 32 + t;
          if (q <= pos || s_key[q] == 0u) continue;
          const int bj = s_idx[q];
          const double l = fmax(0.0, fmin(s_aabb[bi * 6 + 3], s_aabb[bj * 6 + 3]) - fmax(s_aabb[bi * 6], s_aabb[bj * 6]));
          const double wd = fmax(0.0, fmin(s_aabb[bi * 6 + 4], s_aabb[bj * 6 + 4]) - fmax(s_aabb[bi * 6 + 1], s_aabb[bj * 6 + 1]));
          const double h = fmax(0.0, fmin(s_aabb[bi * 6 + 5], s_aabb[bj * 6 + 5]) - fmax(s_aabb[bi * 6 + 2], s_aabb[bj * 6 + 2]));
          const double inter = l * wd * h;
          double o = inter / (s_area[bi] + s_area[bj] - inter);
          if (p.same_class && s_label[bi] != s_label[bj]) o = 0.0;
          if (o > p.threshold) bits |= 1u << t;
        }
      }
      s_sup[static_cast<size_t>(pos) * W + w] = bits;
    }
  }
  __syncthreads();
  // ---- 4. greedy scan by warp 0: lane w owns word w of the removed / kept sets (P <= 1024)
  if (k < 32) {
    uint32_t removed = 0u, kept = 0u;
    for (int pos = 0; pos < P; ++pos) {
      if (s_key[pos] == 0u) break;                                        // padding / background boxes sort last
      const uint32_t word = __shfl_sync(0xffffffffu, removed, pos >> 5);
      if ((word >> (pos & 31)) & 1u) continue;
      if (k == (pos >> 5)) kept |= 1u << (pos & 31);
      if (k < W) removed |= s_sup[static_cast<size_t>(pos) * W + k];
    }
    if (k < W) s_keep[k] = kept;
  }
  __syncthreads();
  // ---- 5. back to box order
  for (int pos = k; pos < P; pos += blockDim.x) {
    if (s_key[pos] == 0u) continue;
    if ((s_keep[pos >> 5] >> (pos & 31)) & 1u) s_label[s_idx[pos]] |= 0x10000;    // mark kept
  }
  __syncthreads();
  if (k < p.K) {
    const long long r = static_cast<long long>(b) * p.K + k;
    const int kept = (s_label[k] >> 16) & 1;
    if (p.nms_mask != nullptr) p.nms_mask[r] = static_cast<uint8_t>(kept);
    p.pred_mask[r] = static_cast<uint8_t>(kept && in_scope);
  }
}

}  // namespace parq
