// libparq_b200.so -- C ABI (include/parq_b200.h) and host-side orchestration of the PARQ decoder
// hot path on sm_100a.  The library owns no device memory: weights, workspace and outputs are
// caller-provided raw device pointers; every launch goes to the caller's stream.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/parq_b200.h"
#include "attn_tc.cuh"
#include "attn2_tc.cuh"
#include "attn3_tc.cuh"
#include "chain_tc.cuh"
#include "gemm_sk.cuh"
#include "fpn.cuh"
#include "gemm_tc.cuh"
#include "gemm2_tc.cuh"
#include "parse_pred.cuh"
#include "project_sample.cuh"
#include "raype.cuh"
#include "rowwise.cuh"

namespace parq {

// ------------------------------------------------------------------------------------ errors --
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return fail(PARQ_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));   \
  } while (0)
#define TRY(expr)              \
  do {                         \
    int _r = (expr);           \
    if (_r != PARQ_OK) return _r; \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------- launch accounting / event profiling --
// Every kernel launch is counted (parq_kernel_launches).  When enabled with parq_profile_enable, launches
// whose tag is in the mask are bracketed by CUDA events on the launching stream; parq_profile_collect
// returns the summed device time and launch count per tag.
enum ProfTag { TAG_KV_PROJ = 0, TAG_SAMPLE = 1, TAG_GEMM = 2, TAG_SELF_ATTN = 3, TAG_CROSS_ATTN = 4, TAG_COMBINE = 5, TAG_ROWWISE = 6, TAG_COUNT = 8 };
struct Profiler {
  bool on = false;
  uint32_t mask = 0;
  int cap = 0, n = 0;
  cudaEvent_t* ev = nullptr;   // 2 per record
  int* tag = nullptr;
};
static thread_local Profiler g_prof;
static thread_local unsigned long long g_launches = 0;
struct ProfScope {
  cudaStream_t st;
  int idx = -1;
  ProfScope(int tag, cudaStream_t s, int nlaunch = 1) : st(s) {
    g_launches += nlaunch;
    if (g_prof.on && (g_prof.mask >> tag & 1u) && g_prof.n < g_prof.cap) {
      idx = g_prof.n++;
      g_prof.tag[idx] = tag;
      cudaEventRecord(g_prof.ev[2 * idx], st);
    }
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof.ev[2 * idx + 1], st);
  }
};

// ------------------------------------------------------------------------------- device info --
struct DeviceInfo {
  int ok = 0;     // 0 unknown, 1 sm_100, -1 other arch / error
  int sms = 148;
};
constexpr int MAX_DEVICES = 64;
static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return (dev >= 0 && dev < MAX_DEVICES) ? dev : 0;
}
// cached per device ordinal (a thread may drive several GPUs one after the other)
static DeviceInfo& device_info() {
  static DeviceInfo infos[MAX_DEVICES];
  static std::mutex mu;
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  DeviceInfo& info = infos[dev];
  if (info.ok == 0) {
    int major = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) {
      info.ok = (major == 10) ? 1 : -1;
      info.sms = sms > 0 ? sms : 148;
    } else {
      cudaGetLastError();
      info.ok = -1;
    }
  }
  return info;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) property of a kernel: the opt-in is
// remembered per (device ordinal, kernel), not per thread.
static int ensure_dyn_smem(const void* fn, size_t bytes) {
  struct Entry { const void* fn; size_t bytes; };
  static std::vector<Entry> tab[MAX_DEVICES];
  static std::mutex mu;
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  for (Entry& e : tab[dev]) {
    if (e.fn == fn) {
      if (e.bytes >= bytes) return PARQ_OK;
      CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
      e.bytes = bytes;
      return PARQ_OK;
    }
  }
  CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  tab[dev].push_back(Entry{fn, bytes});
  return PARQ_OK;
}
#define OPT_IN_SMEM(kernel, bytes) TRY(ensure_dyn_smem(reinterpret_cast<const void*>(kernel), bytes))

static int require_sm100() {
  DeviceInfo& d = device_info();
  if (d.ok != 1) return fail(PARQ_ERR_ARCH, "parq_b200 requires an sm_100 (B200) device; no fallback path exists");
  return PARQ_OK;
}

// ------------------------------------------------------------------------------------ launch --
// Inside parq_decoder_forward every kernel is launched with programmatic stream serialisation (PDL, see
// ptx.cuh): the next kernel's prologue overlaps the tail of the current one, also inside a captured graph.
// The stand-alone entry points launch plainly (their inputs may come from the caller's previous kernel).
static thread_local bool g_pdl = false;
static const bool g_no_dual = getenv("PARQ_NO_DUAL_A") != nullptr;
static const int g_combine_rows = (getenv("PARQ_COMBINE_ROWS") && (atoi(getenv("PARQ_COMBINE_ROWS")) == 4 || atoi(getenv("PARQ_COMBINE_ROWS")) == 16 ||
                                                                   atoi(getenv("PARQ_COMBINE_ROWS")) == 32)) ? atoi(getenv("PARQ_COMBINE_ROWS")) : SK_COMBINE_ROWS;
static const bool g_no_narrow = getenv("PARQ_NO_NARROW") != nullptr;         // A/B switch: 256-column tiles also for GEMMs of a few row tiles
static const bool g_no_chain_v = getenv("PARQ_NO_CHAIN_V") != nullptr;       // A/B switch: self-attention V^T as its own GEMM launch instead of stage 0 of chain P
static const bool g_no_chain = getenv("PARQ_NO_CHAIN") != nullptr;
static const bool g_no_splitk = getenv("PARQ_NO_SPLITK") != nullptr;          // A/B switch: narrow single-CTA tiles instead of the cluster split-K GEMM (gemm_sk.cuh)
static const bool g_no_fork = getenv("PARQ_NO_FORK") != nullptr;              // A/B switch: every launch of the un-chained path on the caller's stream
static const bool g_fused_merge = getenv("PARQ_FUSED_MERGE") != nullptr;  // opt-in: merge the stream-K pieces inside the attention kernel instead of the attn3_combine_kernel launch
static const int g_chain_min_rows = getenv("PARQ_CHAIN_MIN_ROWS") ? atoi(getenv("PARQ_CHAIN_MIN_ROWS")) : 2048;
// Default: every GEMM keeps its low-order activation term.  Dropping it for the three GEMMs whose output is rounded to 16 bits
// right away (mask 7: self-attention Q|K and V^T, cross-attention Q) buys -0.25 +- 0.06 ms per step at config 2 (in-process A/B,
// tools/ab_step.py: the step is power-capped, removed tensor work is removed time) and stays inside the 1e-3 bar (1.0e-4 ->
// 4-5e-4), but it flips the arg-max class of a near-tied query of the reference's own fixture (and 2 of 32 768 at config 2),
// which changes size_unnormalized of that query visibly: parity first, so it stays an opt-in (PARQ_FLAG_HI_ONLY_SET, PARQ_HI_ONLY).
constexpr int HI_ONLY_DEFAULT = 0;
static const int g_hi_only = getenv("PARQ_HI_ONLY") ? atoi(getenv("PARQ_HI_ONLY")) : 0;   // ablation: bit0 sa_qk, bit1 sa_v, bit2 ca_q use the hi activation term only           // A/B switch: separate GEMM + LayerNorm launches instead of chain_tc.cuh
static const bool g_force_pair = getenv("PARQ_FORCE_PAIR") != nullptr;   // every GEMM on the CTA-pair kernel (tests)
static const bool g_no_streamk = getenv("PARQ_NO_STREAMK") != nullptr;       // A/B switch: split-KV grid instead of the stream-K schedule
static const bool g_no_pair_attn = getenv("PARQ_NO_PAIR_ATTN") != nullptr;   // A/B switch: single-CTA attention kernel
static const bool g_no_pair = getenv("PARQ_NO_PAIR") != nullptr;       // A/B switch: single-CTA GEMM instead of the CTA-pair kernel   // A/B switch for the shared-B ring layout
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(g_pdl) { g_pdl = on; }
  ~PdlScope() { g_pdl = prev; }
};
template <typename... KArgs, typename... Args>
static void launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (g_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster.x;
    attr[n].val.clusterDim.y = cluster.y;
    attr[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_kc(kernel, grid, block, smem, st, dim3(1, 1, 1), static_cast<Args&&>(args)...);
}

// ------------------------------------------------------------------------------ side streams --
// Independent launches of one iteration run on two side streams of the library (fork / join through events; inside a stream
// capture they become parallel branches of the graph): with one clip an iteration is a chain of ~20 short dependent launches
// and every launch taken off that chain is ~10 us of latency.  Created lazily outside a capture, per thread and device.
struct SideStreams {
  static constexpr int NEV = 128;          // events are used round-robin: none is re-recorded within one forward (<= 4 per iteration)
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t ev[NEV] = {};
  int next = 0;
  bool ok = false;
  cudaEvent_t take() { cudaEvent_t e = ev[next]; next = (next + 1) % NEV; return e; }
};
static SideStreams* side_streams(cudaStream_t st) {
  static thread_local SideStreams tab[MAX_DEVICES];
  SideStreams& S = tab[current_device()];
  if (!S.ok) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
      cudaGetLastError();
      return nullptr;            // first use inside a capture: stay on one stream
    }
    for (int i = 0; i < 2; ++i) {
      if (cudaStreamCreateWithFlags(&S.s[i], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
    }
    for (int i = 0; i < SideStreams::NEV; ++i) {
      if (cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
    }
    S.ok = true;
  }
  return &S;
}
static int side_fork(cudaStream_t st, SideStreams* S, int i) {       // work launched on S->s[i] from now on starts after what `st` holds
  cudaEvent_t e = S->take();
  CUDA_TRY(cudaEventRecord(e, st));
  CUDA_TRY(cudaStreamWaitEvent(S->s[i], e, 0));
  return PARQ_OK;
}
static int side_join(cudaStream_t st, SideStreams* S, int i) {       // `st` continues after what S->s[i] holds
  cudaEvent_t e = S->take();
  CUDA_TRY(cudaEventRecord(e, S->s[i]));
  CUDA_TRY(cudaStreamWaitEvent(st, e, 0));
  return PARQ_OK;
}

// ------------------------------------------------------------------------------ TMA tensor maps --
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 2-D map over a row-major 16-bit matrix: `cols` contiguous elements, `rows` rows of pitch ld elements;
// box = 64 columns (128 bytes, SWIZZLE_128B) x box_rows.
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(PARQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(PARQ_ERR_SHAPE, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PARQ_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return PARQ_OK;
}

// 2-D map over a row-major fp32 matrix without swizzle: box = box_cols x box_rows (the channels-first addend of gemm_tc.cuh)
static int make_map_f32(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows, bool bf16 = false) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(PARQ_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PARQ_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", static_cast<int>(r));
  return PARQ_OK;
}

// ---------------------------------------------------------------------------------- launchers --
static int launch_gemm(cudaStream_t st, const void* A, uint64_t a_rows, uint64_t a_cols, const void* Bw, uint64_t b_rows,
                       uint64_t b_cols, const GemmParams& gp, int tag = TAG_GEMM) {
  if (gp.K <= 0 || gp.K % gemm::BK != 0) return fail(PARQ_ERR_SHAPE, "GEMM K=%d must be a positive multiple of 64", gp.K);
  if (gp.nterms < 1 || gp.nterms > 3) return fail(PARQ_ERR_SHAPE, "GEMM nterms=%d out of range", gp.nterms);
  // CTA-pair kernel (gemm2_tc.cuh, each CTA stages half of the B tile) for the multi-wave GEMMs (hoisted K / V^T
  // projection, AddRayPE encoder: +6 % measured); the one-wave per-iteration GEMMs stay on the single-CTA kernel, whose
  // shorter prologue (no cluster barriers) is worth more there (-6 % with the pair kernel)
  const long long tiles1 = static_cast<long long>((gp.M + gemm::BM - 1) / gemm::BM) * ((gp.N + gemm::BN - 1) / gemm::BN);
  const bool pairk = !g_no_pair && device_info().sms >= 2 && (tiles1 >= 2LL * device_info().sms || g_force_pair);
  // narrow tiles for GEMMs of a few row tiles (one clip): 4x the CTAs, a quarter of the MMA latency each (not for the
  // channels-first epilogues; the GroupNorm tile sums have one slot per 64 columns)
  const bool narrow = !pairk && !g_no_narrow && gp.ep.nchw_add == nullptr && gp.ep.nchw_out == nullptr && gp.N % 64 == 0 &&
                      tiles1 * 4 <= device_info().sms;
  // cluster split-K kernel (gemm_sk.cuh) for the same few-row GEMMs when the shape fits: activations [hi|lo] x bf16-exact weights
  const int kblocks = gp.K / gemm::BK;
  const bool splitk = narrow && !g_no_splitk && gp.nterms == 2 && gp.b_koff[0] == gp.b_koff[1] && gp.a_koff[0] != gp.a_koff[1] && gp.const_operand != 1 &&
                      !gp.ep.bias_per_row && gp.ep.kv_tiled == 0 && gp.N % gemmsk::BN == 0 && kblocks % gemmsk::S == 0 &&
                      kblocks / gemmsk::S <= gemmsk::MAX_STEPS && (gp.a_split_n == 0 || gp.a_split_n % gemmsk::BN == 0) &&
                      // one wave of clusters: beyond that the output tiles alone fill the machine
                      static_cast<long long>(gemmsk::S) * ((gp.M + gemmsk::BM - 1) / gemmsk::BM) * (gp.N / gemmsk::BN) <= device_info().sms;
  if (splitk) {
    CUtensorMap tmA, tmB;
    TRY(make_map(&tmA, A, a_rows, a_cols, a_cols, gemmsk::BM));
    TRY(make_map(&tmB, Bw, b_rows, b_cols, b_cols, gemmsk::BN));
    OPT_IN_SMEM(gemm_sk_kernel, gemmsk::SMEM_BYTES);
    const int tiles = ((gp.M + gemmsk::BM - 1) / gemmsk::BM) * (gp.N / gemmsk::BN);
    {
      ProfScope ps(tag, st);
      launch_kc(gemm_sk_kernel, dim3(gemmsk::S * tiles), dim3(gemmsk::THREADS), gemmsk::SMEM_BYTES, st, dim3(gemmsk::S, 1, 1), tmA, tmB, gp);
    }
    CUDA_TRY(cudaGetLastError());
    return PARQ_OK;
  }
  const int bn = narrow ? 64 : gemm::BN;
  CUtensorMap tmA, tmB;
  TRY(make_map(&tmA, A, a_rows, a_cols, a_cols, gemm::BM));
  TRY(make_map(&tmB, Bw, b_rows, b_cols, b_cols, pairk ? gemm2::BN / 2 : bn));
  OPT_IN_SMEM(gemm_tc_kernel<false>, gemm::SMEM_BYTES);
  OPT_IN_SMEM(gemm_tc_kernel<true>, gemm::SMEM_BYTES);
  OPT_IN_SMEM(gemm2_tc_kernel<false>, gemm2::SMEM_BYTES);
  OPT_IN_SMEM(gemm2_tc_kernel<true>, gemm2::SMEM_BYTES);
  GemmParams gpl = gp;
  CUtensorMap tmC = tmA;                    // placeholder unless the channels-first addend goes through TMA
  gpl.ep.add_tma = 0;
  if (gp.ep.nchw_add != nullptr && gp.ep.nchw_HW % (gp.ep.nchw_add_bf16 ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(gp.ep.nchw_add) & 15) == 0 &&
      gp.M % gp.ep.nchw_HW == 0) {
    TRY(make_map_f32(&tmC, gp.ep.nchw_add, static_cast<uint64_t>(gp.M / gp.ep.nchw_HW) * gp.N, gp.ep.nchw_HW, 128, 32, gp.ep.nchw_add_bf16 != 0));
    gpl.ep.add_tma = 1;
  }
  gpl.bn = bn;
  gpl.dual_a = (gp.nterms == 2 && gp.b_koff[0] == gp.b_koff[1] && gp.a_koff[0] != gp.a_koff[1] && !g_no_dual) ? 1 : 0;
  const int tiles = ((gp.M + gemm::BM - 1) / gemm::BM) * ((gp.N + bn - 1) / bn);
  const int grid = tiles < device_info().sms ? tiles : device_info().sms;
  {
    ProfScope ps(tag, st);
    const bool nchw = gp.ep.nchw_add != nullptr || gp.ep.nchw_out != nullptr;
    if (pairk) {
      const int ptiles = ((gp.M + 2 * gemm2::BM - 1) / (2 * gemm2::BM)) * ((gp.N + gemm2::BN - 1) / gemm2::BN);
      const int pgrid = 2 * (ptiles < device_info().sms / 2 ? ptiles : device_info().sms / 2);
      if (nchw)
        launch_kc(gemm2_tc_kernel<true>, dim3(pgrid), dim3(gemm2::THREADS), gemm2::SMEM_BYTES, st, dim3(2, 1, 1), tmA, tmB, tmC, gpl);
      else
        launch_kc(gemm2_tc_kernel<false>, dim3(pgrid), dim3(gemm2::THREADS), gemm2::SMEM_BYTES, st, dim3(2, 1, 1), tmA, tmB, tmC, gpl);
    } else if (nchw)
      launch_k(gemm_tc_kernel<true>, dim3(grid), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB, tmC, gpl);
    else
      launch_k(gemm_tc_kernel<false>, dim3(grid), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB, tmC, gpl);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}

// ---- chained GEMMs (chain_tc.cuh): one cluster of 4 CTAs per 128 rows walks a list of dependent linear stages ----
struct ChainBuilder {
  ChainParams p;
  ChainMaps m;
  ChainBuilder(int M) {
    memset(&p, 0, sizeof(p));
    memset(&m, 0, sizeof(m));
    p.M = M;
  }
};
static bool chain_cols_ok(int N) {
  const int ncta = N / chain::CLUSTER;
  if (N % chain::CLUSTER != 0 || ncta > chain::VEC_COLS) return false;
  return ncta <= 256 ? (ncta % 64 == 0) : (ncta % 256 == 0);
}
// A = [hi|lo] activations (M, a_cols), W = packed [hi|lo] weights (N, 2K); S carries N, K and the epilogue
static int chain_add(ChainBuilder& cb, const void* A, uint64_t a_cols, const void* W, bool w_lo, ChainStage S) {
  if (cb.p.nstages >= chain::MAX_STAGES) return fail(PARQ_ERR_SHAPE, "too many chain stages");
  if (S.K <= 0 || S.K % chain::BK != 0 || !chain_cols_ok(S.N)) return fail(PARQ_ERR_SHAPE, "chain stage N=%d K=%d not supported", S.N, S.K);
  const int ncta = S.N / chain::CLUSTER;
  static const int tile_cap = getenv("PARQ_CHAIN_TILE") ? atoi(getenv("PARQ_CHAIN_TILE")) : 256;   // experiment switch (tools/chain_timeline.py)
  S.tile_n = ncta <= tile_cap ? ncta : (ncta % tile_cap == 0 ? tile_cap : ncta);
  if (S.tile_n > 256) S.tile_n = 256;
  S.tiles = ncta / S.tile_n;
  S.nterms = w_lo ? 3 : 2;
  S.a_koff[0] = 0;   S.b_koff[0] = 0;
  S.a_koff[1] = S.K; S.b_koff[1] = 0;
  S.a_koff[2] = 0;   S.b_koff[2] = S.K;
  S.dual_a = (!w_lo && !g_no_dual) ? 1 : 0;
  if (cb.p.nstages > 0 && cb.p.st[0].dual_a != S.dual_a) return fail(PARQ_ERR_SHAPE, "chain stages must share the ring geometry");
  const int i = cb.p.nstages++;
  if (S.dep == -2) S.dep = i - 1;
  if (S.dep >= i) return fail(PARQ_ERR_SHAPE, "chain stage %d cannot depend on stage %d", i, S.dep);
  TRY(make_map(&cb.m.a[i], A, cb.p.M, a_cols, a_cols, chain::BM));
  TRY(make_map(&cb.m.b[i], W, S.N, 2 * static_cast<uint64_t>(S.K), 2 * static_cast<uint64_t>(S.K), S.tile_n));
  cb.p.st[i] = S;
  return PARQ_OK;
}
static thread_local long long* g_chain_dbg = nullptr;     // parq_chain_debug: device buffer for clock stamps, 64 slots per chain launch
static thread_local int g_chain_dbg_launch = 0;
static int g_trace_chain_slot = -1;                        // parq_trace: next per-CTA stamp block of a chain launch, -1 = trace off
static int launch_chain(cudaStream_t st, ChainBuilder& cb) {
  cb.p.trace_slot = g_trace_chain_slot >= 0 ? g_trace_chain_slot++ : -1;
  if (g_chain_dbg != nullptr && g_chain_dbg_launch < 64) cb.p.dbg = g_chain_dbg + 64 * g_chain_dbg_launch++;
  if (cb.p.M % chain::BM != 0) return fail(PARQ_ERR_SHAPE, "chain kernel needs M %% 128 == 0 (M=%d)", cb.p.M);
  OPT_IN_SMEM(chain_tc_kernel, chain::SMEM_BYTES);
  {
    ProfScope ps(TAG_GEMM, st);
    launch_kc(chain_tc_kernel, dim3(chain::CLUSTER * (cb.p.M / chain::BM)), dim3(chain::THREADS), chain::SMEM_BYTES, st, dim3(chain::CLUSTER, 1, 1),
              cb.m, cb.p);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}
static ChainStage chain_stage(int N, int K, int ep, const float* bias) {
  ChainStage S;
  memset(&S, 0, sizeof(S));
  S.N = N; S.K = K; S.ep = ep; S.bias = bias;
  S.dep = -2;                // chain_add: the previous stage (set -1 for a stage that reads no earlier stage of the launch)
  return S;
}

struct SplitPlan {
  int nsplit, tiles_per_split;
};
// Key-split heuristic: minimise waves x (tiles per CTA + fixed per-CTA overhead) over the SM count.
static SplitPlan plan_split(int items, int ntiles, int sms, int force) {
  SplitPlan best{1, ntiles};
  if (force > 0) {
    int tps = (ntiles + force - 1) / force;
    return SplitPlan{(ntiles + tps - 1) / tps, tps};
  }
  double best_cost = 1e30;
  const int smax = ntiles < 32 ? ntiles : 32;
  for (int s = 1; s <= smax; ++s) {
    const int tps = (ntiles + s - 1) / s;
    const int ns = (ntiles + tps - 1) / tps;
    const long long ctas = static_cast<long long>(items) * ns;
    const long long waves = (ctas + sms - 1) / sms;
    // small penalty per extra partial; any split at all costs the combine launch (about three key tiles of latency)
    const double cost = static_cast<double>(waves) * (tps + 3.0) + 0.02 * ns + (ns > 1 ? 3.0 : 0.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = SplitPlan{ns, tps};
    }
  }
  return best;
}

static size_t attn_scratch_bytes(int B, int H, int Nq, int nsplit) {
  const size_t rows = static_cast<size_t>(B) * H * nsplit * Nq;
  return align_up(rows * 256 * sizeof(float), 256) + align_up(rows * sizeof(float2), 256);
}

// scratch of the stream-K schedule (attn3_tc.cuh): one slot of 256 rows per (pair, segment)
static size_t streamk_scratch_bytes(int B, int H, int Nq, int Nk, int sms) {
  if (Nq % (2 * attn::BQ) != 0 || sms < 2) return 0;
  const long long ntiles = (Nk + attn::BKEY - 1) / attn::BKEY;
  const long long units = static_cast<long long>(B) * H * (Nq / (2 * attn::BQ)) * ntiles;
  const long long cap = sms / 2 < SK_MAX_PAIRS ? sms / 2 : SK_MAX_PAIRS;
  const long long npairs = units < cap ? units : cap;
  const long long per = (units + npairs - 1) / npairs;
  const size_t rows = static_cast<size_t>(npairs) * ((per + ntiles - 1) / ntiles + 1) * 256;
  return align_up(rows * 256 * sizeof(float), 256) + align_up(rows * sizeof(float2), 256);
}

static int launch_attention(cudaStream_t st, const void* Q, uint64_t ldq, const void* K, uint64_t ldk, const void* Vt,
                            uint64_t ldv, int B, int H, int Nq, int Nk, bool fp16, void* scratch, size_t scratch_bytes,
                            __nv_bfloat16* out_split, int force_nsplit, bool kv_const = false, bool kv_tiled = false,
                            uint32_t* sk_flags = nullptr) {
  if (Nq % attn::BQ != 0) return fail(PARQ_ERR_SHAPE, "Nq=%d must be a multiple of 128", Nq);
  const int ntiles = (Nk + attn::BKEY - 1) / attn::BKEY;
  const SplitPlan plan = plan_split(B * H * (Nq / attn::BQ), ntiles, device_info().sms, force_nsplit < 0 ? 0 : force_nsplit);
  if (attn_scratch_bytes(B, H, Nq, plan.nsplit) > scratch_bytes)
    return fail(PARQ_ERR_WORKSPACE, "attention scratch too small: need %zu, have %zu", attn_scratch_bytes(B, H, Nq, plan.nsplit),
                scratch_bytes);
  if (!kv_tiled && B > 1 && Nk % 8 != 0)
    return fail(PARQ_ERR_SHAPE, "plain V^T layout: clip b starts at column b*Nk, TMA needs 16-byte aligned box origins -> Nk %% 8 == 0 (Nk=%d)", Nk);
  const uint64_t C = static_cast<uint64_t>(H) * 256;
  // CTA-pair kernel (attn2_tc.cuh): the two 128-query tiles of 256 queries share every K / V^T tile, each CTA stages half
  const bool pairk = !g_no_pair_attn && Nq % (2 * attn::BQ) == 0 && device_info().sms >= 2;
  const uint32_t kbox = pairk ? attn::BKEY / 2 : attn::BKEY, vbox = pairk ? 128 : 256;
  CUtensorMap tmQ, tmK, tmV;
  TRY(make_map(&tmQ, Q, static_cast<uint64_t>(B) * Nq, C, ldq, attn::BQ));
  if (kv_tiled) {
    // tile-contiguous caches written by the K / V^T projection GEMMs (GemmEpilogue::kv_tiled)
    const uint64_t blocks = static_cast<uint64_t>(B) * ntiles * H;
    TRY(make_map(&tmK, K, blocks * attn::BKEY, 256, 256, kbox));
    TRY(make_map(&tmV, Vt, blocks * 256, attn::BKEY, attn::BKEY, vbox));
  } else {
    TRY(make_map(&tmK, K, static_cast<uint64_t>(B) * Nk, C, ldk, kbox));
    TRY(make_map(&tmV, Vt, C, static_cast<uint64_t>(B) * Nk, ldv, vbox));
  }
  OPT_IN_SMEM(attn3_tc_kernel<false>, attn::SMEM_BYTES);
  OPT_IN_SMEM(attn3_tc_kernel<true>, attn::SMEM_BYTES);
  // Stream-K schedule (attn3_tc.cuh) for long key sequences: one CTA pair per SM pair walks a contiguous range of
  // (item, key tile) units; used when every pair gets at least 8 tiles and no split count was forced by the caller.
  if (pairk && !g_no_streamk && force_nsplit <= 0) {
    Attn3Params sp;
    memset(&sp, 0, sizeof(sp));
    sp.B = B; sp.H = H; sp.Nq = Nq; sp.Nk = Nk;
    sp.qpairs = Nq / (2 * attn::BQ);
    sp.ntiles = ntiles;
    sp.units = static_cast<long long>(B) * H * sp.qpairs * ntiles;
    const long long maxp = device_info().sms / 2 < SK_MAX_PAIRS ? device_info().sms / 2 : SK_MAX_PAIRS;
    sp.npairs = static_cast<int>(sp.units < maxp ? sp.units : maxp);
    const long long per = (sp.units + sp.npairs - 1) / sp.npairs;
    sp.slots = static_cast<int>((per + ntiles - 1) / ntiles + 1);
    const size_t rows3 = static_cast<size_t>(sp.npairs) * sp.slots * 256;
    const size_t need3 = align_up(rows3 * 256 * sizeof(float), 256) + align_up(rows3 * sizeof(float2), 256);
    if ((per >= 8 || force_nsplit < 0) && need3 <= scratch_bytes) {
      sp.o_part = reinterpret_cast<float*>(scratch);
      sp.ml_part = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(scratch) + align_up(rows3 * 256 * sizeof(float), 256));
      sp.out = out_split;
      sp.kv_const = kv_const ? 1 : 0;
      sp.kv_tiled = kv_tiled ? 1 : 0;
      // fused merge (attn3_tc.cuh) when no item is cut into more than three pieces: the shortest range covers half an item
      sp.flags = (sk_flags != nullptr && 2 * (sp.units / sp.npairs) >= ntiles && 2 * sp.npairs * 4 <= SK_FLAG_WORDS) ? sk_flags : nullptr;
      {
        ProfScope ps(fp16 ? TAG_SELF_ATTN : TAG_CROSS_ATTN, st);
        if (fp16)
          launch_kc(attn3_tc_kernel<true>, dim3(2, sp.npairs, 1), dim3(attn::THREADS), attn::SMEM_BYTES, st, dim3(2, 1, 1), tmQ, tmK, tmV, sp);
        else
          launch_kc(attn3_tc_kernel<false>, dim3(2, sp.npairs, 1), dim3(attn::THREADS), attn::SMEM_BYTES, st, dim3(2, 1, 1), tmQ, tmK, tmV, sp);
      }
      CUDA_TRY(cudaGetLastError());
      if (sp.flags == nullptr) {
        ProfScope ps(TAG_COMBINE, st);
        const int items = B * H * sp.qpairs;
        if (2 * items <= sp.npairs)      // every item is cut into >= 3 pieces: 2 rows per block, the block's thread groups share a row
          launch_k(attn3_combine_kernel<true>, dim3(items * (256 / 2)), dim3(256), 0, st, sp, 2);
        else
          launch_k(attn3_combine_kernel<false>, dim3(items * (256 / g_combine_rows)), dim3(256), 0, st, sp, g_combine_rows);
      }
      CUDA_TRY(cudaGetLastError());
      return PARQ_OK;
    }
  }
  AttnParams ap;
  ap.B = B; ap.H = H; ap.Nq = Nq; ap.Nk = Nk;
  ap.nsplit = plan.nsplit;
  ap.tiles_per_split = plan.tiles_per_split;
  ap.out_direct = plan.nsplit == 1 ? out_split : nullptr;
  ap.kv_const = kv_const ? 1 : 0;
  ap.kv_tiled = kv_tiled ? 1 : 0;
  ap.ntile = ntiles;
  const size_t rows = static_cast<size_t>(B) * H * plan.nsplit * Nq;
  ap.o_part = reinterpret_cast<float*>(scratch);
  ap.ml_part = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(scratch) + align_up(rows * 256 * sizeof(float), 256));
  OPT_IN_SMEM(attn_tc_kernel<false>, attn::SMEM_BYTES);
  OPT_IN_SMEM(attn_tc_kernel<true>, attn::SMEM_BYTES);
  OPT_IN_SMEM(attn2_tc_kernel<false>, attn::SMEM_BYTES);
  OPT_IN_SMEM(attn2_tc_kernel<true>, attn::SMEM_BYTES);
  dim3 grid(plan.nsplit, Nq / attn::BQ, B * H);
  {
    ProfScope ps(fp16 ? TAG_SELF_ATTN : TAG_CROSS_ATTN, st);
    const dim3 pgrid(Nq / attn::BQ, plan.nsplit, B * H);        // the CTA pair = the two query tiles, adjacent in x
    if (pairk && fp16)
      launch_kc(attn2_tc_kernel<true>, pgrid, dim3(attn::THREADS), attn::SMEM_BYTES, st, dim3(2, 1, 1), tmQ, tmK, tmV, ap);
    else if (pairk)
      launch_kc(attn2_tc_kernel<false>, pgrid, dim3(attn::THREADS), attn::SMEM_BYTES, st, dim3(2, 1, 1), tmQ, tmK, tmV, ap);
    else if (fp16)
      launch_k(attn_tc_kernel<true>, grid, dim3(attn::THREADS), attn::SMEM_BYTES, st, tmQ, tmK, tmV, ap);
    else
      launch_k(attn_tc_kernel<false>, grid, dim3(attn::THREADS), attn::SMEM_BYTES, st, tmQ, tmK, tmV, ap);
  }
  CUDA_TRY(cudaGetLastError());
  if (plan.nsplit > 1) {
    ProfScope ps(TAG_COMBINE, st);
    launch_k(attn_combine_kernel, dim3(B * Nq), dim3(256), 0, st, ap.o_part, ap.ml_part, out_split, H, Nq, plan.nsplit);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}

// ------------------------------------------------------------------------- packed weight layout --
struct Packed {
  // bf16 [rows, 2K] = [hi | lo]
  size_t pe0, pe2, sa_qk, sa_v, sa_out, ca_q, ca_k, ca_v, ca_out, lin1, lin2, hd1, ctr4, rot4;
  // fp32 vectors / small matrices
  size_t pe0_b, pe2_b, sa_qk_b, sa_v_b, sa_out_b, ca_q_b, ca_k_b, ca_v_b, ca_out_b, lin1_b, lin2_b;
  size_t ln1_g, ln1_b, ln2_g, ln2_b, ln3_g, ln3_b;
  size_t ctr1_g, ctr1_b, rot1_g, rot1_b, ctr5_g, ctr5_b, rot5_g, rot5_b;
  size_t cls_w, cls_b, size_w, size_b, ctr8_w, ctr8_b, rot8_w, rot8_b, mean_size, dim_t;
  size_t lo_flag;
  size_t total;
};
static Packed packed_layout(const ParqShape& s) {
  Packed p;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 256);
    return o;
  };
  const size_t C = s.C, F = s.ffn, PE = 384;
  auto mat = [&](size_t rows, size_t K) { return take(rows * 2 * K * 2); };
  p.pe0 = mat(C, PE);  p.pe2 = mat(C, C);
  p.sa_qk = mat(2 * C, C);  p.sa_v = mat(C, C);  p.sa_out = mat(C, C);
  p.ca_q = mat(C, C);  p.ca_k = mat(C, C);  p.ca_v = mat(C, C);  p.ca_out = mat(C, C);
  p.lin1 = mat(F, C);  p.lin2 = mat(C, F);
  p.hd1 = mat(2 * C, C);  p.ctr4 = mat(C, C);  p.rot4 = mat(C, C);
  auto vec = [&](size_t n) { return take(n * 4); };
  p.pe0_b = vec(C);  p.pe2_b = vec(C);  p.sa_qk_b = vec(2 * C);  p.sa_v_b = vec(C);  p.sa_out_b = vec(C);
  p.ca_q_b = vec(C);  p.ca_k_b = vec(C);  p.ca_v_b = vec(C);  p.ca_out_b = vec(C);  p.lin1_b = vec(F);  p.lin2_b = vec(C);
  p.ln1_g = vec(C);  p.ln1_b = vec(C);  p.ln2_g = vec(C);  p.ln2_b = vec(C);  p.ln3_g = vec(C);  p.ln3_b = vec(C);
  p.ctr1_g = vec(C);  p.ctr1_b = vec(C);  p.rot1_g = vec(C);  p.rot1_b = vec(C);
  p.ctr5_g = vec(C);  p.ctr5_b = vec(C);  p.rot5_g = vec(C);  p.rot5_b = vec(C);
  p.cls_w = vec(static_cast<size_t>(s.num_cls) * C);  p.cls_b = vec(s.num_cls);
  p.size_w = vec(3 * C);  p.size_b = vec(3);  p.ctr8_w = vec(3 * C);  p.ctr8_b = vec(3);
  p.rot8_w = vec(6 * C);  p.rot8_b = vec(6);
  p.mean_size = vec(static_cast<size_t>(s.num_cls) * 3);  p.dim_t = vec(128);
  p.lo_flag = take(4);
  p.total = off;
  return p;
}

// dst[r, k] = bf16(mul*src[r,k]), dst[r, K+k] = bf16 residual; flags a non-zero residual.
__global__ void split_weight_kernel(const float* __restrict__ src, int rows, int K, float mul, __nv_bfloat16* __restrict__ dst,
                                    int* __restrict__ lo_flag) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * K) return;
  const long long r = i / K, k = i % K;
  const float v = src[i] * mul;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  dst[r * 2 * K + k] = h;
  dst[r * 2 * K + K + k] = l;
  if (__bfloat162float(l) != 0.f) atomicOr(lo_flag, 1);
}
__global__ void copy_scale_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, float mul) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * mul;
}

// --------------------------------------------------------------------------- workspace layout --
struct Workspace {
  size_t Kc, Vt, T_cl, ref_cur, a_pos, a_peh, pe, a_x, a_xpe, qk_s, vt_s, scratch, a_attn, y, x1, x2, x3, a_x1pe, q_c,
      a_x2, a_ffn, a_x3, h1, a_h1, h2, gn1, gn2, sk_flags;
  size_t scratch_bytes, ldv, ldvs, total;
  int kv_tiled, ntile;      // tile-contiguous K / V^T caches (needs Nk % 32 == 0), key tiles per clip
  SplitPlan cross, self;
};
static Workspace workspace_layout(const ParqShape& s, int sms) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 1024);
    return o;
  };
  const size_t C = s.C, R = static_cast<size_t>(s.B) * s.Nq, Nk = static_cast<size_t>(s.T) * s.H * s.W, Nt = s.B * Nk;
  w.ldv = align_up(Nt, 64);
  w.ldvs = align_up(R, 64);
  w.ntile = static_cast<int>((Nk + attn::BKEY - 1) / attn::BKEY);
  w.kv_tiled = 1;            // every Nk: ragged clips (Nk % 32 != 0) take the per-element path of the GEMM epilogue
  const size_t cache = w.kv_tiled ? static_cast<size_t>(s.B) * w.ntile * attn::BKEY * C * 2 : 0;
  w.Kc = take(w.kv_tiled ? cache : Nt * C * 2);
  w.Vt = take(w.kv_tiled ? cache : C * w.ldv * 2);
  w.T_cl = take(static_cast<size_t>(s.B) * s.T * 12 * 4);
  w.ref_cur = take(R * 3 * 4);
  w.a_pos = take(R * 768 * 2);
  w.a_peh = take(R * 2 * C * 2);
  w.pe = take(R * C * 4);
  w.a_x = take(R * 2 * C * 2);
  w.a_xpe = take(R * 2 * C * 2);
  w.qk_s = take(R * 2 * C * 2);
  w.vt_s = take(C * w.ldvs * 2);
  const int qtiles = s.Nq / attn::BQ;
  w.cross = plan_split(s.B * s.heads * qtiles, static_cast<int>((Nk + attn::BKEY - 1) / attn::BKEY), sms, 0);
  w.self = plan_split(s.B * s.heads * qtiles, (s.Nq + attn::BKEY - 1) / attn::BKEY, sms, 0);
  const int smax = w.cross.nsplit > w.self.nsplit ? w.cross.nsplit : w.self.nsplit;
  w.scratch_bytes = attn_scratch_bytes(s.B, s.heads, s.Nq, smax);
  const size_t sk = streamk_scratch_bytes(s.B, s.heads, s.Nq, static_cast<int>(Nk), sms);
  if (sk > w.scratch_bytes) w.scratch_bytes = sk;
  w.scratch = take(w.scratch_bytes);
  w.a_attn = take(R * 2 * C * 2);
  w.y = take(R * C * 4);
  w.x1 = take(R * C * 4);
  w.x2 = take(R * C * 4);
  w.x3 = take(R * C * 4);
  w.a_x1pe = take(R * 2 * C * 2);
  w.q_c = take(R * C * 2);
  w.a_x2 = take(R * 2 * C * 2);
  w.a_ffn = take(R * 2 * static_cast<size_t>(s.ffn) * 2);
  w.a_x3 = take(R * 2 * C * 2);
  w.h1 = take(R * 2 * C * 4);
  w.a_h1 = take(R * 4 * C * 2);
  w.h2 = take(R * 2 * C * 4);
  w.gn1 = take(R / 128 * GN_SLOTS_PER_MTILE * sizeof(double2));
  w.gn2 = take(R / 128 * GN_SLOTS_PER_MTILE * sizeof(double2));
  w.sk_flags = take(SK_FLAG_WORDS * sizeof(uint32_t));      // fused merge of the stream-K cross-attention: zero between launches
  w.total = off;
  return w;
}

static int check_shape(const ParqShape* s) {
  if (s == nullptr) return fail(PARQ_ERR_SHAPE, "null shape");
  if (s->B < 1 || s->T < 1 || s->H < 1 || s->W < 1 || s->iters < 1) return fail(PARQ_ERR_SHAPE, "non-positive dimension");
  if (s->C != 1024) return fail(PARQ_ERR_SHAPE, "C=%d: this build supports the reference width C=1024 only", s->C);
  if (s->heads * 256 != s->C) return fail(PARQ_ERR_SHAPE, "head_dim must be 256 (C=%d, heads=%d)", s->C, s->heads);
  if (s->Nq % 128 != 0 || s->Nq < 128) return fail(PARQ_ERR_SHAPE, "Nq=%d must be a positive multiple of 128", s->Nq);
  if (s->ffn % 64 != 0 || s->ffn < 64) return fail(PARQ_ERR_SHAPE, "ffn=%d must be a positive multiple of 64", s->ffn);
  if (s->num_cls < 1 || s->num_cls > 16) return fail(PARQ_ERR_SHAPE, "num_cls=%d out of range [1,16]", s->num_cls);
  const long long nt = static_cast<long long>(s->B) * s->T * s->H * s->W;
  if (nt > 0x7fffffffLL - 512) return fail(PARQ_ERR_SHAPE, "too many image tokens for 32-bit TMA coordinates");
  return PARQ_OK;
}

// activation x weight GEMM helper: A = [hi|lo] activations (a_base column offset selects a group),
// W = packed [hi|lo] weights; 2 terms when the weights are bf16-exact, else 3.
static void term_offsets(GemmParams& gp, int K, bool w_lo, int a_base) {
  gp.K = K;
  gp.const_operand = 2;          // activations x weights: the B operand holds constants
  gp.nterms = w_lo ? 3 : 2;
  gp.a_koff[0] = a_base;      gp.b_koff[0] = 0;
  gp.a_koff[1] = a_base + K;  gp.b_koff[1] = 0;
  gp.a_koff[2] = a_base;      gp.b_koff[2] = K;
}
static GemmEpilogue epilogue_none() {
  GemmEpilogue e;
  memset(&e, 0, sizeof(e));
  return e;
}

// K = tokens Wk^T + bk and V^T = Wv tokens^T + bv for `rows` consecutive tokens whose first one is token `tok0` of the batch
// (token index = clip * Nk + key), written into the tile-contiguous caches of the workspace.
static int kv_project_range(const ParqShape& s, cudaStream_t st, const void* tokens, long long rows, long long tok0, const uint8_t* pk,
                            const Packed& P, bool w_lo, uint8_t* ws, const Workspace& W) {
  const int C = s.C;
  const int Nk = s.T * s.H * s.W;
  // K: A = tokens (single bf16 term), B = Wk [hi|lo]
  GemmParams gk;
  memset(&gk, 0, sizeof(gk));
  gk.M = static_cast<int>(rows);  gk.N = C;  gk.K = C;
  // w_lo: weights that are not bf16-exact keep a low-order term (second pass over the tokens).  K and V^T are stored
  // in bf16, so that term is of the size of the storage rounding; PARQ_FLAG_KV_HI_ONLY drops it (measured: parity
  // error 5e-4 -> 9e-4 of the 1e-3 bar, K/V projection 3.7 -> 2.3 ms at config 2), the default keeps it.
  gk.nterms = w_lo ? 2 : 1;
  gk.const_operand = 2;
  gk.a_koff[0] = 0; gk.b_koff[0] = 0; gk.a_koff[1] = 0; gk.b_koff[1] = C;
  gk.ep = epilogue_none();
  gk.ep.bias = reinterpret_cast<const float*>(pk + P.ca_k_b);
  gk.ep.out_lp = ws + W.Kc;  gk.ep.ld_lp = C;
  gk.ep.kv_tiled = 1; gk.ep.kv_Nk = Nk; gk.ep.kv_ntile = W.ntile; gk.ep.kv_H = s.heads; gk.ep.kv_tok_offset = tok0;
  TRY(launch_gemm(st, tokens, rows, C, pk + P.ca_k, C, 2 * C, gk, TAG_KV_PROJ));
  // V^T: A = Wv [hi|lo], B = tokens
  GemmParams gv;
  memset(&gv, 0, sizeof(gv));
  gv.M = C;  gv.N = static_cast<int>(rows);  gv.K = C;
  gv.nterms = w_lo ? 2 : 1;
  gv.const_operand = 1;
  gv.a_koff[0] = 0; gv.b_koff[0] = 0; gv.a_koff[1] = C; gv.b_koff[1] = 0;
  gv.ep = epilogue_none();
  gv.ep.bias = reinterpret_cast<const float*>(pk + P.ca_v_b);
  gv.ep.bias_per_row = 1;
  gv.ep.out_lp = ws + W.Vt;  gv.ep.ld_lp = static_cast<long long>(W.ldv);
  gv.ep.kv_tiled = 2; gv.ep.kv_Nk = Nk; gv.ep.kv_ntile = W.ntile; gv.ep.kv_H = s.heads; gv.ep.kv_tok_offset = tok0;
  TRY(launch_gemm(st, pk + P.ca_v, C, 2 * C, tokens, rows, C, gv, TAG_KV_PROJ));
  return PARQ_OK;
}

static int kv_project(const ParqShape& s, cudaStream_t st, const void* tokens, const uint8_t* pk, const Packed& P, bool w_lo,
                      uint8_t* ws, const Workspace& W) {
  const long long Nt = static_cast<long long>(s.B) * s.T * s.H * s.W;
  const int Nk = s.T * s.H * s.W;
  // keys beyond Nk in a clip's last tile are never written: P is exactly 0 there, but 0 x garbage could be NaN
  if (Nk % attn::BKEY != 0) CUDA_TRY(cudaMemsetAsync(ws + W.Vt, 0, static_cast<size_t>(s.B) * W.ntile * attn::BKEY * s.C * 2, st));
  return kv_project_range(s, st, tokens, Nt, 0, pk, P, w_lo, ws, W);
}

}  // namespace parq

using namespace parq;

// =============================================================================== C ABI ========
extern "C" {

int parq_version(void) { return PARQ_ABI_VERSION; }
const char* parq_last_error(void) { return g_err; }

unsigned long long parq_kernel_launches(void) { return g_launches; }

int parq_profile_enable(uint32_t tag_mask, int max_records) {
  Profiler& p = g_prof;
  if (max_records > p.cap) {
    for (int i = 0; i < 2 * p.cap; ++i) cudaEventDestroy(p.ev[i]);
    delete[] p.ev;
    delete[] p.tag;
    p.ev = new cudaEvent_t[2 * max_records];
    p.tag = new int[max_records];
    for (int i = 0; i < 2 * max_records; ++i) CUDA_TRY(cudaEventCreate(&p.ev[i]));
    p.cap = max_records;
  }
  p.n = 0;
  p.mask = tag_mask;
  p.on = tag_mask != 0 && max_records > 0;
  return PARQ_OK;
}

int parq_profile_collect(float* ms_per_tag, int* launches_per_tag) {
  Profiler& p = g_prof;
  for (int t = 0; t < TAG_COUNT; ++t) { ms_per_tag[t] = 0.f; launches_per_tag[t] = 0; }
  for (int i = 0; i < p.n; ++i) {
    CUDA_TRY(cudaEventSynchronize(p.ev[2 * i + 1]));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, p.ev[2 * i], p.ev[2 * i + 1]));
    ms_per_tag[p.tag[i]] += ms;
    launches_per_tag[p.tag[i]] += 1;
  }
  const int dropped = (p.n >= p.cap) ? 1 : 0;
  p.n = 0;
  return dropped;
}

long long parq_workspace_offset(const ParqShape* shape, const char* name) {
  if (check_shape(shape) != PARQ_OK || name == nullptr) return -1;
  const Workspace W = workspace_layout(*shape, device_info().sms);
  const struct { const char* n; size_t off; } tab[] = {
      {"Kc", W.Kc}, {"Vt", W.Vt}, {"T_cl", W.T_cl}, {"ref_cur", W.ref_cur}, {"pe", W.pe}, {"qk_s", W.qk_s}, {"vt_s", W.vt_s},
      {"a_attn", W.a_attn}, {"y", W.y}, {"x1", W.x1}, {"x2", W.x2}, {"x3", W.x3}, {"q_c", W.q_c}, {"h1", W.h1}, {"h2", W.h2},
      {"ldv", W.ldv}, {"ldvs", W.ldvs}, {"kv_tiled", static_cast<size_t>(W.kv_tiled)}, {"ntile", static_cast<size_t>(W.ntile)}, {"cross_nsplit", static_cast<size_t>(W.cross.nsplit)},
      {"self_nsplit", static_cast<size_t>(W.self.nsplit)}};
  for (const auto& e : tab)
    if (strcmp(e.n, name) == 0) return static_cast<long long>(e.off);
  fail(PARQ_ERR_SHAPE, "unknown workspace region '%s'", name);
  return -1;
}

size_t parq_packed_bytes(const ParqShape* shape) {
  if (check_shape(shape) != PARQ_OK) return 0;
  return packed_layout(*shape).total;
}
size_t parq_workspace_bytes(const ParqShape* shape) {
  if (check_shape(shape) != PARQ_OK) return 0;
  return workspace_layout(*shape, device_info().sms).total;
}

int parq_pack_weights(const ParqShape* shape, const ParqWeightsF32* w, void* packed, size_t packed_bytes, void* stream) {
  TRY(check_shape(shape));
  if (w == nullptr || packed == nullptr) return fail(PARQ_ERR_SHAPE, "null pointer");
  const ParqShape& s = *shape;
  const Packed P = packed_layout(s);
  if (packed_bytes < P.total) return fail(PARQ_ERR_WORKSPACE, "packed buffer too small: need %zu, have %zu", P.total, packed_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  int* flag = reinterpret_cast<int*>(pk + P.lo_flag);
  CUDA_TRY(cudaMemsetAsync(flag, 0, 4, st));
  const int C = s.C, F = s.ffn;
  auto split = [&](const float* src, int rows, int K, float mul, size_t dst_off, size_t row0) -> int {
    if (src == nullptr) return fail(PARQ_ERR_SHAPE, "null weight pointer");
    const long long n = static_cast<long long>(rows) * K;
    split_weight_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
        src, rows, K, mul, reinterpret_cast<__nv_bfloat16*>(pk + dst_off) + row0 * 2 * K, flag);
    return PARQ_OK;
  };
  auto vec = [&](const float* src, int n, float mul, size_t dst_off, size_t el0) -> int {
    if (src == nullptr) return fail(PARQ_ERR_SHAPE, "null weight pointer");
    copy_scale_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, reinterpret_cast<float*>(pk + dst_off) + el0, n, mul);
    return PARQ_OK;
  };
  const float qs = 0.0625f;   // 1/sqrt(head_dim 256): exact power of two, folded into Wq / bq
  const long long CC = static_cast<long long>(C) * C;
  TRY(split(w->pe0_w, C, 384, 1.f, P.pe0, 0));
  TRY(split(w->pe2_w, C, C, 1.f, P.pe2, 0));
  TRY(split(w->sa_in_w, C, C, qs, P.sa_qk, 0));
  TRY(split(w->sa_in_w + CC, C, C, 1.f, P.sa_qk, C));
  TRY(split(w->sa_in_w + 2 * CC, C, C, 1.f, P.sa_v, 0));
  TRY(split(w->sa_out_w, C, C, 1.f, P.sa_out, 0));
  TRY(split(w->ca_in_w, C, C, qs, P.ca_q, 0));
  TRY(split(w->ca_in_w + CC, C, C, 1.f, P.ca_k, 0));
  TRY(split(w->ca_in_w + 2 * CC, C, C, 1.f, P.ca_v, 0));
  TRY(split(w->ca_out_w, C, C, 1.f, P.ca_out, 0));
  TRY(split(w->lin1_w, F, C, 1.f, P.lin1, 0));
  TRY(split(w->lin2_w, C, F, 1.f, P.lin2, 0));
  TRY(split(w->ctr0_w, C, C, 1.f, P.hd1, 0));
  TRY(split(w->rot0_w, C, C, 1.f, P.hd1, C));
  TRY(split(w->ctr4_w, C, C, 1.f, P.ctr4, 0));
  TRY(split(w->rot4_w, C, C, 1.f, P.rot4, 0));
  TRY(vec(w->pe0_b, C, 1.f, P.pe0_b, 0));
  TRY(vec(w->pe2_b, C, 1.f, P.pe2_b, 0));
  TRY(vec(w->sa_in_b, C, qs, P.sa_qk_b, 0));
  TRY(vec(w->sa_in_b + C, C, 1.f, P.sa_qk_b, C));
  TRY(vec(w->sa_in_b + 2 * C, C, 1.f, P.sa_v_b, 0));
  TRY(vec(w->sa_out_b, C, 1.f, P.sa_out_b, 0));
  TRY(vec(w->ca_in_b, C, qs, P.ca_q_b, 0));
  TRY(vec(w->ca_in_b + C, C, 1.f, P.ca_k_b, 0));
  TRY(vec(w->ca_in_b + 2 * C, C, 1.f, P.ca_v_b, 0));
  TRY(vec(w->ca_out_b, C, 1.f, P.ca_out_b, 0));
  TRY(vec(w->lin1_b, F, 1.f, P.lin1_b, 0));
  TRY(vec(w->lin2_b, C, 1.f, P.lin2_b, 0));
  TRY(vec(w->ln1_g, C, 1.f, P.ln1_g, 0));  TRY(vec(w->ln1_b, C, 1.f, P.ln1_b, 0));
  TRY(vec(w->ln2_g, C, 1.f, P.ln2_g, 0));  TRY(vec(w->ln2_b, C, 1.f, P.ln2_b, 0));
  TRY(vec(w->ln3_g, C, 1.f, P.ln3_g, 0));  TRY(vec(w->ln3_b, C, 1.f, P.ln3_b, 0));
  TRY(vec(w->ctr1_g, C, 1.f, P.ctr1_g, 0));  TRY(vec(w->ctr1_b, C, 1.f, P.ctr1_b, 0));
  TRY(vec(w->rot1_g, C, 1.f, P.rot1_g, 0));  TRY(vec(w->rot1_b, C, 1.f, P.rot1_b, 0));
  TRY(vec(w->ctr5_g, C, 1.f, P.ctr5_g, 0));  TRY(vec(w->ctr5_b, C, 1.f, P.ctr5_b, 0));
  TRY(vec(w->rot5_g, C, 1.f, P.rot5_g, 0));  TRY(vec(w->rot5_b, C, 1.f, P.rot5_b, 0));
  TRY(vec(w->cls_w, s.num_cls * C, 1.f, P.cls_w, 0));  TRY(vec(w->cls_b, s.num_cls, 1.f, P.cls_b, 0));
  TRY(vec(w->size_w, 3 * C, 1.f, P.size_w, 0));  TRY(vec(w->size_b, 3, 1.f, P.size_b, 0));
  TRY(vec(w->ctr8_w, 3 * C, 1.f, P.ctr8_w, 0));  TRY(vec(w->ctr8_b, 3, 1.f, P.ctr8_b, 0));
  TRY(vec(w->rot8_w, 6 * C, 1.f, P.rot8_w, 0));  TRY(vec(w->rot8_b, 6, 1.f, P.rot8_b, 0));
  TRY(vec(w->mean_size, s.num_cls * 3, 1.f, P.mean_size, 0));
  TRY(vec(w->dim_t, 128, 1.f, P.dim_t, 0));
  CUDA_TRY(cudaGetLastError());
  int w_lo = 0;
  CUDA_TRY(cudaMemcpyAsync(&w_lo, flag, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return w_lo != 0 ? 1 : 0;
}

int parq_pose_chain(const float* T_cp, const float* T_wp, const float* T_wl, float* T_cl, int B, int T, void* stream) {
  TRY(require_sm100());
  if (!T_cp || !T_wp || !T_wl || !T_cl || B < 1 || T < 1) return fail(PARQ_ERR_SHAPE, "bad pose_chain arguments");
  {
    ProfScope ps(TAG_ROWWISE, static_cast<cudaStream_t>(stream));
    pose_chain_kernel<<<(B * T + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(T_cp, T_wp, T_wl, T_cl, B, T);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}

static int launch_sample(cudaStream_t st, const SampleParams& sp) {
  if ((reinterpret_cast<uintptr_t>(sp.tokens) & 15) != 0 || (reinterpret_cast<uintptr_t>(sp.tokens_lo) & 15) != 0)
    return fail(PARQ_ERR_SHAPE, "token planes must be 16-byte aligned");
  const int R = sp.B * sp.Nq;
  const size_t smem = static_cast<size_t>(SAMPLE_QPB) * sp.T * sizeof(ViewTap);
  {
    ProfScope ps(TAG_SAMPLE, st);
    if (sp.tokens_lo != nullptr)
      launch_k(project_sample_kernel<true>, dim3(R / SAMPLE_QPB), dim3(sp.C / 8), smem, st, sp);
    else
      launch_k(project_sample_kernel<false>, dim3(R / SAMPLE_QPB), dim3(sp.C / 8), smem, st, sp);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}

static void fill_sample_params(SampleParams& sp, const ParqShape& s) {
  memset(&sp, 0, sizeof(sp));
  sp.B = s.B; sp.T = s.T; sp.H = s.H; sp.W = s.W; sp.C = s.C; sp.Nq = s.Nq;
  for (int i = 0; i < 3; ++i) {
    // (hi - lo) is evaluated in double from the config floats, then enters the fp32 op as a scalar
    sp.span[i] = static_cast<float>(static_cast<double>(s.scale[2 * i + 1]) - static_cast<double>(s.scale[2 * i]));
    sp.lo[i] = s.scale[2 * i];
  }
}

int parq_project_sample(const ParqShape* shape, const void* tokens_bf16, const void* tokens_lo_bf16, const float* ref, const float* T_cl,
                        const float* camera, float* features, float* center_im, uint8_t* center_valid, float* coord_pos, void* stream) {
  TRY(require_sm100());
  TRY(check_shape(shape));
  if (!tokens_bf16 || !ref || !T_cl || !camera || !features) return fail(PARQ_ERR_SHAPE, "null pointer");
  SampleParams sp;
  fill_sample_params(sp, *shape);
  sp.tokens = static_cast<const __nv_bfloat16*>(tokens_bf16);
  sp.tokens_lo = static_cast<const __nv_bfloat16*>(tokens_lo_bf16);
  sp.ref = ref; sp.T_cl = T_cl; sp.camera = camera;
  sp.feat = features; sp.center_im = center_im; sp.valid = center_valid; sp.coord_pos = coord_pos;
  return launch_sample(static_cast<cudaStream_t>(stream), sp);
}

// fp32 tokens -> exact bf16 pair: hi = bf16(x), lo = bf16(x - hi)  (x = hi + lo to 16 mantissa bits)
__global__ void split_tokens_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo, long long n4) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = src[i];
    const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
    hi[i] = make_uint2(h0, h1);
    lo[i] = make_uint2(pack_bf16x2(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xFFFF0000u)),
                       pack_bf16x2(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xFFFF0000u)));
  }
}

int parq_split_tokens(const float* tokens_f32, void* hi_bf16, void* lo_bf16, long long n, void* stream) {
  TRY(require_sm100());
  if (!tokens_f32 || !hi_bf16 || !lo_bf16 || n < 0 || n % 4 != 0) return fail(PARQ_ERR_SHAPE, "bad split_tokens arguments (n must be a multiple of 4)");
  if (((reinterpret_cast<uintptr_t>(tokens_f32) | reinterpret_cast<uintptr_t>(hi_bf16) | reinterpret_cast<uintptr_t>(lo_bf16)) & 15) != 0)
    return fail(PARQ_ERR_SHAPE, "split_tokens needs 16-byte aligned buffers");
  if (n == 0) return PARQ_OK;
  const long long n4 = n / 4;
  const long long want = (n4 + 255) / 256;
  const int blocks = static_cast<int>(want < 148LL * 16 ? want : 148LL * 16);
  {
    ProfScope ps(TAG_ROWWISE, static_cast<cudaStream_t>(stream));
    split_tokens_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(tokens_f32), static_cast<uint2*>(hi_bf16),
                                                                              static_cast<uint2*>(lo_bf16), n4);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}

// ---- f-3: FPN upsample + concat ---------------------------------------------------------------------------------
static int fpn_concat_impl(const void* l0, const void* l1, const void* l2, const void* l3, int in_bf16, const int32_t* level_hw, int BT,
                           int channels_per_level, int target_level, void* out_nchw, int out_bf16, void* stream);

int parq_fpn_concat(const float* l0, const float* l1, const float* l2, const float* l3, const int32_t* level_hw, int BT, int channels_per_level,
                    int target_level, float* out_nchw, void* stream) {
  return fpn_concat_impl(l0, l1, l2, l3, 0, level_hw, BT, channels_per_level, target_level, out_nchw, 0, stream);
}
int parq_fpn_concat_bf16(const void* l0, const void* l1, const void* l2, const void* l3, const int32_t* level_hw, int BT, int channels_per_level,
                         int target_level, float* out_nchw, void* stream) {
  return fpn_concat_impl(l0, l1, l2, l3, 1, level_hw, BT, channels_per_level, target_level, out_nchw, 0, stream);
}
int parq_fpn_concat_ex(const void* l0, const void* l1, const void* l2, const void* l3, int levels_bf16, const int32_t* level_hw, int BT,
                       int channels_per_level, int target_level, void* out_nchw, int out_bf16, void* stream) {
  return fpn_concat_impl(l0, l1, l2, l3, levels_bf16 ? 1 : 0, level_hw, BT, channels_per_level, target_level, out_nchw, out_bf16 ? 1 : 0, stream);
}

static int fpn_concat_impl(const void* l0, const void* l1, const void* l2, const void* l3, int in_bf16, const int32_t* level_hw, int BT,
                           int channels_per_level, int target_level, void* out_nchw, int out_bf16, void* stream) {
  TRY(require_sm100());
  if (!l0 || !l1 || !l2 || !l3 || !level_hw || !out_nchw) return fail(PARQ_ERR_SHAPE, "null pointer");
  if (BT < 1 || channels_per_level < 1 || target_level < 0 || target_level > 3) return fail(PARQ_ERR_SHAPE, "bad fpn_concat arguments");
  FpnParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.level[0] = l0; fp.level[1] = l1; fp.level[2] = l2; fp.level[3] = l3;
  fp.in_bf16 = in_bf16;
  for (int l = 0; l < 4; ++l) {
    fp.h[l] = level_hw[2 * l];
    fp.w[l] = level_hw[2 * l + 1];
    if (fp.h[l] < 1 || fp.w[l] < 1) return fail(PARQ_ERR_SHAPE, "non-positive level size");
  }
  fp.BT = BT; fp.Cl = channels_per_level; fp.H = fp.h[target_level]; fp.W = fp.w[target_level];
  fp.out = out_nchw;
  const long long planes = static_cast<long long>(BT) * 4 * channels_per_level;
  if (planes > 0x7fffffffLL) return fail(PARQ_ERR_SHAPE, "too many feature planes");
  if (channels_per_level % FPN_CH != 0) return fail(PARQ_ERR_SHAPE, "channels_per_level=%d must be a multiple of %d", channels_per_level, FPN_CH);
  {
    fp.plane0 = 0;
    const dim3 grid(static_cast<unsigned>(planes / FPN_CH));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ProfScope ps(TAG_ROWWISE, st);
    if (in_bf16 && out_bf16) fpn_concat_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(fp);
    else if (in_bf16) fpn_concat_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(fp);
    else if (out_bf16) fpn_concat_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(fp);
    else fpn_concat_kernel<float, float><<<grid, 256, 0, st>>>(fp);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}

// ---- f-1: AddRayPE + tokeniser ---------------------------------------------------------------------------------
struct RayPacked { size_t w0, w2, b0, b2, lo_flag, total; };
static RayPacked raype_packed_layout(int C, int n) {
  RayPacked r;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  r.w0 = take(static_cast<size_t>(C) * 2 * 3 * n * 2);
  r.w2 = take(static_cast<size_t>(C) * 2 * C * 2);
  r.b0 = take(static_cast<size_t>(C) * 4);
  r.b2 = take(static_cast<size_t>(C) * 4);
  r.lo_flag = take(4);
  r.total = off;
  return r;
}
struct RayWorkspace { size_t aff, feat, hidden, total; };
static RayWorkspace raype_workspace_layout(int B, int T, int H, int W, int C, int n) {
  RayWorkspace r;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t ntok = static_cast<size_t>(B) * T * H * W;
  r.aff = take(static_cast<size_t>(B) * T * 12 * 4);
  r.feat = take(ntok * 2 * 3 * n * 2);
  r.hidden = take(ntok * 2 * C * 2);      // [hi | lo] when PARQ_RAYPE_SPLIT_HIDDEN, else hi only
  r.total = off;
  return r;
}
static int raype_check(int B, int T, int H, int W, int C, int n) {
  if (B < 1 || T < 1 || H < 1 || W < 1) return fail(PARQ_ERR_SHAPE, "non-positive dimension");
  if (C < 64 || C % 64 != 0) return fail(PARQ_ERR_SHAPE, "C=%d must be a positive multiple of 64", C);
  if (n < 1 || (3 * n) % 64 != 0) return fail(PARQ_ERR_SHAPE, "3*num_samples=%d must be a multiple of 64", 3 * n);
  if (static_cast<long long>(B) * T * H * W > 0x7fffffffLL - 512) return fail(PARQ_ERR_SHAPE, "too many pixels for 32-bit TMA coordinates");
  return PARQ_OK;
}

size_t parq_raype_packed_bytes(int C, int num_samples) {
  if (raype_check(1, 1, 1, 1, C, num_samples) != PARQ_OK) return 0;
  return raype_packed_layout(C, num_samples).total;
}
size_t parq_raype_workspace_bytes(int B, int T, int H, int W, int C, int num_samples) {
  if (raype_check(B, T, H, W, C, num_samples) != PARQ_OK) return 0;
  return raype_workspace_layout(B, T, H, W, C, num_samples).total;
}

int parq_raype_pack_weights(int C, int num_samples, const float* w0, const float* b0, const float* w2, const float* b2, void* packed,
                            size_t packed_bytes, void* stream) {
  TRY(raype_check(1, 1, 1, 1, C, num_samples));
  if (!w0 || !b0 || !w2 || !b2 || !packed) return fail(PARQ_ERR_SHAPE, "null pointer");
  const RayPacked P = raype_packed_layout(C, num_samples);
  if (packed_bytes < P.total) return fail(PARQ_ERR_WORKSPACE, "packed buffer too small: need %zu, have %zu", P.total, packed_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  int* flag = reinterpret_cast<int*>(pk + P.lo_flag);
  CUDA_TRY(cudaMemsetAsync(flag, 0, 4, st));
  const int F = 3 * num_samples;
  split_weight_kernel<<<(C * F + 255) / 256, 256, 0, st>>>(w0, C, F, 1.f, reinterpret_cast<__nv_bfloat16*>(pk + P.w0), flag);
  split_weight_kernel<<<(C * C + 255) / 256, 256, 0, st>>>(w2, C, C, 1.f, reinterpret_cast<__nv_bfloat16*>(pk + P.w2), flag);
  copy_scale_kernel<<<(C + 255) / 256, 256, 0, st>>>(b0, reinterpret_cast<float*>(pk + P.b0), C, 1.f);
  copy_scale_kernel<<<(C + 255) / 256, 256, 0, st>>>(b2, reinterpret_cast<float*>(pk + P.b2), C, 1.f);
  CUDA_TRY(cudaGetLastError());
  int w_lo = 0;
  CUDA_TRY(cudaMemcpyAsync(&w_lo, flag, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return w_lo != 0 ? 1 : 0;
}

int parq_raype_forward(int B, int T, int H, int W, int C, int num_samples, const float* feat_nchw, const float* camera,
                       const float* T_cp, const float* T_wp, const float* T_wl, const float* depth_planes, const float* ray_points_scale,
                       const void* packed, void* workspace, size_t workspace_bytes, void* tokens_bf16, float* encoding_nchw,
                       uint32_t flags, void* stream) {
  TRY(require_sm100());
  TRY(raype_check(B, T, H, W, C, num_samples));
  if (!camera || !T_cp || !T_wp || !T_wl || !depth_planes || !ray_points_scale || !packed || !workspace)
    return fail(PARQ_ERR_SHAPE, "null pointer");
  if (!tokens_bf16 && !encoding_nchw) return fail(PARQ_ERR_SHAPE, "need a tokens or an encoding output");
  if (tokens_bf16 && !feat_nchw) return fail(PARQ_ERR_SHAPE, "tokens = features + encoding needs the feature maps");
  const RayWorkspace Wk = raype_workspace_layout(B, T, H, W, C, num_samples);
  if (workspace_bytes < Wk.total) return fail(PARQ_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", Wk.total, workspace_bytes);
  const RayPacked P = raype_packed_layout(C, num_samples);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const bool w_lo = (flags & PARQ_FLAG_WEIGHT_LO) != 0;
  const int F = 3 * num_samples;
  const long long ntok = static_cast<long long>(B) * T * H * W;
  {
    ProfScope ps(TAG_ROWWISE, st, 2);
    raype_affine_kernel<<<(B * T + 127) / 128, 128, 0, st>>>(T_cp, T_wp, T_wl, reinterpret_cast<float*>(ws + Wk.aff), B, T);
    RayFeatParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.camera = camera; rp.aff = reinterpret_cast<const float*>(ws + Wk.aff); rp.depth = depth_planes;
    rp.out = reinterpret_cast<__nv_bfloat16*>(ws + Wk.feat);
    rp.BT = B * T; rp.H = H; rp.W = W; rp.n = num_samples;
    for (int i = 0; i < 3; ++i) {
      rp.lo[i] = ray_points_scale[2 * i];
      rp.span[i] = static_cast<float>(static_cast<double>(ray_points_scale[2 * i + 1]) - static_cast<double>(ray_points_scale[2 * i]));
    }
    const long long nthreads = ntok * num_samples;
    ray_features_kernel<<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, st>>>(rp);
  }
  CUDA_TRY(cudaGetLastError());
  // layer 0: relu(W0 f + b0), inputs as an exact [hi|lo] split (positions are precision-critical), hidden kept in bf16
  GemmParams g;
  memset(&g, 0, sizeof(g));
  g.M = static_cast<int>(ntok); g.N = C;
  term_offsets(g, F, w_lo, 0);
  g.ep = epilogue_none();
  g.ep.bias = reinterpret_cast<const float*>(pk + P.b0); g.ep.relu = 1;
  const bool split_hidden = (flags & PARQ_RAYPE_SPLIT_HIDDEN) != 0;
  g.ep.out_lp = ws + Wk.hidden; g.ep.ld_lp = split_hidden ? 2 * C : C;
  g.ep.lp_lo_off = split_hidden ? C : 0;
  TRY(launch_gemm(st, ws + Wk.feat, ntok, 2 * F, pk + P.w0, C, 2 * F, g));
  // layer 2: W2 h + b2 (+ channels-first features) -> channels-last bf16 tokens and / or the channels-first fp32 encoding
  memset(&g, 0, sizeof(g));
  g.M = static_cast<int>(ntok); g.N = C;
  if (split_hidden) {
    term_offsets(g, C, w_lo, 0);                      // h_hi W + h_lo W (+ h_hi W_lo)
  } else {
    g.K = C;
    g.nterms = w_lo ? 2 : 1;
    g.const_operand = 2;
    g.a_koff[0] = 0; g.b_koff[0] = 0; g.a_koff[1] = 0; g.b_koff[1] = C;
  }
  g.ep = epilogue_none();
  g.ep.bias = reinterpret_cast<const float*>(pk + P.b2);
  g.ep.nchw_HW = H * W;
  g.ep.nchw_out = encoding_nchw;
  if (tokens_bf16) {
    g.ep.nchw_add = feat_nchw;
    g.ep.nchw_add_bf16 = (flags & PARQ_RAYPE_FEAT_BF16) ? 1 : 0;
    g.ep.out_lp = tokens_bf16; g.ep.ld_lp = C;
  }
  TRY(launch_gemm(st, ws + Wk.hidden, ntok, split_hidden ? 2 * C : C, pk + P.w2, C, 2 * C, g));
  return PARQ_OK;
}

int parq_parse_pred(const float* center, const float* size, const float* ortho6d, const float* prob, int B, int K, int num_cls,
                    const float* track_scale, double overlap_threshold, uint32_t mode, uint8_t* pred_mask, uint8_t* nms_mask,
                    float* scores, int32_t* labels, float* obbs, void* stream) {
  TRY(require_sm100());
  if (!center || !size || !ortho6d || !prob || !pred_mask || !track_scale) return fail(PARQ_ERR_SHAPE, "null pointer");
  if (B < 1 || K < 1 || K > 1024 || num_cls < 2 || num_cls > 64) return fail(PARQ_ERR_SHAPE, "parse_pred supports 1..1024 boxes per clip");
  ParsePredParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.center = center; pp.size = size; pp.ortho6d = ortho6d; pp.prob = prob;
  pp.B = B; pp.K = K; pp.num_cls = num_cls;
  pp.background = num_cls - 1;
  pp.same_class = (mode & PARQ_NMS_SAME_CLASS) ? 1 : 0;
  pp.apply_track_scale = (mode & PARQ_NMS_NO_TRACK_SCALE) ? 0 : 1;
  for (int i = 0; i < 6; ++i) pp.track_scale[i] = track_scale[i];
  pp.threshold = overlap_threshold;
  pp.pred_mask = pred_mask; pp.nms_mask = nms_mask; pp.scores = scores; pp.labels = labels; pp.obbs = obbs;
  int P = 32;
  while (P < K) P <<= 1;
  const size_t smem = parse_pred_smem(P);
  OPT_IN_SMEM(parse_pred_kernel, smem);
  {
    ProfScope ps(TAG_ROWWISE, static_cast<cudaStream_t>(stream));
    launch_k(parse_pred_kernel, dim3(B), dim3((K + 31) / 32 * 32), smem, static_cast<cudaStream_t>(stream), pp, P);
  }
  CUDA_TRY(cudaGetLastError());
  return PARQ_OK;
}

int parq_gemm_bf16(const void* A, int64_t a_rows, int64_t a_cols, const void* Bw, int64_t b_rows, int64_t b_cols, int M, int N,
                   int K, int nterms, const int32_t* a_koff, const int32_t* b_koff, const float* bias, int bias_per_row, int relu,
                   float* out_f32, int64_t ld_f32, void* out_lp, int64_t ld_lp, int lp_fp16, int64_t lp_lo_off, void* stream) {
  TRY(require_sm100());
  if (!A || !Bw || M < 1 || N < 1 || !a_koff || !b_koff) return fail(PARQ_ERR_SHAPE, "bad GEMM arguments");
  GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.M = M; gp.N = N; gp.K = K; gp.nterms = nterms;
  for (int t = 0; t < nterms && t < 3; ++t) { gp.a_koff[t] = a_koff[t]; gp.b_koff[t] = b_koff[t]; }
  gp.ep = epilogue_none();
  gp.ep.bias = bias; gp.ep.bias_per_row = bias_per_row; gp.ep.relu = relu;
  gp.ep.out_f32 = out_f32; gp.ep.ld_f32 = ld_f32;
  gp.ep.out_lp = out_lp; gp.ep.ld_lp = ld_lp; gp.ep.lp_fp16 = lp_fp16; gp.ep.lp_lo_off = lp_lo_off;
  return launch_gemm(static_cast<cudaStream_t>(stream), A, a_rows, a_cols, Bw, b_rows, b_cols, gp);
}

/* Unit-test entry of the chained kernel: y = LayerNorm(x_split W1^T + b1 + resid) * gamma + beta, then z = relu(y W2^T + b2),
 * as ONE launch of two chained stages (N1 = 1024 columns for the LayerNorm stage). */
int parq_chain_ln_linear(const void* a_split, const void* w1_split, const float* b1, const float* resid_cm, const float* gamma, const float* beta,
                         const void* w2_split, const float* b2, int M, int K1, int N2, int w_lo, float* y_f32, void* y_split, void* z_split,
                         void* stream) {
  TRY(require_sm100());
  if (!a_split || !w1_split || !resid_cm || !gamma || !beta || !w2_split || !y_f32 || !y_split || !z_split)
    return fail(PARQ_ERR_SHAPE, "null pointer");
  const int N1 = 1024;
  ChainBuilder cb(M);
  ChainStage S = chain_stage(N1, K1, CH_EP_LN, b1);
  S.resid_cm = resid_cm; S.gamma = gamma; S.beta = beta;
  S.out_f32 = y_f32; S.a_out = static_cast<__nv_bfloat16*>(y_split);
  TRY(chain_add(cb, a_split, 2 * static_cast<uint64_t>(K1), w1_split, w_lo != 0, S));
  S = chain_stage(N2, N1, CH_EP_SPLIT, b2);
  S.relu = 1; S.a_out = static_cast<__nv_bfloat16*>(z_split);
  TRY(chain_add(cb, y_split, 2 * N1, w2_split, w_lo != 0, S));
  return launch_chain(static_cast<cudaStream_t>(stream), cb);
}

size_t parq_attention_scratch_bytes(int B, int H, int Nq, int Nk) {
  const int ntiles = (Nk + attn::BKEY - 1) / attn::BKEY;
  const size_t a = attn_scratch_bytes(B, H, Nq, ntiles < 32 ? ntiles : 32), b = streamk_scratch_bytes(B, H, Nq, Nk, device_info().sms);
  return a > b ? a : b;
}

int parq_attention(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* Vt, int64_t ldv, int B, int H, int Nq, int Nk,
                   int fp16, void* scratch, size_t scratch_bytes, void* out_split, int force_nsplit, void* stream) {
  TRY(require_sm100());
  if (!Q || !K || !Vt || !scratch || !out_split || B < 1 || H < 1 || Nk < 1) return fail(PARQ_ERR_SHAPE, "bad attention arguments");
  return launch_attention(static_cast<cudaStream_t>(stream), Q, ldq, K, ldk, Vt, ldv, B, H, Nq, Nk, fp16 != 0, scratch, scratch_bytes,
                          static_cast<__nv_bfloat16*>(out_split), force_nsplit);
}

int parq_kv_project(const ParqShape* shape, const void* tokens_bf16, const void* packed, void* workspace, size_t workspace_bytes,
                    uint32_t flags, void* stream) {
  TRY(require_sm100());
  TRY(check_shape(shape));
  if (!tokens_bf16 || !packed || !workspace) return fail(PARQ_ERR_SHAPE, "null pointer");
  const Workspace W = workspace_layout(*shape, device_info().sms);
  if (workspace_bytes < W.total) return fail(PARQ_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", W.total, workspace_bytes);
  const Packed P = packed_layout(*shape);
  return kv_project(*shape, static_cast<cudaStream_t>(stream), tokens_bf16, static_cast<const uint8_t*>(packed), P,
                    (flags & PARQ_FLAG_WEIGHT_LO) != 0 && !(flags & PARQ_FLAG_KV_HI_ONLY), static_cast<uint8_t*>(workspace), W);
}

/* instrumentation: global-timer stamp of every following kernel launch at the moment its stream dependency resolved (see ptx.cuh);
 * buf = device memory of `capacity` uint64 (slot 0 = number of stamps so far, zeroed here), NULL switches it off */
int parq_trace(void* buf, int capacity) {
  unsigned long long* b = static_cast<unsigned long long*>(buf);
  unsigned int cap = buf != nullptr && capacity > 1 ? static_cast<unsigned int>(capacity) : 0u;
  if (cap == 0) b = nullptr;
  if (b != nullptr) CUDA_TRY(cudaMemset(b, 0, sizeof(unsigned long long)));
  CUDA_TRY(cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap)));
  CUDA_TRY(cudaMemcpyToSymbol(g_trace_buf, &b, sizeof(b)));
  CUDA_TRY(cudaDeviceSynchronize());
  g_trace_chain_slot = b != nullptr ? 0 : -1;
  return PARQ_OK;
}

/* instrumentation: clock64 stamps of CTA 0 of the next <= 64 chain launches into buf (64 slots each); NULL switches it off */
int parq_chain_debug(void* buf) {
  g_chain_dbg = static_cast<long long*>(buf);
  g_chain_dbg_launch = 0;
  return PARQ_OK;
}

int parq_kv_project_views(const ParqShape* shape, const void* view_tokens_bf16, int slot0, int n_views, const void* packed, void* workspace,
                          size_t workspace_bytes, uint32_t flags, void* stream) {
  TRY(require_sm100());
  TRY(check_shape(shape));
  if (!view_tokens_bf16 || !packed || !workspace) return fail(PARQ_ERR_SHAPE, "null pointer");
  const ParqShape& s = *shape;
  if (slot0 < 0 || n_views < 1 || slot0 + n_views > s.T) return fail(PARQ_ERR_SHAPE, "view slots [%d, %d) outside the %d views of the window", slot0, slot0 + n_views, s.T);
  const long long hw = static_cast<long long>(s.H) * s.W;
  if (hw % 32 != 0) return fail(PARQ_ERR_SHAPE, "per-view K / V^T updates need H*W %% 32 == 0 (H*W = %lld)", hw);
  const Workspace W = workspace_layout(s, device_info().sms);
  if (workspace_bytes < W.total) return fail(PARQ_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", W.total, workspace_bytes);
  const Packed P = packed_layout(s);
  const bool w_lo = (flags & PARQ_FLAG_WEIGHT_LO) != 0 && !(flags & PARQ_FLAG_KV_HI_ONLY);
  const long long rows = hw * n_views, Nk = hw * s.T;
  for (int b = 0; b < s.B; ++b) {
    const uint8_t* tok = static_cast<const uint8_t*>(view_tokens_bf16) + static_cast<size_t>(b) * rows * s.C * 2;
    TRY(kv_project_range(s, static_cast<cudaStream_t>(stream), tok, rows, b * Nk + slot0 * hw, static_cast<const uint8_t*>(packed), P, w_lo,
                         static_cast<uint8_t*>(workspace), W));
  }
  return PARQ_OK;
}

// (un-chained launch path) pe = W2 relu(W1 posemb + b1) + b2 as two GEMM launches on `st`
static int launch_pe_mlp(cudaStream_t st, const ParqShape& s, const Workspace& W, const Packed& P, uint8_t* ws, const uint8_t* pk, bool w_lo) {
  const int C = s.C, R = s.B * s.Nq;
  GemmParams g; memset(&g, 0, sizeof(g));
  g.M = R; g.N = C; term_offsets(g, 384, w_lo, 0);
  g.ep = epilogue_none(); g.ep.bias = reinterpret_cast<const float*>(pk + P.pe0_b); g.ep.relu = 1;
  g.ep.out_lp = ws + W.a_peh; g.ep.ld_lp = 2 * C; g.ep.lp_lo_off = C;
  TRY(launch_gemm(st, ws + W.a_pos, R, 768, pk + P.pe0, C, 768, g));
  memset(&g, 0, sizeof(g));
  g.M = R; g.N = C; term_offsets(g, C, w_lo, 0);
  g.ep = epilogue_none(); g.ep.bias = reinterpret_cast<const float*>(pk + P.pe2_b);
  g.ep.out_f32 = reinterpret_cast<float*>(ws + W.pe); g.ep.ld_f32 = C;
  return launch_gemm(st, ws + W.a_peh, R, 2 * C, pk + P.pe2, C, 2 * C, g);
}
// (un-chained launch path) V^T = Wv x^T + bv of the self-attention: weights are the A operand, activations the B operand
static int launch_sa_v(cudaStream_t st, const ParqShape& s, const Workspace& W, const Packed& P, uint8_t* ws, const uint8_t* pk, bool w_lo,
                       bool hi_only) {
  const int C = s.C, R = s.B * s.Nq;
  GemmParams g; memset(&g, 0, sizeof(g));
  g.M = C; g.N = R; g.K = C; g.nterms = w_lo ? 3 : (hi_only ? 1 : 2); g.const_operand = 1;
  g.a_koff[0] = 0; g.b_koff[0] = 0; g.a_koff[1] = 0; g.b_koff[1] = C; g.a_koff[2] = C; g.b_koff[2] = 0;
  g.ep = epilogue_none(); g.ep.bias = reinterpret_cast<const float*>(pk + P.sa_v_b); g.ep.bias_per_row = 1;
  g.ep.out_lp = ws + W.vt_s; g.ep.ld_lp = static_cast<long long>(W.ldvs); g.ep.lp_fp16 = 1;
  return launch_gemm(st, pk + P.sa_v, C, 2 * C, ws + W.a_x, R, 2 * C, g);
}

int parq_decoder_forward(const ParqShape* shape, const void* tokens_bf16, const void* tokens_lo_bf16, const float* camera, const float* T_cp,
                         const float* T_wp, const float* T_wl, const float* ref0, const float* forced_refs, const void* packed, void* workspace,
                         size_t workspace_bytes, const ParqOutputs* out, uint32_t flags, void* stream) {
  TRY(require_sm100());
  TRY(check_shape(shape));
  if (!tokens_bf16 || !camera || !T_cp || !T_wp || !T_wl || !packed || !workspace || !out) return fail(PARQ_ERR_SHAPE, "null pointer");
  if (!ref0 && !forced_refs) return fail(PARQ_ERR_SHAPE, "need ref0 or forced_refs");
  if (!out->pred_logits || !out->center_unnormalized || !out->size_unnormalized || !out->ortho6d || !out->sem_cls_prob || !out->coord_pos)
    return fail(PARQ_ERR_SHAPE, "required output pointer is null");
  const ParqShape& s = *shape;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Workspace W = workspace_layout(s, device_info().sms);
  if (workspace_bytes < W.total) return fail(PARQ_ERR_WORKSPACE, "workspace too small: need %zu, have %zu", W.total, workspace_bytes);
  const Packed P = packed_layout(s);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const bool w_lo = (flags & PARQ_FLAG_WEIGHT_LO) != 0;
  const PdlScope pdl((flags & PARQ_FLAG_NO_PDL) == 0);
  const int C = s.C, F = s.ffn, R = s.B * s.Nq;
  // the row-local linears of an iteration as three chained launches (chain_tc.cuh) instead of ten GEMMs + three LayerNorms
  // Low-order activation term of the three GEMMs whose output is rounded to 16 bits (self-attention Q|K and V^T in fp16,
  // cross-attention Q in bf16): PARQ_HI_ONLY (environment) overrides the default for the ablation of DESIGN.md
  // bits: 0 sa_qk, 1 sa_v, 2 ca_q (both launch paths), 3 pe0, 4 pe2, 5 sa_out, 6 ca_out, 7 lin1, 8 lin2, 9 hd1, 10 hd2 (chained
  // path only); bf16-exact weights
  const int hi_only = getenv("PARQ_HI_ONLY") ? g_hi_only
                      : ((flags & PARQ_FLAG_HI_ONLY_SET) ? static_cast<int>((flags & PARQ_FLAG_HI_ONLY_MASK) >> PARQ_FLAG_HI_ONLY_SHIFT) & 0x7FF
                                                         : HI_ONLY_DEFAULT);
  auto hio = [&](int bit) { return (!w_lo && ((hi_only >> bit) & 1)) ? 1 : 0; };
  // It pays when the one-wave GEMMs it replaces fill the machine (R = 4096 rows at config 2: -0.46 ms per step); with a few
  // hundred rows (one clip: 2-4 clusters) the separate launches are faster (measured at config 5: 2.7 vs 3.5 ms per window).
  const bool chained = !(flags & PARQ_FLAG_NO_CHAIN) && !g_no_chain && chain_cols_ok(C) && chain_cols_ok(2 * C) && chain_cols_ok(F) && R % chain::BM == 0 &&
                       (R >= g_chain_min_rows || (flags & PARQ_FLAG_FORCE_CHAIN));
  auto F32 = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto BF = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto PF = [&](size_t off) { return reinterpret_cast<const float*>(pk + off); };

  // piece flags of the fused stream-K merge (attn3_tc.cuh): every launch leaves them zero, this covers the first use of a workspace
  // (opt-in: in the power-capped steady state of config 2 the step time is the same with either merge)
  const bool fused_merge = (flags & PARQ_FLAG_FUSED_MERGE) != 0 || g_fused_merge;
  if (fused_merge) CUDA_TRY(cudaMemsetAsync(ws + W.sk_flags, 0, SK_FLAG_WORDS * sizeof(uint32_t), st));
  // K0: pose chain
  { ProfScope ps(TAG_ROWWISE, st); launch_k(pose_chain_kernel, dim3((s.B * s.T + 127) / 128), dim3(128), 0, st, T_cp, T_wp, T_wl, F32(W.T_cl), s.B, s.T); }
  CUDA_TRY(cudaGetLastError());
  // K4: hoisted K / V^T projection of the image tokens (iteration invariant)
  if (!(flags & PARQ_FLAG_SKIP_KV)) TRY(kv_project(s, st, tokens_bf16, pk, P, w_lo && !(flags & PARQ_FLAG_KV_HI_ONLY), ws, W));

  SampleParams sp;
  fill_sample_params(sp, s);
  HeadsParams hp;
  memset(&hp, 0, sizeof(hp));
  for (int i = 0; i < 3; ++i) { hp.span[i] = sp.span[i]; hp.lo[i] = sp.lo[i]; }

  for (int it = 0; it < s.iters; ++it) {
    const float* ref = forced_refs ? forced_refs + static_cast<size_t>(it) * R * 3 : (it == 0 ? ref0 : F32(W.ref_cur));
    float* x3 = out->decoder_out ? out->decoder_out + static_cast<size_t>(it) * R * C : F32(W.x3);
    const int Nk = s.T * s.H * s.W;
    // sinusoidal embedding of the reference points (input of the reference-point MLP); launched BEFORE the sampling kernel so
    // that the reference points are two launches old when the sampler projects them in its pre-wait prologue
    // (one clip, free-running iterations > 0: the heads kernel of the previous iteration already wrote it next to ref_next --
    // one dependent launch less where launches are the cost; with thousands of rows the 12 sin / cos per lane and row make the
    // heads kernel 19 us slower, more than this 9 us kernel and its launch)
    const bool posemb_folded = forced_refs == nullptr && R < g_chain_min_rows;
    if (!posemb_folded || it == 0) {
      ProfScope ps(TAG_ROWWISE, st);
      launch_k(posemb_kernel, dim3((R * 384 + 255) / 256), dim3(256), 0, st, ref, PF(P.dim_t), BF(W.a_pos), R);
    }
    CUDA_TRY(cudaGetLastError());
    // K1: projection + multi-view bilinear gather -> the [hi|lo] split of the sampled features (the query content x)
    sp.tokens = static_cast<const __nv_bfloat16*>(tokens_bf16);
    sp.tokens_lo = static_cast<const __nv_bfloat16*>(tokens_lo_bf16);
    sp.ref = ref; sp.T_cl = F32(W.T_cl); sp.camera = camera;
    sp.ref_is_fresh = (posemb_folded && it > 0) ? 1 : 0;      // then the heads kernel is the direct predecessor (no posemb launch between)
    sp.feat = out->features ? out->features + static_cast<size_t>(it) * R * C : nullptr;
    sp.a_x = BF(W.a_x);
    sp.center_im = out->center_im ? out->center_im + static_cast<size_t>(it) * s.B * s.T * s.Nq * 2 : nullptr;
    sp.valid = out->center_valid ? out->center_valid + static_cast<size_t>(it) * s.B * s.T * s.Nq : nullptr;
    sp.coord_pos = nullptr;
    // un-chained launch path (one clip): the reference-point MLP needs only the sinusoidal embedding, not the sampled features --
    // it runs on a side stream next to the sampling kernel; V^T of the self-attention (needs only x) next to the Q|K projection
    SideStreams* sd = (!chained && !(flags & PARQ_FLAG_NO_FORK) && !g_no_fork) ? side_streams(st) : nullptr;
    if (sd != nullptr) {
      TRY(side_fork(st, sd, 0));
      TRY(launch_pe_mlp(sd->s[0], s, W, P, ws, pk, w_lo));
    }
    TRY(launch_sample(st, sp));
    if (chained) {
      // ---- chain P: pe = W2 relu(W1 posemb + b1) + b2 (+ x -> split(x + pe)) -> self-attention Q|K projection
      {
        ChainBuilder cb(R);
        // stage order: pe0 first (short, K = 384), then V (independent of it: its MMAs run while pe0's epilogue drains), then
        // pe2 (reads stage 0, long since complete: its MMAs cover V's epilogue), then Q|K
        ChainStage S = chain_stage(C, 384, CH_EP_SPLIT, PF(P.pe0_b));
        S.relu = 1; S.a_out = BF(W.a_peh); S.hi_only = hio(3);
        TRY(chain_add(cb, ws + W.a_pos, 768, pk + P.pe0, w_lo, S));
        if (!g_no_chain_v) {
          // V = x Wv^T + bv of the self-attention, stored transposed (V^T, K-major for P.V) in fp16: it needs only x
          S = chain_stage(C, C, CH_EP_LP_T, PF(P.sa_v_b));
          S.out_lp = ws + W.vt_s; S.ld_lp = static_cast<long long>(W.ldvs); S.lp_fp16 = 1; S.dep = -1;
          S.hi_only = hio(1);
          TRY(chain_add(cb, ws + W.a_x, 2 * C, pk + P.sa_v, w_lo, S));
        }
        // (column-major private streams of the chained path: W.pe = pe, W.y = x, W.x1 = x1, W.x2 = x2 -- see chain_tc.cuh)
        S = chain_stage(C, C, CH_EP_F32, PF(P.pe2_b));
        S.out_cm = F32(W.pe); S.add_split = BF(W.a_x); S.out_sum_split = BF(W.a_xpe); S.add_cm_out = F32(W.y); S.dep = 0;
        S.hi_only = hio(4);
        TRY(chain_add(cb, ws + W.a_peh, 2 * C, pk + P.pe2, w_lo, S));
        S = chain_stage(2 * C, C, CH_EP_LP, PF(P.sa_qk_b));
        S.out_lp = ws + W.qk_s; S.ld_lp = 2 * C; S.lp_fp16 = 1; S.hi_only = hio(0);
        TRY(chain_add(cb, ws + W.a_xpe, 2 * C, pk + P.sa_qk, w_lo, S));
        TRY(launch_chain(st, cb));
      }
      // (A/B switch PARQ_NO_CHAIN_V) V^T = Wv x^T + bv of the self-attention as its own GEMM: weights are the A operand, activations the B operand
      if (g_no_chain_v) {
        GemmParams g; memset(&g, 0, sizeof(g));
        g.M = C; g.N = R; g.K = C; g.nterms = w_lo ? 3 : (hio(1) ? 1 : 2); g.const_operand = 1;
        g.a_koff[0] = 0; g.b_koff[0] = 0; g.a_koff[1] = 0; g.b_koff[1] = C; g.a_koff[2] = C; g.b_koff[2] = 0;
        g.ep = epilogue_none(); g.ep.bias = PF(P.sa_v_b); g.ep.bias_per_row = 1;
        g.ep.out_lp = ws + W.vt_s; g.ep.ld_lp = static_cast<long long>(W.ldvs); g.ep.lp_fp16 = 1;
        TRY(launch_gemm(st, pk + P.sa_v, C, 2 * C, ws + W.a_x, R, 2 * C, g));
      }
      TRY(launch_attention(st, ws + W.qk_s, 2 * C, ws + W.qk_s + static_cast<size_t>(C) * 2, 2 * C, ws + W.vt_s, W.ldvs, s.B, s.heads,
                           s.Nq, s.Nq, true, ws + W.scratch, W.scratch_bytes, BF(W.a_attn), W.self.nsplit));
      // ---- chain A: self-attention out-projection + residual + LN1 (+pe) -> cross-attention Q projection
      {
        ChainBuilder cb(R);
        ChainStage S = chain_stage(C, C, CH_EP_LN, PF(P.sa_out_b));
        S.resid_cm = F32(W.y); S.gamma = PF(P.ln1_g); S.beta = PF(P.ln1_b); S.pe_cm = F32(W.pe);
        S.out_cm = F32(W.x1); S.a_out_pe = BF(W.a_x1pe); S.hi_only = hio(5);
        TRY(chain_add(cb, ws + W.a_attn, 2 * C, pk + P.sa_out, w_lo, S));
        S = chain_stage(C, C, CH_EP_LP, PF(P.ca_q_b));
        S.out_lp = ws + W.q_c; S.ld_lp = C; S.hi_only = hio(2);
        TRY(chain_add(cb, ws + W.a_x1pe, 2 * C, pk + P.ca_q, w_lo, S));
        TRY(launch_chain(st, cb));
      }
      TRY(launch_attention(st, ws + W.q_c, C, ws + W.Kc, C, ws + W.Vt, W.ldv, s.B, s.heads, s.Nq, Nk, false, ws + W.scratch,
                           W.scratch_bytes, BF(W.a_attn), /*library's choice: stream-K or the planned split*/ 0, /*kv_const=*/true,
                           W.kv_tiled != 0, fused_merge ? reinterpret_cast<uint32_t*>(ws + W.sk_flags) : nullptr));
      // ---- chain B: cross-attention out-projection + residual + LN2 -> FFN -> + residual + LN3 -> first head layer
      {
        ChainBuilder cb(R);
        ChainStage S = chain_stage(C, C, CH_EP_LN, PF(P.ca_out_b));
        S.resid_cm = F32(W.x1); S.gamma = PF(P.ln2_g); S.beta = PF(P.ln2_b);
        S.out_cm = F32(W.x2); S.a_out = BF(W.a_x2); S.hi_only = hio(6);
        TRY(chain_add(cb, ws + W.a_attn, 2 * C, pk + P.ca_out, w_lo, S));
        S = chain_stage(F, C, CH_EP_SPLIT, PF(P.lin1_b));
        S.relu = 1; S.a_out = BF(W.a_ffn); S.hi_only = hio(7);
        TRY(chain_add(cb, ws + W.a_x2, 2 * C, pk + P.lin1, w_lo, S));
        S = chain_stage(C, F, CH_EP_LN, PF(P.lin2_b));
        S.resid_cm = F32(W.x2); S.gamma = PF(P.ln3_g); S.beta = PF(P.ln3_b);
        S.out_f32 = x3; S.a_out = BF(W.a_x3); S.hi_only = hio(8);
        TRY(chain_add(cb, ws + W.a_ffn, 2 * F, pk + P.lin2, w_lo, S));
        S = chain_stage(2 * C, C, CH_EP_F32, nullptr);
        S.out_f32 = F32(W.h1); S.gn_out = reinterpret_cast<double2*>(ws + W.gn1); S.gn_stride = GN_SLOTS_PER_MTILE; S.hi_only = hio(9);
        TRY(chain_add(cb, ws + W.a_x3, 2 * C, pk + P.hd1, w_lo, S));
        TRY(launch_chain(st, cb));
      }
    } else {
    // K2: reference-point positional feature  pe = W2 relu(W1 posemb + b1) + b2, then the split of x + pe, the query / key input
    // of the self-attention (transformer_parq.py:372)
    {
      if (sd != nullptr) {
        TRY(side_fork(st, sd, 1));                // after the sampler: V^T = Wv x^T + bv on side stream 1
        TRY(launch_sa_v(sd->s[1], s, W, P, ws, pk, w_lo, hio(1) != 0));
        TRY(side_join(st, sd, 0));                // pe is there
      } else {
        TRY(launch_pe_mlp(st, s, W, P, ws, pk, w_lo));
      }
      { ProfScope ps(TAG_ROWWISE, st); launch_k(split_sum_kernel, dim3((R * (C / 8) + 255) / 256), dim3(256), 0, st, BF(W.a_x), F32(W.pe), BF(W.a_xpe), R, C); }
      CUDA_TRY(cudaGetLastError());
    }
    // K3: self-attention among the queries (fp16 operands), out-projection, residual + LN1
    {
      GemmParams g; memset(&g, 0, sizeof(g));
      g.M = R; g.N = 2 * C; term_offsets(g, C, w_lo, 0);
      if (hio(0)) g.nterms = 1;
      g.ep = epilogue_none(); g.ep.bias = PF(P.sa_qk_b);
      g.ep.out_lp = ws + W.qk_s; g.ep.ld_lp = 2 * C; g.ep.lp_fp16 = 1;
      TRY(launch_gemm(st, ws + W.a_xpe, R, 2 * C, pk + P.sa_qk, 2 * C, 2 * C, g));
      if (sd != nullptr) TRY(side_join(st, sd, 1));
      else TRY(launch_sa_v(st, s, W, P, ws, pk, w_lo, hio(1) != 0));
      TRY(launch_attention(st, ws + W.qk_s, 2 * C, ws + W.qk_s + static_cast<size_t>(C) * 2, 2 * C, ws + W.vt_s, W.ldvs, s.B, s.heads,
                           s.Nq, s.Nq, true, ws + W.scratch, W.scratch_bytes, BF(W.a_attn), W.self.nsplit));
      memset(&g, 0, sizeof(g));
      g.M = R; g.N = C; term_offsets(g, C, w_lo, 0);
      g.ep = epilogue_none(); g.ep.bias = PF(P.sa_out_b);
      g.ep.out_f32 = F32(W.y); g.ep.ld_f32 = C;
      TRY(launch_gemm(st, ws + W.a_attn, R, 2 * C, pk + P.sa_out, C, 2 * C, g));
      // the residual x is read back from its [hi|lo] split (the only form in which the sampled features are stored)
      { ProfScope ps(TAG_ROWWISE, st); launch_k(add_ln_kernel<1024>, dim3((R + 3) / 4), dim3(128), 0, st, static_cast<const float*>(nullptr), BF(W.a_x), F32(W.y),
                                               PF(P.ln1_g), PF(P.ln1_b), F32(W.pe), F32(W.x1), nullptr, BF(W.a_x1pe), R); }
      CUDA_TRY(cudaGetLastError());
    }
    // K5: cross-attention over all image tokens (bf16 operands), out-projection, residual + LN2
    {
      GemmParams g; memset(&g, 0, sizeof(g));
      g.M = R; g.N = C; term_offsets(g, C, w_lo, 0);
      if (hio(2)) g.nterms = 1;
      g.ep = epilogue_none(); g.ep.bias = PF(P.ca_q_b);
      g.ep.out_lp = ws + W.q_c; g.ep.ld_lp = C;
      TRY(launch_gemm(st, ws + W.a_x1pe, R, 2 * C, pk + P.ca_q, C, 2 * C, g));
      TRY(launch_attention(st, ws + W.q_c, C, ws + W.Kc, C, ws + W.Vt, W.ldv, s.B, s.heads, s.Nq, Nk, false, ws + W.scratch,
                           W.scratch_bytes, BF(W.a_attn), /*library's choice: stream-K or the planned split*/ 0, /*kv_const=*/true,
                           W.kv_tiled != 0, fused_merge ? reinterpret_cast<uint32_t*>(ws + W.sk_flags) : nullptr));
      memset(&g, 0, sizeof(g));
      g.M = R; g.N = C; term_offsets(g, C, w_lo, 0);
      g.ep = epilogue_none(); g.ep.bias = PF(P.ca_out_b);
      g.ep.out_f32 = F32(W.y); g.ep.ld_f32 = C;
      TRY(launch_gemm(st, ws + W.a_attn, R, 2 * C, pk + P.ca_out, C, 2 * C, g));
      { ProfScope ps(TAG_ROWWISE, st); launch_k(add_ln_kernel<1024>, dim3((R + 3) / 4), dim3(128), 0, st, F32(W.x1), static_cast<const __nv_bfloat16*>(nullptr), F32(W.y), PF(P.ln2_g), PF(P.ln2_b), nullptr, F32(W.x2),
                                               BF(W.a_x2), nullptr, R); }
      CUDA_TRY(cudaGetLastError());
    }
    // K6: FFN, residual + LN3
    {
      GemmParams g; memset(&g, 0, sizeof(g));
      g.M = R; g.N = F; term_offsets(g, C, w_lo, 0);
      g.ep = epilogue_none(); g.ep.bias = PF(P.lin1_b); g.ep.relu = 1;
      g.ep.out_lp = ws + W.a_ffn; g.ep.ld_lp = 2 * F; g.ep.lp_lo_off = F;
      TRY(launch_gemm(st, ws + W.a_x2, R, 2 * C, pk + P.lin1, F, 2 * C, g));
      memset(&g, 0, sizeof(g));
      g.M = R; g.N = C; term_offsets(g, F, w_lo, 0);
      g.ep = epilogue_none(); g.ep.bias = PF(P.lin2_b);
      g.ep.out_f32 = F32(W.y); g.ep.ld_f32 = C;
      TRY(launch_gemm(st, ws + W.a_ffn, R, 2 * F, pk + P.lin2, C, 2 * F, g));
      { ProfScope ps(TAG_ROWWISE, st); launch_k(add_ln_kernel<1024>, dim3((R + 3) / 4), dim3(128), 0, st, F32(W.x2), static_cast<const __nv_bfloat16*>(nullptr), F32(W.y), PF(P.ln3_g), PF(P.ln3_b), nullptr, x3, BF(W.a_x3), nullptr, R); }
      CUDA_TRY(cudaGetLastError());
    }
    // first hidden layer of the centre / rotation heads (one GEMM, N = 2C) with the GroupNorm tile sums in its epilogue
    {
      GemmParams g; memset(&g, 0, sizeof(g));
      g.M = R; g.N = 2 * C; term_offsets(g, C, w_lo, 0);
      g.ep = epilogue_none();
      g.ep.out_f32 = F32(W.h1); g.ep.ld_f32 = 2 * C;
      g.ep.gn_out = reinterpret_cast<double2*>(ws + W.gn1); g.ep.gn_stride = GN_SLOTS_PER_MTILE;
      TRY(launch_gemm(st, ws + W.a_x3, R, 2 * C, pk + P.hd1, 2 * C, 2 * C, g));
    }
    }   // !chained
    // K7: heads (two hidden layers with per-clip GroupNorm) + box update
    {
      GemmParams g;
      { ProfScope ps(TAG_ROWWISE, st);
        // 8 rows per block from a few thousand rows on (the statistics prologue once per block), one row per block below
        if (R >= 2048)
          launch_k(gn_apply_kernel<8>, dim3((R + 7) / 8), dim3(256), 0, st, F32(W.h1), 2 * C, C, s.Nq, 2, reinterpret_cast<const double2*>(ws + W.gn1),
                   PF(P.ctr1_g), PF(P.ctr1_b), PF(P.rot1_g), PF(P.rot1_b), BF(W.a_h1), R);
        else
          launch_k(gn_apply_kernel<1>, dim3(R), dim3(256), 0, st, F32(W.h1), 2 * C, C, s.Nq, 2, reinterpret_cast<const double2*>(ws + W.gn1),
                   PF(P.ctr1_g), PF(P.ctr1_b), PF(P.rot1_g), PF(P.rot1_b), BF(W.a_h1), R);
      }
      CUDA_TRY(cudaGetLastError());
      // second hidden layer of the centre and rotation heads as ONE launch: the two weight matrices are adjacent in the
      // packed buffer (one B operand of 2C rows), output columns >= C read the rotation half of a_h1
      memset(&g, 0, sizeof(g));
      g.M = R; g.N = 2 * C; term_offsets(g, C, w_lo, 0);
      if (chained && hio(10)) g.nterms = 1;
      g.a_split_n = C; g.a_split_off = 2 * C;
      g.ep = epilogue_none();
      g.ep.out_f32 = F32(W.h2); g.ep.ld_f32 = 2 * C;
      g.ep.gn_out = reinterpret_cast<double2*>(ws + W.gn2); g.ep.gn_stride = GN_SLOTS_PER_MTILE;
      TRY(launch_gemm(st, ws + W.a_h1, R, 4 * C, pk + P.ctr4, 2 * C, 2 * C, g));
      hp.x = x3; hp.h2 = F32(W.h2); hp.partial = reinterpret_cast<const double2*>(ws + W.gn2);
      hp.gamma_c = PF(P.ctr5_g); hp.beta_c = PF(P.ctr5_b); hp.gamma_r = PF(P.rot5_g); hp.beta_r = PF(P.rot5_b);
      hp.w_cls = PF(P.cls_w); hp.b_cls = PF(P.cls_b); hp.w_size = PF(P.size_w); hp.b_size = PF(P.size_b);
      hp.w_c3 = PF(P.ctr8_w); hp.b_c3 = PF(P.ctr8_b); hp.w_r3 = PF(P.rot8_w); hp.b_r3 = PF(P.rot8_b);
      hp.ref = ref; hp.mean_size = PF(P.mean_size);
      const size_t o = static_cast<size_t>(it) * R;
      hp.logits = out->pred_logits + o * s.num_cls; hp.prob = out->sem_cls_prob + o * s.num_cls;
      hp.center = out->center_unnormalized + o * 3; hp.size = out->size_unnormalized + o * 3;
      hp.ortho6d = out->ortho6d + o * 6; hp.coord_pos = out->coord_pos + o * 3;
      hp.rot = out->rotation ? out->rotation + o * 9 : nullptr;
      hp.ref_next = F32(W.ref_cur);
      hp.posemb_next = (posemb_folded && it + 1 < s.iters) ? BF(W.a_pos) : nullptr;
      hp.dim_t = PF(P.dim_t);
      hp.R = R; hp.Nq = s.Nq; hp.C = C; hp.num_cls = s.num_cls;
      { ProfScope ps(TAG_ROWWISE, st); {
        const int rpb = (R + device_info().sms - 1) / device_info().sms;     // rows per block: one block per SM
        const size_t hsm = heads_smem_bytes<1024>();
        if (R <= 1024) {
          // a few rows (one clip): weights in registers, channels split over the warps of a block (rowwise.cuh)
          const dim3 grid((R + HEADS_SMALL_ROWS - 1) / HEADS_SMALL_ROWS);
          if (hp.posemb_next != nullptr) launch_k(heads_final_small_kernel<1024, true>, grid, dim3(256), 0, st, hp);
          else launch_k(heads_final_small_kernel<1024, false>, grid, dim3(256), 0, st, hp);
        } else if (hp.posemb_next != nullptr) {
          OPT_IN_SMEM((heads_final_kernel<1024, true>), hsm);
          launch_k(heads_final_kernel<1024, true>, dim3((R + rpb - 1) / rpb), dim3(HEADS_THREADS), hsm, st, hp, rpb);
        } else {
          OPT_IN_SMEM((heads_final_kernel<1024, false>), hsm);
          launch_k(heads_final_kernel<1024, false>, dim3((R + rpb - 1) / rpb), dim3(HEADS_THREADS), hsm, st, hp, rpb);
        }
      } }
      CUDA_TRY(cudaGetLastError());
    }
  }
  return PARQ_OK;
}

}  // extern "C"
