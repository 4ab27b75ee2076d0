"""ctypes binding of libparq_b200.so (the C ABI declared in include/parq_b200.h).

There is no fallback: if the shared library is missing or a call fails, this
module raises.  The library is looked up in-tree (parq_b200/libparq_b200.so,
built by parq_b200.build / __graft_entry__.build()).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libparq_b200.so")

PARQ_FLAG_SKIP_KV = 1
PARQ_FLAG_WEIGHT_LO = 2
PARQ_FLAG_NO_PDL = 4
PARQ_FLAG_KV_HI_ONLY = 16
PARQ_FLAG_NO_CHAIN = 64
PARQ_FLAG_FORCE_CHAIN = 128
PARQ_FLAG_FUSED_MERGE = 512
PARQ_FLAG_NO_FORK = 1024
PARQ_FLAG_HI_ONLY_SHIFT = 16
PARQ_FLAG_HI_ONLY_SET = 0x08000000
PARQ_RAYPE_SPLIT_HIDDEN = 8
PARQ_RAYPE_FEAT_BF16 = 256
PARQ_NMS_SAME_CLASS = 1
PARQ_NMS_NO_TRACK_SCALE = 2

EXPORTS = [
    "parq_version", "parq_last_error", "parq_packed_bytes", "parq_workspace_bytes", "parq_pack_weights",
    "parq_pose_chain", "parq_split_tokens", "parq_project_sample", "parq_kv_project", "parq_kv_project_views", "parq_chain_debug", "parq_trace", "parq_decoder_forward",
    "parq_gemm_bf16", "parq_chain_ln_linear", "parq_attention_scratch_bytes", "parq_attention",
    "parq_parse_pred", "parq_fpn_concat", "parq_fpn_concat_bf16", "parq_fpn_concat_ex", "parq_raype_packed_bytes", "parq_raype_workspace_bytes", "parq_raype_pack_weights",
    "parq_raype_forward", "parq_kernel_launches", "parq_profile_enable", "parq_profile_collect", "parq_workspace_offset",
]
PROFILE_TAGS = ("kv_proj", "project_sample", "gemm", "self_attn", "cross_attn", "combine", "rowwise")


class ParqShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "T", "H", "W", "C", "Nq", "heads", "ffn", "iters", "num_cls")] + \
               [("scale", C.c_float * 6)]


WEIGHT_FIELDS = [
    "pe0_w", "pe0_b", "pe2_w", "pe2_b",
    "sa_in_w", "sa_in_b", "sa_out_w", "sa_out_b",
    "ca_in_w", "ca_in_b", "ca_out_w", "ca_out_b",
    "lin1_w", "lin1_b", "lin2_w", "lin2_b",
    "ln1_g", "ln1_b", "ln2_g", "ln2_b", "ln3_g", "ln3_b",
    "cls_w", "cls_b",
    "ctr0_w", "ctr1_g", "ctr1_b", "ctr4_w", "ctr5_g", "ctr5_b", "ctr8_w", "ctr8_b",
    "size_w", "size_b",
    "rot0_w", "rot1_g", "rot1_b", "rot4_w", "rot5_g", "rot5_b", "rot8_w", "rot8_b",
    "mean_size", "dim_t",
]


class ParqWeightsF32(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in WEIGHT_FIELDS]


OUTPUT_FIELDS = ["pred_logits", "center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob", "coord_pos",
                 "rotation", "center_im", "center_valid", "features", "decoder_out"]


class ParqOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUTPUT_FIELDS]


class ParqError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ParqError("libparq_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU or PyTorch fallback for the decoder hot path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz, u32, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_uint32, C.c_void_p
    lib.parq_version.restype = C.c_int
    lib.parq_last_error.restype = C.c_char_p
    lib.parq_packed_bytes.restype = sz
    lib.parq_packed_bytes.argtypes = [C.POINTER(ParqShape)]
    lib.parq_workspace_bytes.restype = sz
    lib.parq_workspace_bytes.argtypes = [C.POINTER(ParqShape)]
    lib.parq_pack_weights.restype = C.c_int
    lib.parq_pack_weights.argtypes = [C.POINTER(ParqShape), C.POINTER(ParqWeightsF32), vp, sz, vp]
    lib.parq_pose_chain.restype = C.c_int
    lib.parq_pose_chain.argtypes = [f32p, f32p, f32p, f32p, i32, i32, vp]
    lib.parq_project_sample.restype = C.c_int
    lib.parq_project_sample.argtypes = [C.POINTER(ParqShape), vp, vp, f32p, f32p, f32p, f32p, f32p, vp, f32p, vp]
    lib.parq_split_tokens.restype = C.c_int
    lib.parq_split_tokens.argtypes = [f32p, vp, vp, C.c_longlong, vp]
    lib.parq_kv_project.restype = C.c_int
    lib.parq_kv_project.argtypes = [C.POINTER(ParqShape), vp, vp, vp, sz, u32, vp]
    lib.parq_kv_project_views.restype = C.c_int
    lib.parq_kv_project_views.argtypes = [C.POINTER(ParqShape), vp, i32, i32, vp, vp, sz, u32, vp]
    lib.parq_chain_debug.restype = C.c_int
    lib.parq_chain_debug.argtypes = [vp]
    lib.parq_trace.restype = C.c_int
    lib.parq_trace.argtypes = [vp, i32]
    lib.parq_decoder_forward.restype = C.c_int
    lib.parq_decoder_forward.argtypes = [C.POINTER(ParqShape), vp, vp, f32p, f32p, f32p, f32p, f32p, f32p, vp, vp, sz,
                                         C.POINTER(ParqOutputs), u32, vp]
    lib.parq_gemm_bf16.restype = C.c_int
    lib.parq_gemm_bf16.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                   f32p, i32, i32, f32p, i64, vp, i64, i32, i64, vp]
    lib.parq_chain_ln_linear.restype = C.c_int
    lib.parq_chain_ln_linear.argtypes = [vp, vp, f32p, f32p, f32p, f32p, vp, f32p, i32, i32, i32, i32, f32p, vp, vp, vp]
    lib.parq_attention_scratch_bytes.restype = sz
    lib.parq_attention_scratch_bytes.argtypes = [i32, i32, i32, i32]
    lib.parq_attention.restype = C.c_int
    lib.parq_attention.argtypes = [vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, i32, vp, sz, vp, i32, vp]
    lib.parq_parse_pred.restype = C.c_int
    lib.parq_parse_pred.argtypes = [f32p, f32p, f32p, f32p, i32, i32, i32, C.POINTER(C.c_float), C.c_double, u32, vp, vp, f32p, vp,
                                    f32p, vp]
    lib.parq_fpn_concat.restype = C.c_int
    lib.parq_fpn_concat.argtypes = [f32p, f32p, f32p, f32p, C.POINTER(C.c_int32), i32, i32, i32, f32p, vp]
    lib.parq_fpn_concat_bf16.restype = C.c_int
    lib.parq_fpn_concat_bf16.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int32), i32, i32, i32, f32p, vp]
    lib.parq_fpn_concat_ex.restype = C.c_int
    lib.parq_fpn_concat_ex.argtypes = [vp, vp, vp, vp, i32, C.POINTER(C.c_int32), i32, i32, i32, vp, i32, vp]
    lib.parq_raype_packed_bytes.restype = sz
    lib.parq_raype_packed_bytes.argtypes = [i32, i32]
    lib.parq_raype_workspace_bytes.restype = sz
    lib.parq_raype_workspace_bytes.argtypes = [i32] * 6
    lib.parq_raype_pack_weights.restype = C.c_int
    lib.parq_raype_pack_weights.argtypes = [i32, i32, f32p, f32p, f32p, f32p, vp, sz, vp]
    lib.parq_raype_forward.restype = C.c_int
    lib.parq_raype_forward.argtypes = [i32] * 6 + [f32p] * 6 + [C.POINTER(C.c_float), vp, vp, sz, vp, f32p, u32, vp]
    lib.parq_workspace_offset.restype = C.c_longlong
    lib.parq_workspace_offset.argtypes = [C.POINTER(ParqShape), C.c_char_p]
    lib.parq_kernel_launches.restype = C.c_ulonglong
    lib.parq_profile_enable.restype = C.c_int
    lib.parq_profile_enable.argtypes = [u32, i32]
    lib.parq_profile_collect.restype = C.c_int
    lib.parq_profile_collect.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int)]
    _lib = lib
    return lib


def profile_enable(tags, max_records=4096):
    """Arm event profiling for the named tags (see PROFILE_TAGS); empty disarms."""
    mask = 0
    for t in tags:
        mask |= 1 << PROFILE_TAGS.index(t)
    check(load().parq_profile_enable(mask, max_records), "parq_profile_enable")


def profile_collect():
    """{tag: (total_ms, launches)} for the launches recorded since the last call."""
    ms = (C.c_float * 8)()
    n = (C.c_int * 8)()
    dropped = check(load().parq_profile_collect(ms, n), "parq_profile_collect")
    out = {t: (float(ms[i]), int(n[i])) for i, t in enumerate(PROFILE_TAGS)}
    out["_dropped"] = bool(dropped)
    return out


def check(rc, what):
    if rc < 0:
        raise ParqError("%s failed (%d): %s" % (what, rc, load().parq_last_error().decode("utf-8", "replace")))
    return rc
