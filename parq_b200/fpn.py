"""Host side of the "next" row f-3: the FPN upsample + concat of the reference's ``ResnetFPN.forward``
(/root/reference/model/resnet_fpn.py:73-90) on the device in one pass, plus the camera rescale that goes with it."""
import ctypes as C

import torch

from . import _lib
from .decoder import _ptr, _stream
from .wrappers import Camera, raw


def fpn_concat(features, layer=0, out_dtype=torch.float32):
    """``features``: the backbone's dict {"0": (N,Cl,h0,w0), "1": ..., "2": ..., "3": ...} fp32 (or all-bf16) CUDA tensors
    (torchvision ``resnet_fpn_backbone`` output; "pool" is ignored).  Returns (N, 4*Cl, h_layer, w_layer) fp32 =
    ``torch.cat([F.interpolate(features[str(l)], features[str(layer)].shape[-2:], mode="bilinear") for l in range(4)], 1)``.
    ``out_dtype=torch.bfloat16`` writes the result in bf16 -- half the bytes of this memory-bound pass and of its read-back in
    ``AddRayPEB200.tokens`` (which accepts it directly); the fp32 default is the reference's ``all_features``."""
    lv = [features[str(l)] for l in range(4)]
    if lv[0].device.type != "cuda":
        raise NotImplementedError("fpn_concat needs CUDA tensors on an sm_100 device (no CPU fallback)")
    bf16 = all(t.dtype == torch.bfloat16 for t in lv)
    lv = [t.detach().contiguous() if bf16 else t.detach().float().contiguous() for t in lv]
    N, Cl = lv[0].shape[:2]
    if any(t.shape[0] != N or t.shape[1] != Cl or t.dim() != 4 for t in lv):
        raise ValueError("pyramid levels must be (N, Cl, h, w) with the same N and Cl")
    hw = (C.c_int32 * 8)(*[int(v) for t in lv for v in t.shape[-2:]])
    H, W = lv[int(layer)].shape[-2:]
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("out_dtype must be float32 or bfloat16")
    out = torch.empty(N, 4 * Cl, H, W, dtype=out_dtype, device=lv[0].device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().parq_fpn_concat_ex(_ptr(lv[0]), _ptr(lv[1]), _ptr(lv[2]), _ptr(lv[3]), int(bf16), hw, N, Cl, int(layer), _ptr(out),
                                                   int(out_dtype == torch.bfloat16), _stream()), "parq_fpn_concat_ex")
    return out


def camera_feature(camera, layer=0):
    """``camera.scale(1 / 2**(layer+2))`` (resnet_fpn.py:88-90, utils/wrappers.py:478-488) on the raw (…,6) camera tensor."""
    return Camera(raw(camera)).scale(1.0 / (2 ** (int(layer) + 2)))
