"""Host side of the "next" row f-1: a drop-in for the reference's ``AddRayPE`` fused with the tokeniser.

``AddRayPEB200`` keeps the reference module's constructor arguments, parameter names (``encoder.0/2.weight/bias``, the
``add_ray_pe.`` checkpoint prefix of utils/weight_convert.py) and ``forward`` contract
(/root/reference/model/ray_positional_encoding.py:29-139): it returns the (B,T,C,H,W) fp32 encoding, so
``parq_lightning.py:72`` works unchanged.  ``tokens()`` is the fused producer: features + encoding written straight
into the channels-last bf16 token tensor the decoder consumes (replaces parq_lightning.py:72-85 -- no fp32 encoding
tensor, no einops transpose copy, no cast).  All arithmetic runs in libparq_b200.so; there is no fallback.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from .decoder import _ptr, _stream
from .wrappers import raw


class AddRayPEB200(nn.Module):
    def __init__(self, dim_out, ray_points_scale=(-2, 2, -1.5, 0, 0.25, 4.25), num_samples=64, min_depth=0.25, max_depth=5.25):
        super().__init__()
        self.dim_out, self.num_samples = dim_out, num_samples
        self.ray_points_scale = [float(x) for x in ray_points_scale]
        self.min_depth, self.max_depth = min_depth, max_depth
        self.encoder = nn.Sequential(nn.Linear(3 * num_samples, dim_out), nn.ReLU(), nn.Linear(dim_out, dim_out))
        self._packed = None
        self._key = None
        self._ws = None

    def _depth_planes(self, device):
        # the reference's own torch expression (utils/encoding_utils.py:82-88), fp32
        ramp = torch.linspace(0, 1, self.num_samples)
        mn, mx = torch.tensor([self.min_depth])[0], torch.tensor([self.max_depth])[0]
        return torch.exp(torch.log(mn) + torch.log(mx / mn) * ramp).to(device).contiguous()

    def _prepare(self, device):
        lib = _lib.load()
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._key != key:
            f32 = lambda t: t.detach().to(device, torch.float32).contiguous()
            w0, b0, w2, b2 = f32(self.encoder[0].weight), f32(self.encoder[0].bias), f32(self.encoder[2].weight), f32(self.encoder[2].bias)
            n = lib.parq_raype_packed_bytes(self.dim_out, self.num_samples)
            if n == 0:
                raise _lib.ParqError("parq_raype_packed_bytes: " + lib.parq_last_error().decode())
            packed = torch.empty(n, dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                rc = _lib.check(lib.parq_raype_pack_weights(self.dim_out, self.num_samples, _ptr(w0), _ptr(b0), _ptr(w2), _ptr(b2),
                                                            _ptr(packed), n, _stream()), "parq_raype_pack_weights")
            self._packed, self._flags, self._depth, self._key = packed, (_lib.PARQ_FLAG_WEIGHT_LO if rc else 0), self._depth_planes(device), key
        return lib

    def _run(self, images_feat, camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, want_tokens, want_encoding, out=None):
        if self.training:
            raise NotImplementedError("AddRayPEB200 is inference-only: call .eval() (no training fallback exists)")
        if images_feat.device.type != "cuda":
            raise NotImplementedError("AddRayPEB200 needs CUDA tensors on an sm_100 device (no CPU fallback)")
        dev = images_feat.device
        lib = self._prepare(dev)
        B, T, Cc, H, W = images_feat.shape
        if Cc != self.dim_out:
            raise ValueError("images_feat has %d channels, module was built for %d" % (Cc, self.dim_out))
        f32 = lambda t: raw(t).detach().to(dev, torch.float32).contiguous()
        # bf16 features (fpn_concat(..., out_dtype=torch.bfloat16)) are consumed as they are by the fused token producer
        feat_bf16 = want_tokens and images_feat.dtype == torch.bfloat16
        feat = images_feat.detach().contiguous() if feat_bf16 else f32(images_feat)
        cam, Tcp, Twp, Twl = f32(camera), f32(T_camera_pseudoCam), f32(T_world_pseudoCam), f32(T_world_local)
        nws = lib.parq_raype_workspace_bytes(B, T, H, W, Cc, self.num_samples)
        if self._ws is None or self._ws.numel() < nws or self._ws.device != dev:
            self._ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        tokens = None
        if want_tokens:
            tokens = out if out is not None else torch.empty(B, T * H * W, Cc, dtype=torch.bfloat16, device=dev)
            if tuple(tokens.shape) != (B, T * H * W, Cc) or tokens.dtype != torch.bfloat16 or not tokens.is_contiguous() or tokens.device != dev:
                raise ValueError("out must be a contiguous (B, T*H*W, C) bf16 tensor on the features' device")
        enc = torch.empty(B, T, Cc, H, W, dtype=torch.float32, device=dev) if want_encoding else None
        scale = (C.c_float * 6)(*self.ray_points_scale)
        with torch.cuda.device(dev), torch.no_grad():
            _lib.check(lib.parq_raype_forward(B, T, H, W, Cc, self.num_samples, _ptr(feat), _ptr(cam), _ptr(Tcp), _ptr(Twp), _ptr(Twl),
                                              _ptr(self._depth), scale, _ptr(self._packed), _ptr(self._ws), self._ws.numel(),
                                              _ptr(tokens), _ptr(enc), self._flags | (_lib.PARQ_RAYPE_SPLIT_HIDDEN if want_encoding else 0) |
                                              (_lib.PARQ_RAYPE_FEAT_BF16 if feat_bf16 else 0),
                                              _stream()), "parq_raype_forward")
        return tokens, enc

    def forward(self, images_feat, camera=None, T_camera_pseudoCam=None, T_world_pseudoCam=None, T_world_local=None):
        """The reference's contract: (B,T,C,H,W) features in, (B,T,C,H,W) fp32 ray positional encoding out
        (hidden layer kept as an exact bf16 split: fp32-grade result)."""
        return self._run(images_feat, camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, False, True)[1]

    def tokens(self, images_feat, camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, out=None):
        """Fused producer: (B, T*H*W, C) bf16 channels-last tokens = features + encoding, the decoder's input.
        ``out``: optional preallocated token buffer (a fixed address lets the decoder replay one captured graph)."""
        return self._run(images_feat, camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local, True, False, out)[0]
