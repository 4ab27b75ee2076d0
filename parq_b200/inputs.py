"""Deterministic synthetic inputs and weights for parity tests and the benchmark.

Everything is generated from explicit ``torch.Generator`` seeds on the CPU so
that the build container (where the golden fixtures are produced with the real
reference) and the GPU box regenerate bit-identical tensors; the fixtures store
checksums of what they were generated from.

Shapes/semantics follow the reference's data pipeline:
  tokens   (B, T*H*W, C)  channels-last image tokens, token order (t, h, w)
           (model/parq_lightning.py:78-85)
  camera   (B, T, 6)      [w, h, fx, fy, cx, cy] at feature-map scale
           (model/resnet_fpn.py:89-90, utils/wrappers.py:478-488)
  T_camera_pseudoCam, T_world_pseudoCam (B, T, 12); T_world_local (B, 1, 12)
           = pseudo-camera pose of the middle view (datasets/transforms.py:201-208)
"""
import hashlib
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .wrappers import Camera, Pose

DEC_DIM = 1024
NUM_HEADS = 4
FFN_DIM = 768
NUM_CLS = 10          # NUM_SEMCLS + 1 (background)
POS_FEATS = 128       # pos2posemb3d num_pos_feats (transformer_parq.py:45)
SCALE = (-3.0, 3.0, -2.0, 0.5, 0.25, 5.25)   # config/eval.yaml:55


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """fp32 tensor whose values are exactly representable in bf16."""
    return x.to(torch.bfloat16).to(torch.float32)


def state_dict_spec(num_queries=256, dim=DEC_DIM, ffn=FFN_DIM, num_cls=NUM_CLS):
    """(key, shape) list of the reference decoder's state dict, in its order
    (model/parq_decoder.py:35-132, model/transformer_parq.py:164-181,340-357)."""
    spec = []

    def heads(prefix):
        spec.append((prefix + "sem_cls_head.layers.0.weight", (num_cls, dim, 1)))
        spec.append((prefix + "sem_cls_head.layers.0.bias", (num_cls,)))
        for name, out in (("center_head", 3), (None, None), ("rotation_head", 6)):
            if name is None:
                spec.append((prefix + "size_head.layers.0.weight", (3, dim, 1)))
                spec.append((prefix + "size_head.layers.0.bias", (3,)))
                continue
            spec.append((prefix + name + ".layers.0.weight", (dim, dim, 1)))
            spec.append((prefix + name + ".layers.1.weight", (dim,)))
            spec.append((prefix + name + ".layers.1.bias", (dim,)))
            spec.append((prefix + name + ".layers.4.weight", (dim, dim, 1)))
            spec.append((prefix + name + ".layers.5.weight", (dim,)))
            spec.append((prefix + name + ".layers.5.bias", (dim,)))
            spec.append((prefix + name + ".layers.8.weight", (out, dim, 1)))
            spec.append((prefix + name + ".layers.8.bias", (out,)))

    heads("mlp_heads.")
    L = "parq_module.decoder.layers.0."
    for attn in ("self_attn", "multihead_attn"):
        spec.append((L + attn + ".in_proj_weight", (3 * dim, dim)))
        spec.append((L + attn + ".in_proj_bias", (3 * dim,)))
        spec.append((L + attn + ".out_proj.weight", (dim, dim)))
        spec.append((L + attn + ".out_proj.bias", (dim,)))
    spec.append((L + "linear1.weight", (ffn, dim)))
    spec.append((L + "linear1.bias", (ffn,)))
    spec.append((L + "linear2.weight", (dim, ffn)))
    spec.append((L + "linear2.bias", (dim,)))
    for n in ("norm1", "norm2", "norm3"):
        spec.append((L + n + ".weight", (dim,)))
        spec.append((L + n + ".bias", (dim,)))
    spec.append(("parq_module.decoder.norm.weight", (dim,)))
    spec.append(("parq_module.decoder.norm.bias", (dim,)))
    P = "parq_module.decoder.position_encoder."
    spec.append((P + "0.weight", (dim, 3 * POS_FEATS)))
    spec.append((P + "0.bias", (dim,)))
    spec.append((P + "2.weight", (dim, dim)))
    spec.append((P + "2.bias", (dim,)))
    heads("parq_module.decoder.mlp_heads.")
    spec.append(("refpoint.weight", (num_queries, 3)))
    return spec


def make_weights(seed=0, num_queries=256, bf16_exact=True, n_layers=1):
    """Random-init decoder weights as a state dict with the reference's 65 keys.

    Matrices use the reference's init families (xavier-uniform inside
    ``parq_module``, transformer_parq.py:89-92; kaiming-uniform Conv1d default on
    the heads); vectors (biases, LayerNorm/GroupNorm affine) are perturbed away
    from their 0/1 defaults so that every bias/affine path is exercised by the
    parity tests.  With ``bf16_exact`` every matrix is rounded to
    bf16-representable fp32 so both sides of a parity test consume identical
    values.  ``n_layers`` > 1: SHARE_WEIGHTS False, one distinct decoder layer per iteration
    (keys ``parq_module.decoder.layers.{i}.*``, transformer_parq.py:168-171).
    """
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd = OrderedDict()
    spec = []
    for key, shape in state_dict_spec(num_queries):
        if key.startswith("parq_module.decoder.layers.0.") and n_layers > 1:
            continue
        spec.append((key, shape))
    if n_layers > 1:
        layer0 = [(k, sh) for k, sh in state_dict_spec(num_queries) if k.startswith("parq_module.decoder.layers.0.")]
        at = next(i for i, (k, _) in enumerate(spec) if k.startswith("parq_module.decoder.norm."))
        spec[at:at] = [(k.replace("layers.0.", "layers.%d." % i), sh) for i in range(n_layers) for k, sh in layer0]
    for key, shape in spec:
        if key.startswith("parq_module.decoder.mlp_heads."):
            sd[key] = sd[key[len("parq_module.decoder."):]]      # alias, parq_decoder.py:66
            continue
        if key == "refpoint.weight":
            w = torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_out, fan_in = shape[0], shape[1]
            if key.startswith("mlp_heads."):
                bound = 1.0 / math.sqrt(fan_in)                    # kaiming_uniform(a=sqrt(5))
            else:
                bound = math.sqrt(6.0 / (fan_in + fan_out))        # xavier_uniform
            w = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if bf16_exact:
                w = bf16_round(w)
        elif key.endswith(".weight"):                              # 1-D weight = LayerNorm/GroupNorm scale
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            w = 0.05 * torch.randn(shape, generator=g)
        sd[key] = w.contiguous()
    return sd


def make_tokens(B, T, H, W, C=DEC_DIM, seed=0, smooth=True):
    """(B, T*H*W, C) fp32, bf16-representable: a unit-variance smooth Gaussian
    field (coarse noise bilinearly upsampled, align_corners=True) + 0.05 white
    noise; ``smooth=False`` gives pure white noise (stress variant)."""
    out = torch.empty(B, T * H * W, C)
    for b in range(B):
        g = torch.Generator().manual_seed(1234 + seed * 7919 + b)
        if smooth:
            hc, wc = max(2, (H + 7) // 8), max(2, (W + 7) // 8)
            coarse = torch.randn(T, C, hc, wc, generator=g)
            fine = F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True)
            fine = fine / fine.std().clamp(min=1e-6)
            fine = fine + 0.05 * torch.randn(T, C, H, W, generator=g)
        else:
            fine = torch.randn(T, C, H, W, generator=g)
        out[b] = fine.permute(0, 2, 3, 1).reshape(T * H * W, C)
    return bf16_round(out)


def make_raype_weights(seed=0, dim=DEC_DIM, num_samples=64, bf16_exact=True):
    """State dict of the reference's AddRayPE encoder (model/ray_positional_encoding.py:55-59):
    Linear(3*num_samples, dim) -> ReLU -> Linear(dim, dim), torch's default Linear init families."""
    g = torch.Generator().manual_seed(7000003 * seed + 29)
    sd = OrderedDict()
    for name, (fo, fi) in (("encoder.0", (dim, 3 * num_samples)), ("encoder.2", (dim, dim))):
        bound = 1.0 / math.sqrt(fi)
        w = (torch.rand(fo, fi, generator=g) * 2 - 1) * bound
        sd[name + ".weight"] = (bf16_round(w) if bf16_exact else w).contiguous()
        sd[name + ".bias"] = ((torch.rand(fo, generator=g) * 2 - 1) * bound).contiguous()
    return sd


def make_features(B, T, H, W, C=DEC_DIM, seed=0):
    """(B, T, C, H, W) fp32 backbone-style feature maps (the layout ResnetFPN hands to AddRayPE,
    model/resnet_fpn.py:73-85): the same smooth field as make_tokens, channels-first."""
    tok = make_tokens(B, T, H, W, C, seed=seed + 500)
    return tok.view(B, T, H, W, C).permute(0, 1, 4, 2, 3).contiguous()


def make_pyramid(N, H, W, Cl=256, seed=0):
    """The four FPN levels a torchvision resnet_fpn_backbone returns for level-0 size (H, W):
    {"0": (N,Cl,H,W), "1": ceil/2, "2": ceil/4, "3": ceil/8} fp32 (model/resnet_fpn.py:66-75)."""
    g = torch.Generator().manual_seed(9000011 * seed + 3)
    out, h, w = {}, H, W
    for l in range(4):
        out[str(l)] = torch.randn(N, Cl, h, w, generator=g)
        h, w = (h + 1) // 2, (w + 1) // 2
    return out


def _rot(axis, ang):
    c, s = torch.cos(ang), torch.sin(ang)
    o, z = torch.ones_like(ang), torch.zeros_like(ang)
    if axis == "x":
        m = [o, z, z, z, c, -s, z, s, c]
    elif axis == "y":
        m = [c, z, s, z, o, z, -s, z, c]
    else:
        m = [c, -s, z, s, c, z, z, z, o]
    return torch.stack(m, -1).reshape(ang.shape + (3, 3))


def make_geometry(B, T, H, W, seed=0, wild=False):
    """ScanNet-like camera and poses.  ``wild=True`` draws
    unconstrained rotations so that most views are invalid/behind the camera
    (edge-case coverage for the zero-padding and valid-count semantics)."""
    g = torch.Generator().manual_seed(4321 + seed * 104729)
    cam = torch.tensor([W, H, 0.9025 * W, 0.9025 * W, (W - 1) / 2 + 0.125, (H - 1) / 2 + 0.125], dtype=torch.float32)
    camera = Camera(cam.expand(B, T, 6).contiguous())
    deg = math.pi / 180.0
    if wild:
        a = torch.rand(B, T, 3, generator=g) * 2 * math.pi
        R_cp = _rot("x", a[..., 0]) @ _rot("y", a[..., 1]) @ _rot("z", a[..., 2])
        a = torch.rand(B, T, 3, generator=g) * 2 * math.pi
        R_wp = _rot("y", a[..., 0]) @ _rot("x", a[..., 1]) @ _rot("z", a[..., 2])
        t_wp = torch.randn(B, T, 3, generator=g)
        t_cp = 0.1 * torch.randn(B, T, 3, generator=g)
    else:
        pitch = 5 * deg * torch.randn(B, T, generator=g)
        roll = 5 * deg * torch.randn(B, T, generator=g)
        R_cp = _rot("x", pitch) @ _rot("z", roll)
        t_cp = torch.zeros(B, T, 3)
        k = torch.arange(T, dtype=torch.float32) - T / 2
        yaw = (15 * deg * k)[None, :] + 3 * deg * torch.randn(B, T, generator=g)
        R_wp = _rot("y", yaw)
        t_wp = torch.zeros(B, T, 3)
        t_wp[..., 0] = 0.1 * k[None, :]
        t_wp = t_wp + 0.05 * torch.randn(B, T, 3, generator=g)
    T_camera_pseudoCam = Pose.from_Rt(R_cp.float(), t_cp.float())
    T_world_pseudoCam = Pose.from_Rt(R_wp.float(), t_wp.float())
    T_world_local = Pose(T_world_pseudoCam._data[:, T // 2: T // 2 + 1].clone())
    return camera, T_camera_pseudoCam, T_world_pseudoCam, T_world_local


def tensor_checksum(*tensors) -> str:
    """sha256 over the raw bytes of the given tensors (fixture provenance)."""
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]
