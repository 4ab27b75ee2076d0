"""Generates tests/golden/*.npz by running the UNMODIFIED reference decoder
(/root/reference, imported through oracle/ref_loader.py) on the deterministic synthetic
inputs of parq_b200.inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin oracle/parq_oracle.py (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py).  Inputs/weights are NOT stored: they are regenerated from the
seeds and verified against the checksums recorded here.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_loader import decoder_cfg, load_reference  # noqa: E402
from parq_b200 import inputs as I  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FEAT_STRIDE = 16     # sampled features are stored for every 16th channel

CASES = {
    # name: (B, T, H, W, Nq, seed, wild, smooth)
    "small": (2, 3, 12, 16, 256, 0, False, True),
    "ragged_wild": (1, 2, 10, 14, 128, 1, True, True),
    "white_noise": (1, 4, 15, 20, 256, 2, False, False),
}
RAYPE_CASES = {
    # AddRayPE + tokeniser (f-1): (B, T, H, W, seed)
    "raype_small": (2, 3, 12, 16, 6),
    "raype_c1_view": (1, 2, 60, 80, 7),
}
PROJ_CASES = {
    # projection-only cases at benchmark geometry: (B, T, H, W, Nq, seed, wild)
    "proj_c1": (1, 8, 60, 80, 256, 3, False),
    "proj_c4_views": (1, 32, 30, 40, 512, 4, False),
    "proj_wild": (2, 5, 24, 32, 128, 5, True),
}


def main():
    ns = load_reference()
    torch.manual_seed(0)
    for name, (B, T, H, W, Nq, seed, wild, smooth) in CASES.items():
        sd = I.make_weights(seed, Nq)
        ref = ns.PARQDecoder(decoder_cfg(Nq)).eval()
        ref.load_state_dict(sd, strict=True)
        tokens = I.make_tokens(B, T, H, W, seed=seed, smooth=smooth)
        cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed, wild=wild)
        rc, rp = ns.Camera(cam._data), ns.Pose
        with torch.no_grad():
            outs = ref(tokens, rc, rp(Tcp._data), rp(Twp._data), rp(Twl._data))
            Tcl = rp(Tcp._data) @ (rp(Twp._data).inverse() @ rp(Twl._data))
            memory_hw = tokens.view(B * T, H, W, -1).permute(0, 3, 1, 2)
            feats, cims, vals = [], [], []
            for o in outs:
                f, ci, cv = ns.project(memory_hw, o["coord_pos"], Tcl, rc)
                feats.append(f[..., ::FEAT_STRIDE].numpy())
                cims.append(ci.numpy())
                vals.append(cv.numpy())
        data = {"shape": np.array([B, T, H, W, Nq, seed, int(wild), int(smooth)]),
                "weights_sum": np.array(I.tensor_checksum(*[sd[k] for k in sorted(sd)])),
                "inputs_sum": np.array(I.tensor_checksum(tokens, cam._data, Tcp._data, Twp._data, Twl._data)),
                "T_camera_local": Tcl._data.numpy(),
                "features": np.stack(feats), "center_im": np.stack(cims), "center_valid": np.stack(vals)}
        for k in outs[0]:
            data[k] = np.stack([o[k].numpy() for o in outs])
        # post-NMS detection set of the reference's own parse_pred (parq_decoder.py:372-424).  Its only CUDA
        # dependency is `obbs_pred.cuda()` (:403, a device move of the result); it is made a no-op here so the
        # unmodified method runs on the CPU of the build container.
        ns.Obb3D.cuda = lambda self: self
        parsed = ref.parse_pred([dict(o) for o in outs])
        data["pred_mask"] = parsed["pred_mask"].numpy()
        data["obbs_pred"] = parsed["obbs_pred"]._data.numpy()
        # the NMS decision alone (before the track-scale filter), from the reference's own nms() (utils/nms.py:20-32)
        scores = torch.max(outs[-1]["sem_cls_prob"], -1)[0]
        data["nms_mask"] = np.asarray(ns.decoder_module.nms(parsed["obbs_pred"], scores, 9, 0.1, "nms_3d_faster"))
        # the FOR_VIS branch of the same unmodified method (:407-421): same-class NMS at IoU 0.2, no track-scale filter
        ref.for_vis = True
        parsed_vis = ref.parse_pred([dict(o) for o in outs])
        ref.for_vis = False
        data["pred_mask_vis"] = parsed_vis["pred_mask"].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: v.shape for k, v in data.items() if hasattr(v, "shape")})
    # SHARE_WEIGHTS False (transformer_parq.py:168-171, 311-314): one distinct decoder layer per iteration
    for name, (B, T, H, W, Nq, seed, layers) in {"unshared": (1, 2, 10, 14, 128, 5, 3)}.items():
        from oracle.ref_loader import build_decoder
        sd = I.make_weights(seed, Nq, n_layers=layers)
        ref = build_decoder(sd, Nq, layers)
        tokens = I.make_tokens(B, T, H, W, seed=seed)
        cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
        with torch.no_grad():
            outs = ref(tokens, ns.Camera(cam._data), ns.Pose(Tcp._data), ns.Pose(Twp._data), ns.Pose(Twl._data))
        data = {"shape": np.array([B, T, H, W, Nq, seed, layers]),
                "weights_sum": np.array(I.tensor_checksum(*[sd[k] for k in sorted(sd)])),
                "inputs_sum": np.array(I.tensor_checksum(tokens, cam._data, Tcp._data, Twp._data, Twl._data))}
        for k in outs[0]:
            data[k] = np.stack([o[k].numpy() for o in outs])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, {k: v.shape for k, v in data.items() if hasattr(v, "shape")})
    from oracle.ref_loader import load_module
    rpe = load_module("model.ray_positional_encoding")
    for name, (B, T, H, W, seed) in RAYPE_CASES.items():
        sd = I.make_raype_weights(seed)
        ref = rpe.AddRayPE(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()      # config/eval.yaml:30-35
        ref.load_state_dict(sd, strict=True)
        feat = I.make_features(B, T, H, W, seed=seed)
        cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
        with torch.no_grad():
            enc = ref(feat, ns.Camera(cam._data), ns.Pose(Tcp._data), ns.Pose(Twp._data), ns.Pose(Twl._data))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array([B, T, H, W, seed]),
                            inputs_sum=np.array(I.tensor_checksum(feat, cam._data, Tcp._data, Twp._data, Twl._data, *[sd[k] for k in sorted(sd)])),
                            encoding=enc[:, :, ::FEAT_STRIDE].numpy())
        print(name, tuple(enc.shape), "max |enc| %.3f" % enc.abs().max().item())
    # f-3: the concat part of the reference's ResnetFPN.forward (model/resnet_fpn.py:56-91), run UNMODIFIED on an
    # instance whose backbone is replaced by a stub that returns a seeded pyramid (the torchvision backbone itself is
    # out of scope and needs downloaded weights)
    rfpn = load_module("model.resnet_fpn")
    for name, (B, T, H, W, seed) in {"fpn_small": (1, 2, 60, 80, 8), "fpn_odd": (2, 1, 15, 21, 9)}.items():
        pyr = I.make_pyramid(B * T, H, W, seed=seed)
        obj = rfpn.ResnetFPN.__new__(rfpn.ResnetFPN)
        torch.nn.Module.__init__(obj)
        obj.resnet_fpn, obj.transform, obj.freeze, obj.layer = (lambda x: pyr), (lambda x: x), False, "0"
        cam = I.make_geometry(B, T, H * 4, W * 4, seed=seed)[0]
        with torch.no_grad():
            batch = obj.forward({"rgb_img": torch.zeros(B, T, 3, 4 * H, 4 * W), "camera": ns.Camera(cam._data)})
        af = batch["all_features"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array([B, T, H, W, seed]),
                            inputs_sum=np.array(I.tensor_checksum(*[pyr[str(l)] for l in range(4)])),
                            all_features=af[:, :, ::FEAT_STRIDE].numpy(), camera_feature=batch["camera_feature"]._data.numpy())
        print(name, tuple(af.shape))
    for name, (B, T, H, W, Nq, seed, wild) in PROJ_CASES.items():
        tokens = I.make_tokens(B, T, H, W, seed=seed)
        cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed, wild=wild)
        g = torch.Generator().manual_seed(seed)
        pts = torch.rand(B, Nq, 3, generator=g) * torch.tensor([6.0, 2.5, 5.0]) + torch.tensor([-3.0, -2.0, 0.25])
        rc, rp = ns.Camera(cam._data), ns.Pose
        with torch.no_grad():
            Tcl = rp(Tcp._data) @ (rp(Twp._data).inverse() @ rp(Twl._data))
            f, ci, cv = ns.project(tokens.view(B * T, H, W, -1).permute(0, 3, 1, 2), pts, Tcl, rc)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            shape=np.array([B, T, H, W, Nq, seed, int(wild), 1]),
                            inputs_sum=np.array(I.tensor_checksum(tokens, cam._data, Tcp._data, Twp._data, Twl._data, pts)),
                            points=pts.numpy(), T_camera_local=Tcl._data.numpy(), features=f[..., ::FEAT_STRIDE].numpy(),
                            center_im=ci.numpy(), center_valid=cv.numpy())
        print(name, "valid fraction %.3f" % cv.float().mean().item())


if __name__ == "__main__":
    main()
