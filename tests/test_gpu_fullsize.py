"""Direct parity at the FULL sizes of BASELINE.json's configurations (2, 4, 5), teacher-forced per iteration.

The checker at these sizes is the UNMODIFIED reference module itself (oracle/ref_loader.py: /root/reference, or its
bytecode tree oracle/_ref on the GPU box) executed by stock PyTorch in fp32 on the same GPU with TF32 switched off
("GPU fp32 oracle with TF32 disabled at full shapes"); it is first cross-checked against the pinned CPU
oracle on one clip.  Projection bit-exactness is always judged against the machine-independent numpy restatement
(oracle/parq_oracle.py), because cuBLAS rounds the tiny pose matmuls differently from the CPU reference.
Bars: center_im / center_valid / coord_pos bit-exact; logits, centre, ortho6d, probabilities max|d|/max|ref| <= 1e-3;
size where the arg-max class agrees (flips counted and bounded).
"""
import numpy as np
import pytest
import torch

from conftest import OUT_KEYS, bit_equal, relerr
from oracle import parq_oracle as O
from oracle import ref_loader as RL
from parq_b200 import inputs as I
from parq_b200.decoder import DecoderEngine

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10, "needs an sm_100 device"
    return torch.device("cuda:0")


def reference_outputs(sd, Nq, tokens, cam, Tcp, Twp, Twl, dev, iters=8):
    """The checker: list of `iters` dicts (CPU tensors) from the unmodified reference on the GPU in fp32 (TF32 off);
    the CPU oracle port, clip by clip, when no reference tree travelled to this box."""
    if RL.reference_available():
        ns = RL.load_reference()
        m = RL.build_decoder(sd, Nq, iters, device=dev)
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.no_grad():
                outs = m(tokens.float().to(dev), ns.Camera(cam.to(dev)), ns.Pose(Tcp.to(dev)), ns.Pose(Twp.to(dev)), ns.Pose(Twl.to(dev)))
            outs = [{k: v.float().cpu() for k, v in o.items()} for o in outs]
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
            del m
            torch.cuda.empty_cache()
        return outs, "reference module on the GPU (fp32, TF32 off)"
    per_clip = [O.decoder_forward(tokens[b:b + 1].float(), cam[b:b + 1], Tcp[b:b + 1], Twp[b:b + 1], Twl[b:b + 1], sd, iters=iters)
                for b in range(tokens.shape[0])]
    return [{k: torch.cat([pc[i][k] for pc in per_clip]) for k in OUT_KEYS} for i in range(iters)], "CPU oracle port"


def check_teacher_forced(got, outs, iters, max_flips):
    flips, worst = 0, {}
    for i in range(iters):
        same = got["sem_cls_prob"][i].cpu().argmax(-1) == outs[i]["sem_cls_prob"].argmax(-1)
        flips += int((~same).sum())
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            e = relerr(got[k][i].cpu(), outs[i][k])
            worst[k] = max(worst.get(k, 0.0), e)
            assert e <= TOL, (k, i, e)
        e = relerr(got["size_unnormalized"][i].cpu()[same], outs[i]["size_unnormalized"][same])
        worst["size_unnormalized"] = max(worst.get("size_unnormalized", 0.0), e)
        assert e <= TOL, ("size_unnormalized", i, e)
    assert flips <= max_flips, "%d arg-max class flips" % flips
    return worst, flips


def check_projection_bit_exact(got, outs, refs, cam, Tcp, Twp, Twl, iters):
    """coord_pos / center_im / center_valid of every iteration, bit for bit, against the machine-independent numpy
    restatement (oracle/parq_oracle.py) applied to the teacher-forced reference points.  The checker's own coord_pos is
    compared to 1e-6: a GPU reference may contract p*span+lo into an FMA, the CPU reference (the named oracle device of the
    bit-exactness bar) does not."""
    Tcl = O.camera_from_local(Tcp.numpy(), Twp.numpy(), Twl.numpy())
    scale = (-3, 3, -2, 0.5, 0.25, 5.25)
    for i in range(iters):
        cp = O.denormalize(refs[i], scale)
        cp = cp if isinstance(cp, torch.Tensor) else torch.from_numpy(np.asarray(cp))
        assert bit_equal(got["coord_pos"][i], cp), "coord_pos, iteration %d" % i
        assert relerr(got["coord_pos"][i].cpu(), outs[i]["coord_pos"]) <= 1e-6
        pc = O.transform_points(Tcl, cp.numpy())
        cim, val = O.pinhole_project(cam.numpy(), pc)
        assert bit_equal(got["center_im"][i], cim), "center_im, iteration %d" % i
        assert np.array_equal(got["center_valid"][i].cpu().numpy(), val), "center_valid, iteration %d" % i


def test_config2_full_size_all_clips_against_reference(dev):
    # BASELINE.json configs[1], the benchmark workload itself: 16 clips x 8 views x 60x80 tokens, 256 queries, 8 iterations.
    # At B = 16 the stream-K plan cuts items in the middle of a clip (64 items over 74 CTA pairs).
    B, T, H, W, Nq, seed = 16, 8, 60, 80, 256, 31
    sd = I.make_weights(seed, Nq)
    tokens = torch.cat([I.make_tokens(1, T, H, W, seed=seed * 100 + b) for b in range(B)])
    cam, Tcp, Twp, Twl = (t._data for t in I.make_geometry(B, T, H, W, seed=seed))
    outs, how = reference_outputs(sd, Nq, tokens, cam, Tcp, Twp, Twl, dev)
    if how.startswith("reference"):
        # the GPU checker against the pinned CPU oracle on one clip: fp32 vs fp32, teacher-forcing not needed for iteration 0-1
        b = 5
        cpu = O.decoder_forward(tokens[b:b + 1], cam[b:b + 1], Tcp[b:b + 1], Twp[b:b + 1], Twl[b:b + 1], sd, iters=2)
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            assert relerr(outs[0][k][b:b + 1], cpu[0][k]) <= 5e-5, (k, "GPU fp32 checker vs CPU oracle")
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev)
    got = eng.forward(tokens.to(dev).bfloat16(), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev), debug=True)
    torch.cuda.synchronize()
    check_projection_bit_exact(got, outs, refs, cam, Tcp, Twp, Twl, 8)
    worst, flips = check_teacher_forced(got, outs, 8, max_flips=B * 2)
    print("config 2 full size vs %s: worst max|d|/max|ref| %s, %d arg-max flips of %d" % (how, {k: "%.1e" % v for k, v in worst.items()}, flips, 8 * B * Nq))
    # the same batch through the CUDA graph the benchmark replays
    rep = eng.forward(tokens.to(dev).bfloat16(), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev), debug=True, graph=True)
    torch.cuda.synchronize()
    for k in OUT_KEYS:
        assert torch.equal(rep[k], got[k]), k
    # the opt-in merge of the stream-K pieces inside the attention kernel (flags + spin wait) against the separate merge launch:
    # same pieces, another summation order
    fm = eng.forward(tokens.to(dev).bfloat16(), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev), debug=True,
                     fused_merge=True)
    torch.cuda.synchronize()
    for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob", "decoder_out"):
        assert relerr(fm[k].cpu(), got[k].cpu()) <= 1e-4, (k, "fused merge")
    # ... and the opt-in that drops the low-order activation term of the three 16-bit-output GEMMs (hi_only=7): inside the bar, a few
    # arg-max flips of near-tied queries (reported; the default keeps every term)
    ho = eng.forward(tokens.to(dev).bfloat16(), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev), debug=True,
                     hi_only=7)
    torch.cuda.synchronize()
    worst7, flips7 = check_teacher_forced(ho, outs, 8, max_flips=B * 2)
    print("config 2 with hi_only=7: worst %s, %d flips" % ({k: "%.1e" % v for k, v in worst7.items()}, flips7))


def test_config4_full_size_against_reference(dev):
    # BASELINE.json configs[3]: 1 clip, 32 views of 120x160 tokens (614 400 keys), 512 queries, 8 iterations; white-noise
    # tokens (the stress variant of SURVEY.md 8d)
    B, T, H, W, Nq, seed = 1, 32, 120, 160, 512, 43
    sd = I.make_weights(seed, Nq)
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randn(B, T * H * W, 1024, generator=g).bfloat16()
    cam, Tcp, Twp, Twl = (t._data for t in I.make_geometry(B, T, H, W, seed=seed))
    iters = 8 if RL.reference_available() else 2          # the CPU port at this size: first two iterations only
    outs, how = reference_outputs(sd, Nq, tokens, cam, Tcp, Twp, Twl, dev, iters=iters)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev, iters=iters)
    got = eng.forward(tokens.to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev), debug=True)
    torch.cuda.synchronize()
    check_projection_bit_exact(got, outs, refs, cam, Tcp, Twp, Twl, iters)
    worst, flips = check_teacher_forced(got, outs, iters, max_flips=4)
    print("config 4 full size vs %s (%d iterations): worst %s, %d flips" % (how, iters, {k: "%.1e" % v for k, v in worst.items()}, flips))


def test_config5_sliding_windows_against_reference(dev):
    # BASELINE.json configs[4]: one clip, an 8-view window sliding by one view; every window is compared with the
    # reference run on that window (local frame = pseudo-camera of the window's middle view, datasets/transforms.py:201-208)
    T, H, W, Nq, seed, nwin = 8, 60, 80, 256, 51, 3
    sd = I.make_weights(seed, Nq)
    stream = I.make_tokens(1, T + nwin - 1, H, W, seed=seed)[0].view(T + nwin - 1, H * W, 1024)
    cam, Tcp, Twp, _ = (t._data for t in I.make_geometry(1, T + nwin - 1, H, W, seed=seed))
    eng = DecoderEngine(sd, dev)
    window = torch.empty(1, T * H * W, 1024, dtype=torch.bfloat16, device=dev)       # the application's window buffer (fixed address)
    for s in range(nwin):
        tok = stream[s:s + T].reshape(1, T * H * W, 1024).contiguous()
        window.copy_(tok)
        c, tcp, twp = cam[:, s:s + T].contiguous(), Tcp[:, s:s + T].contiguous(), Twp[:, s:s + T].contiguous()
        twl = Twp[:, s + T // 2: s + T // 2 + 1].contiguous()
        outs, how = reference_outputs(sd, Nq, tok, c, tcp, twp, twl, dev)
        refs = O.refs_from_outputs(outs, sd)
        got = eng.forward(window, c.to(dev), tcp.to(dev), twp.to(dev), twl.to(dev), H, W, forced_refs=refs.to(dev), debug=True, graph=True)
        torch.cuda.synchronize()
        check_projection_bit_exact(got, outs, refs, c, tcp, twp, twl, 8)
        check_teacher_forced(got, outs, 8, max_flips=2)
    assert len(eng._graphs) == 1                # one captured graph serves every window (shape-keyed, static inputs)


def test_fp32_tokens_report(dev):
    """VERDICT r1 weak #2 / ADVICE: the pipeline feeds arbitrary fp32 tokens (parq_lightning.py:75-85), not bf16-representable
    ones.  GPU path on NON-representable fp32 tokens and fp32 (checkpoint-style) weights against the fp32 checker on the same
    tokens, teacher-forced per iteration; written to gpurun_out/fp32_tokens.md and gated at the 1e-3 bar."""
    import os
    B, T, H, W, Nq, seed = 2, 8, 60, 80, 256, 91
    sd = I.make_weights(seed, Nq, bf16_exact=False)
    g = torch.Generator().manual_seed(seed)
    tokens = I.make_tokens(B, T, H, W, seed=seed).float()
    tokens = tokens * (1 + 2.0 ** -9 * (torch.rand(tokens.shape, generator=g) - 0.5))       # off the bf16 grid
    assert (tokens.bfloat16().float() != tokens).float().mean() > 0.9
    cam, Tcp, Twp, Twl = (t._data for t in I.make_geometry(B, T, H, W, seed=seed))
    outs, how = reference_outputs(sd, Nq, tokens, cam, Tcp, Twp, Twl, dev)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev)
    got = eng.forward(tokens.to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev), debug=True)
    torch.cuda.synchronize()
    keys = ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob")
    lines = ["| iteration | " + " | ".join(keys) + " |", "|---|" + "---|" * len(keys)]
    worst = 0.0
    for i in range(8):
        row = [relerr(got[k][i].cpu(), outs[i][k]) for k in keys]
        worst = max(worst, max(row))
        lines.append("| %d | " % i + " | ".join("%.2e" % v for v in row) + " |")
    report = "\n".join(lines)
    print("\nfp32 (non-bf16-representable) tokens + fp32 weights vs %s, teacher-forced:\n%s" % (how, report))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/fp32_tokens.md", "w") as f:
            f.write("GPU path on fp32 tokens that are NOT bf16-representable (+ fp32 weights) vs %s on the same tokens; max|d|/max|ref|, "
                    "teacher-forced per iteration, %d clips x %d views x %dx%d, %d queries\n\n%s\n" % (how, B, T, H, W, Nq, report))
    except OSError:
        pass
    assert worst <= TOL, "fp32-token parity %.2e exceeds the 1e-3 bar" % worst
