"""Streaming window cache (f-4, BASELINE.json configs[4]): per-view K / V^T updates + decode over the cached window.

Parity argument, in three steps:
  1. the cache is exact: after every push, decode() equals a from-scratch forward over the ring's views, bit for bit
     (a view's K / V^T rows do not depend on which GEMM tile computed them);
  2. against the reference run on each window in temporal order (its own local frame), teacher-forced: the 1e-3 bar
     (the ring permutes the views, which only changes fp32 summation orders);
  3. anchor-frame decoding (tokens never re-encoded, reference points re-expressed) changes semantics: its deviation from
     the per-window reference is REPORTED (gpurun_out/streaming_anchor.md), together with the invariants that do hold
     (projected pixels and sampled features are unchanged by the change of frame).
"""
import os

import pytest
import torch

from conftest import OUT_KEYS, relerr
from oracle import parq_oracle as O
from parq_b200 import inputs as I
from parq_b200.decoder import DecoderEngine
from parq_b200.streaming import StreamingWindow
from test_gpu_fullsize import check_teacher_forced, reference_outputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10, "needs an sm_100 device"
    return torch.device("cuda:0")


def _stream_inputs(T, H, W, n, seed):
    tokens = I.make_tokens(1, n, H, W, seed=seed)[0].view(n, H * W, 1024)
    cam, Tcp, Twp, _ = (t._data for t in I.make_geometry(1, n, H, W, seed=seed))
    return tokens, cam, Tcp, Twp


def test_streaming_cache_is_exact_and_matches_reference(dev):
    T, H, W, Nq, seed, n = 8, 60, 80, 256, 51, 11
    sd = I.make_weights(seed, Nq)
    tokens, cam, Tcp, Twp = _stream_inputs(T, H, W, n, seed)
    eng, fresh = DecoderEngine(sd, dev), DecoderEngine(sd, dev)
    sw = StreamingWindow(eng, T, H, W)
    for v in range(n):
        sw.push(tokens[v:v + 1].to(dev).bfloat16(), cam[:, v].to(dev), Tcp[:, v].to(dev), Twp[:, v].to(dev))
        if not sw.full:
            continue
        s = v - T + 1                                               # window = views s .. s+T-1, local frame = its middle view
        twl = Twp[:, s + T // 2: s + T // 2 + 1].contiguous()
        got = {k: x.clone() for k, x in sw.decode(twl.to(dev)).items()}
        torch.cuda.synchronize()
        # 1. from-scratch forward over the ring's views (all K / V^T re-projected): bit-identical
        ref = fresh.forward(sw.tokens.clone(), sw.camera, sw.T_cp, sw.T_wp, twl.to(dev), H, W)
        torch.cuda.synchronize()
        for k in OUT_KEYS:
            assert torch.equal(got[k], ref[k]), (k, v)
        # 2. the reference on the window in temporal order, teacher-forced
        if v in (T - 1, n - 1):
            tok = tokens[s:s + T].reshape(1, T * H * W, 1024)
            c, tcp, twp = cam[:, s:s + T].contiguous(), Tcp[:, s:s + T].contiguous(), Twp[:, s:s + T].contiguous()
            outs, how = reference_outputs(sd, Nq, tok, c, tcp, twp, twl, dev)
            refs = O.refs_from_outputs(outs, sd)
            forced = sw.decode(twl.to(dev), forced_refs=refs.to(dev))
            torch.cuda.synchronize()
            check_teacher_forced(forced, outs, 8, max_flips=2)
    assert sw.slot_order() == [(n % T + i) % T for i in range(T)]


def test_streaming_anchor_frame_report(dev):
    T, H, W, Nq, seed, n = 8, 32, 40, 256, 53, 12
    sd = I.make_weights(seed, Nq)
    tokens, cam, Tcp, Twp = _stream_inputs(T, H, W, n, seed)
    eng = DecoderEngine(sd, dev)
    sw = StreamingWindow(eng, T, H, W)
    anchor = Twp[:, T // 2].contiguous()                                # frame of the first window's middle view
    lines = ["| window | frame offset [m] | center_im max dev [px] | features | centre it0 | logits it0 | centre it7 | logits it7 |", "|---|---|---|---|---|---|---|---|"]
    for v in range(n):
        sw.push(tokens[v:v + 1].to(dev).bfloat16(), cam[:, v].to(dev), Tcp[:, v].to(dev), Twp[:, v].to(dev))
        if not sw.full:
            continue
        s = v - T + 1
        twl = Twp[:, s + T // 2].contiguous()
        exact = {k: x.clone() for k, x in sw.decode(twl.to(dev).view(1, 1, 12), debug=True).items()}
        outs_l, outs_a = sw.decode_anchored(anchor.to(dev), twl.to(dev))
        dbg = {k: x.clone() for k, x in sw.decode(anchor.to(dev).view(1, 1, 12), ref0=None, debug=True).items()}
        torch.cuda.synchronize()
        for k in OUT_KEYS:
            assert torch.isfinite(outs_l[k]).all(), k
        if s == 0:
            # anchor == local frame: the two modes coincide (identity change of frame, up to the rounding of R R^T)
            assert relerr(outs_l["center_unnormalized"][0], exact["center_unnormalized"][0]) <= 1e-4
            assert relerr(outs_l["pred_logits"][0], exact["pred_logits"][0]) <= 1e-3
        off = float((twl[0, 9:] - anchor[0, 9:]).norm())
        # invariants of the change of frame at iteration 0: same pixels, same sampled features (need the anchored run's debug outputs)
        lo, span = torch.tensor(eng.scale[0::2], device=dev), torch.tensor(eng.scale[1::2], device=dev) - torch.tensor(eng.scale[0::2], device=dev)
        ref0_a = ((outs_a["coord_pos"][0] - lo) / span).contiguous()
        a_dbg = sw.decode(anchor.to(dev).view(1, 1, 12), ref0=ref0_a, debug=True)
        torch.cuda.synchronize()
        both = a_dbg["center_valid"][0] & exact["center_valid"][0]
        px = float((a_dbg["center_im"][0] - exact["center_im"][0]).abs()[both].max()) if both.any() else 0.0
        lines.append("| %d | %.2f | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (
            s, off, px, relerr(a_dbg["features"][0], exact["features"][0]),
            relerr(outs_l["center_unnormalized"][0], exact["center_unnormalized"][0]), relerr(outs_l["pred_logits"][0], exact["pred_logits"][0]),
            relerr(outs_l["center_unnormalized"][7], exact["center_unnormalized"][7]), relerr(outs_l["pred_logits"][7], exact["pred_logits"][7])))
        assert px <= 1e-2                                           # the projection is frame independent (fp32 rounding of the pose products)
        del dbg
    report = "\n".join(lines)
    print("\n" + report)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/streaming_anchor.md", "w") as f:
            f.write("Anchor-frame streaming decode (tokens encoded once, reference points re-expressed) vs the exact per-window decode; "
                    "max|d|/max|ref|, random-init weights (the decoder is not frame-equivariant: the reference-point MLP and the box update see "
                    "coordinates in another frame), 1 clip, 8-view window of 32x40 tokens sliding by one view\n\n" + report + "\n")
    except OSError:
        pass
