"""Parity of the CUDA path (through the C ABI) against the golden fixtures of the unmodified
reference, against the CPU oracle on seeded inputs, and -- at benchmark sizes -- through
size-independent properties.  Run on a B200: `pytest -m gpu`.

Bars: center_im / center_valid / coord_pos bit-exact in fp32; sampled features,
box parameters, logits and probabilities max|d|/max|ref| <= 1e-3, teacher-forced per iteration.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import OUT_KEYS, bit_equal, load_golden, regenerate_case, relerr
from oracle import parq_oracle as O
from parq_b200 import _lib, inputs as I
from parq_b200.decoder import DecoderEngine, PARQDecoderB200, _ptr, _stream, default_cfg, pose_chain, project

pytestmark = pytest.mark.gpu
TOL = 1e-3          # north_star: "within 1e-3 relative (bf16 attention)"
FEAT_TOL = 1e-5     # the gather itself is fp32 on identical bf16 token values


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10, "needs an sm_100 device"
    return torch.device("cuda:0")


def _engine_forward(eng, c, dev, **kw):
    out = eng.forward(c["tokens"].to(dev), c["camera"].to(dev), c["T_cp"].to(dev), c["T_wp"].to(dev), c["T_wl"].to(dev),
                      c["H"], c["W"], **kw)
    torch.cuda.synchronize()
    return out


# ---------------------------------------------------------------- projection: bit-exact ----
@pytest.mark.parametrize("name", ["proj_c1", "proj_c4_views", "proj_wild"])
def test_projection_against_reference_golden(dev, name):
    gold = load_golden(name)
    c = regenerate_case(gold)
    Tcl = pose_chain(c["T_cp"].to(dev), c["T_wp"].to(dev), c["T_wl"].to(dev))
    assert bit_equal(Tcl, gold["T_camera_local"]), "T_camera_local differs from the reference"
    feat, cim, val = project(c["tokens"].to(dev), c["points"].to(dev), Tcl, c["camera"].to(dev), c["H"], c["W"])
    assert bit_equal(cim, gold["center_im"]), "%d coordinates differ" % int((cim.cpu().numpy() != gold["center_im"]).sum())
    assert np.array_equal(val.cpu().numpy(), gold["center_valid"])
    assert relerr(feat.cpu()[..., ::16], gold["features"]) <= FEAT_TOL


def test_projection_edge_cases(dev):
    # points exactly on the image border, behind the camera, at the clamp depth, and far outside
    B, T, H, W, Nq = 1, 2, 6, 8, 128
    tokens = I.make_tokens(B, T, H, W, seed=11)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=11)
    eye = torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0]).expand(B, T, 12).contiguous()
    pts = torch.zeros(B, Nq, 3)
    pts[0, :, 2] = 1.0
    f, cx, cy = cam._data[0, 0, 2].item(), cam._data[0, 0, 4].item(), cam._data[0, 0, 5].item()
    special = [((0 - cx) / f, (0 - cy) / f, 1.0), ((W - 1 - cx) / f, (H - 1 - cy) / f, 1.0), (0, 0, -1.0), (0, 0, 1e-3),
               (0, 0, 9.9e-4), (50.0, 50.0, 1.0), (-50.0, 0.0, 1.0), ((W - 1.5 - cx) / f, (H - 0.5 - cy) / f, 1.0)]
    for i, p in enumerate(special):
        pts[0, i] = torch.tensor(p)
    f_ref, ci_ref, cv_ref = O.project_sample(tokens, pts, eye.numpy(), cam._data.numpy(), H, W)
    feat, cim, val = project(tokens.to(dev), pts.to(dev), eye.to(dev), cam._data.to(dev), H, W)
    assert bit_equal(cim, ci_ref) and torch.equal(val.cpu(), cv_ref)
    assert relerr(feat.cpu(), f_ref) <= FEAT_TOL


# ------------------------------------------------------- full decoder: teacher-forced ----
@pytest.mark.parametrize("chain", [None, True], ids=["auto", "chained"])
@pytest.mark.parametrize("name", ["small", "ragged_wild", "white_noise"])
def test_decoder_against_reference_golden(dev, name, chain):
    # chain=True forces the chained cluster kernel (chain_tc.cuh), which the library picks by itself only from 2048 rows
    gold = load_golden(name)
    c = regenerate_case(gold)
    sd = c["sd"]
    iters = gold["coord_pos"].shape[0]
    gold_outs = [{k: torch.from_numpy(gold[k][i]) for k in OUT_KEYS} for i in range(iters)]
    refs = O.refs_from_outputs(gold_outs, sd)
    eng = DecoderEngine(sd, dev)
    got = _engine_forward(eng, c, dev, forced_refs=refs.to(dev), debug=True, chain=chain)
    flips = 0
    for i in range(iters):
        assert bit_equal(got["coord_pos"][i], gold["coord_pos"][i]), "iteration %d" % i
        assert bit_equal(got["center_im"][i], gold["center_im"][i]), "iteration %d" % i
        assert np.array_equal(got["center_valid"][i].cpu().numpy(), gold["center_valid"][i])
        assert relerr(got["features"][i].cpu()[..., ::16], gold["features"][i]) <= FEAT_TOL
        same_cls = got["sem_cls_prob"][i].cpu().argmax(-1) == torch.from_numpy(gold["sem_cls_prob"][i]).argmax(-1)
        flips += int((~same_cls).sum())
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            assert relerr(got[k][i].cpu(), gold[k][i]) <= TOL, (k, i)
        # size = exp(s) * mean_size[argmax]: compare where the arg-max class agrees (a flip changes the
        # looked-up mean size discontinuously; flips are counted and bounded below)
        sz, gz = got["size_unnormalized"][i].cpu()[same_cls], torch.from_numpy(gold["size_unnormalized"][i])[same_cls]
        assert relerr(sz, gz) <= TOL, ("size_unnormalized", i)
    assert flips <= 2, "%d arg-max class flips" % flips


def test_decoder_c1_shape_against_oracle(dev):
    # BASELINE config 1 geometry: 1 clip, 8 views of 60x80 tokens, 256 queries, 8 iterations
    B, T, H, W, Nq, seed = 1, 8, 60, 80, 256, 21
    sd = I.make_weights(seed, Nq)
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    c = dict(tokens=tokens, camera=cam._data, T_cp=Tcp._data, T_wp=Twp._data, T_wl=Twl._data, H=H, W=W)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    outs, auxs = O.decoder_forward(tokens, cam._data, Tcp._data, Twp._data, Twl._data, sd, return_aux=True)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev)
    got = _engine_forward(eng, c, dev, forced_refs=refs.to(dev), debug=True)
    for i in range(8):
        assert bit_equal(got["center_im"][i], auxs[i]["center_im"]) and torch.equal(got["center_valid"][i].cpu(), auxs[i]["center_valid"])
        assert relerr(got["features"][i].cpu(), auxs[i]["features"]) <= FEAT_TOL
        assert relerr(got["decoder_out"][i].cpu(), auxs[i]["decoder_out"]) <= TOL
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            assert relerr(got[k][i].cpu(), outs[i][k]) <= TOL, (k, i)
    # rotation epilogue = compute_rotation_matrix_from_ortho6d of the emitted ortho6d
    R = O.rotation_from_ortho6d(got["ortho6d"][7].cpu().reshape(-1, 6)).view(B, Nq, 3, 3)
    assert relerr(got["rotation"][7].cpu(), R) <= 1e-5
    # free-running (no teacher forcing) stays close for the first iterations; later ones are reported only
    free = _engine_forward(eng, c, dev)
    assert relerr(free["center_unnormalized"][0].cpu(), outs[0]["center_unnormalized"]) <= TOL
    assert relerr(free["center_unnormalized"][1].cpu(), outs[1]["center_unnormalized"]) <= 5 * TOL


def test_module_forward_matches_engine_and_reference_api(dev):
    gold = load_golden("small")
    c = regenerate_case(gold)
    m = PARQDecoderB200(default_cfg(c["Nq"])).eval()
    m.load_state_dict(c["sd"], strict=True)
    m = m.to(dev)
    cam, Tcp, Twp, Twl = I.make_geometry(c["B"], c["T"], c["H"], c["W"], seed=c["seed"])
    out = m(c["tokens"].to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    torch.cuda.synchronize()
    assert isinstance(out, list) and len(out) == 8 and set(out[0].keys()) == set(OUT_KEYS)
    assert out[0]["pred_logits"].shape == (c["B"], c["Nq"], 10) and out[0]["ortho6d"].shape == (c["B"], c["Nq"], 6)
    for k in OUT_KEYS:   # iteration 0 consumes sigmoid(refpoint) computed on the device
        assert relerr(out[0][k].cpu(), gold[k][0]) <= TOL, k
    with pytest.raises(NotImplementedError):
        m.train()(c["tokens"].to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))


# ------------------------------------------------------------------ kernel unit tests ----
def _gemm(dev, A, Bw, M, N, K, **kw):
    lib = _lib.load()
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
    ak, bk = (C.c_int32 * 3)(*kw.get("a_koff", (0, 0, 0))), (C.c_int32 * 3)(*kw.get("b_koff", (0, 0, 0)))
    _lib.check(lib.parq_gemm_bf16(_ptr(A), A.shape[0], A.shape[1], _ptr(Bw), Bw.shape[0], Bw.shape[1], M, N, K, kw.get("nterms", 1),
                                  ak, bk, _ptr(kw.get("bias")), 0, kw.get("relu", 0), _ptr(out), N, None, 0, 0, 0, _stream()), "gemm")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 768, 1024), (4096, 1024, 384), (1, 16, 64)])
def test_gemm_against_fp32(dev, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(dev).bfloat16()
    Bw = torch.randn(N, K, generator=g).to(dev).bfloat16()
    assert relerr(_gemm(dev, A, Bw, M, N, K), A.float() @ Bw.float().t()) <= 1e-5


def test_gemm_three_term_split_recovers_fp32_product(dev):
    g = torch.Generator().manual_seed(5)
    M, N, K = 256, 512, 256
    x, W = torch.randn(M, K, generator=g).to(dev), (torch.randn(N, K, generator=g) * 0.05).to(dev)
    sp = lambda t: torch.cat([t.bfloat16(), (t - t.bfloat16().float()).bfloat16()], 1).contiguous()
    bias = torch.randn(N, generator=g).to(dev)
    out = _gemm(dev, sp(x), sp(W), M, N, K, nterms=3, a_koff=(0, K, 0), b_koff=(0, 0, K), bias=bias, relu=1)
    ref = torch.relu((x.double() @ W.double().t()).float() + bias)
    assert relerr(out, ref) <= 3e-5


@pytest.mark.parametrize("M,N,K", [(256, 1024, 1024), (256, 768, 1024), (256, 1024, 768), (200, 2048, 1024), (512, 1024, 256), (128, 2048, 512)])
def test_gemm_cluster_split_k_against_float64(dev, M, N, K):
    # gemm_sk.cuh: a GEMM of a few row tiles, activations [hi|lo] x bf16-exact weights -> a cluster of 4 CTAs per 128 x 128
    # tile splits K and reduce-scatters the partial accumulators over distributed shared memory (bias + ReLU epilogue)
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    xs = torch.cat([x.bfloat16(), (x - x.bfloat16().float()).bfloat16()], 1).contiguous()
    bias = torch.randn(N, generator=g).to(dev)
    Ws = torch.cat([W, torch.zeros_like(W)], 1).contiguous()          # packed [hi|lo] weights with an empty low part
    out = _gemm(dev, xs, Ws, M, N, K, nterms=2, a_koff=(0, K, 0), b_koff=(0, 0, 0), bias=bias, relu=1)
    x16 = xs[:, :K].double() + xs[:, K:].double()
    ref = torch.relu(x16 @ W.double().t() + bias.double()).float()
    assert relerr(out, ref) <= 2e-6


@pytest.mark.parametrize("M,K1,N2,w_lo", [(128, 1024, 1024, 0), (512, 768, 768, 0), (256, 1024, 2048, 1)])
def test_chained_kernel_linear_layernorm_linear(dev, M, K1, N2, w_lo):
    # chain_tc.cuh in isolation: cluster of 4 CTAs per 128 rows, LayerNorm statistics exchanged over DSMEM, the second stage
    # streams the first stage's output back as its A operand; against float64 torch on the same (split-representable) inputs
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + K1 + N2)
    sp = lambda t: torch.cat([t.bfloat16(), (t - t.bfloat16().float()).bfloat16()], 1).contiguous()
    rec = lambda s_: s_[:, : s_.shape[1] // 2].double() + s_[:, s_.shape[1] // 2:].double()
    x = torch.randn(M, K1, generator=g).to(dev)
    W1 = (torch.randn(1024, K1, generator=g) * 0.04).to(dev)
    W2 = (torch.randn(N2, 1024, generator=g) * 0.04).to(dev)
    if not w_lo:
        W1, W2 = W1.bfloat16().float(), W2.bfloat16().float()
    b1, b2 = torch.randn(1024, generator=g).to(dev), torch.randn(N2, generator=g).to(dev)
    gamma, beta = (1 + 0.1 * torch.randn(1024, generator=g)).to(dev), (0.1 * torch.randn(1024, generator=g)).to(dev)
    resid = torch.randn(M, 1024, generator=g).to(dev)
    xs, w1s, w2s = sp(x), sp(W1), sp(W2)
    y = torch.full((M, 1024), float("nan"), device=dev)
    ys = torch.zeros(M, 2048, dtype=torch.bfloat16, device=dev)
    zs = torch.zeros(M, 2 * N2, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.parq_chain_ln_linear(_ptr(xs), _ptr(w1s), _ptr(b1), _ptr(resid.t().contiguous()), _ptr(gamma), _ptr(beta), _ptr(w2s), _ptr(b2),
                                        M, K1, N2, w_lo, _ptr(y), _ptr(ys), _ptr(zs), _stream()), "parq_chain_ln_linear")
    torch.cuda.synchronize()
    w1 = rec(w1s) if w_lo else W1.double()
    z64 = rec(xs) @ w1.t() + b1.double() + resid.double()
    y64 = torch.nn.functional.layer_norm(z64, (1024,), gamma.double(), beta.double(), 1e-5)
    assert relerr(y, y64) <= 3e-5
    assert relerr(rec(ys), y64) <= 3e-5                        # the operand split carries 16 mantissa bits
    w2 = rec(w2s) if w_lo else W2.double()
    out64 = torch.relu(rec(ys) @ w2.t() + b2.double())         # second stage on the operand it actually consumed
    assert relerr(rec(zs), out64) <= 3e-5


def _attention(dev, B, H, Nq, Nk, fp16, nsplit, seed, scale=1.0, spike=False):
    lib = _lib.load()
    g = torch.Generator().manual_seed(seed)
    dt, Cc = (torch.float16 if fp16 else torch.bfloat16), H * 256
    Q = (torch.randn(B * Nq, Cc, generator=g) * scale / 16).to(dev).to(dt)
    K = torch.randn(B * Nk, Cc, generator=g).to(dev).to(dt)
    V = torch.randn(B * Nk, Cc, generator=g).to(dev).to(dt)
    if spike:
        K[Nk - 3::Nk] *= 12
    ldv = (B * Nk + 63) // 64 * 64
    Vt = torch.zeros(Cc, ldv, dtype=dt, device=dev)
    Vt[:, : B * Nk] = V.t()
    nb = lib.parq_attention_scratch_bytes(B, H, Nq, Nk)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    out = torch.zeros(B * Nq, 2 * Cc, dtype=torch.bfloat16, device=dev)
    _lib.check(lib.parq_attention(_ptr(Q), Cc, _ptr(K), Cc, _ptr(Vt), ldv, B, H, Nq, Nk, int(fp16), _ptr(scratch), nb, _ptr(out),
                                  nsplit, _stream()), "attention")
    torch.cuda.synchronize()
    got = (out[:, :Cc].float() + out[:, Cc:].float()).view(B, Nq, H, 256).permute(0, 2, 1, 3)
    hv = lambda t, n: t.float().view(B, n, H, 256).permute(0, 2, 1, 3).double()
    p = torch.softmax(torch.einsum("bhqd,bhkd->bhqk", hv(Q, Nq), hv(K, Nk)), -1)
    return relerr(got, torch.einsum("bhqk,bhkd->bhqd", p, hv(V, Nk)).float())


@pytest.mark.parametrize("B,H,Nq,Nk,fp16,nsplit", [(1, 1, 128, 128, False, 1), (1, 1, 128, 384, False, 3), (1, 2, 256, 420, False, 2),
                                                   (2, 4, 256, 1000, False, 0), (2, 4, 256, 256, True, 1), (1, 1, 128, 1, False, 1)])
def test_attention_against_fp32_softmax(dev, B, H, Nq, Nk, fp16, nsplit):
    assert _attention(dev, B, H, Nq, Nk, fp16, nsplit, seed=Nk) <= (2e-3 if fp16 else 1e-2)


@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 4, 256, 4000), (1, 2, 512, 1500), (3, 1, 256, 136), (16, 4, 256, 2048)])
def test_attention_stream_k_schedule(dev, B, H, Nq, Nk):
    # nsplit = -1 forces the stream-K schedule: segments that start / end in the middle of an item, items merged from
    # two or three partials, items written directly, ragged last key tiles
    assert _attention(dev, B, H, Nq, Nk, False, -1, seed=Nk + B) <= 1e-2


def test_attention_plain_layout_alignment_is_checked(dev):
    # stand-alone entry point with plain K / V^T matrices: clip b's keys start at column b*Nk of V^T, which TMA can only
    # address at 16-byte granularity -> refused with an error instead of faulting (the decoder's tiled caches have no
    # such restriction, see test_decoder_ragged_clip_sizes)
    with pytest.raises(_lib.ParqError, match="Nk"):
        _attention(dev, 3, 1, 256, 129, False, 0, seed=1)


def test_attention_lazy_rescale_path(dev):
    assert _attention(dev, 1, 2, 128, 1024, False, 1, seed=3, scale=4.0, spike=True) <= 1e-2


# ------------------------------------------------ benchmark-size properties (config 2) ----
def test_config2_properties(dev):
    B, T, H, W, Nq, seed = 16, 8, 60, 80, 256, 31
    sd = I.make_weights(seed, Nq)
    eng = DecoderEngine(sd, dev)
    tokens = I.make_tokens(B, T, H, W, seed=seed).to(dev).bfloat16()
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    g = torch.Generator().manual_seed(seed)
    refs = torch.rand(8, B, Nq, 3, generator=g).to(dev)          # fixed reference points: no recurrence
    args = lambda sl: (tokens[sl], cam._data[sl].to(dev), Tcp._data[sl].to(dev), Twp._data[sl].to(dev), Twl._data[sl].to(dev), H, W)
    full = eng.forward(*args(slice(0, B)), forced_refs=refs, debug=True)
    torch.cuda.synchronize()
    for k in OUT_KEYS:
        assert torch.isfinite(full[k]).all(), k
    assert (full["sem_cls_prob"].sum(-1) - 1).abs().max() <= 1e-5
    R = full["rotation"][-1].reshape(-1, 3, 3)
    assert (R.transpose(1, 2) @ R - torch.eye(3, device=dev)).abs().max() <= 1e-4
    # coord_pos is the exact denormalisation of the forced reference points
    lo, span = torch.tensor([-3.0, -2.0, 0.25], device=dev), torch.tensor([6.0, 2.5, 5.0], device=dev)
    assert torch.equal(full["coord_pos"], refs * span + lo)
    # clip independence: a clip decoded alone equals the same clip decoded inside the batch
    # (different key-split plans => only the split-combine order changes)
    full = {k: v.clone() for k, v in full.items()}
    for b in (0, 9):
        one = eng.forward(*args(slice(b, b + 1)), forced_refs=refs[:, b:b + 1].contiguous(), debug=True)
        torch.cuda.synchronize()
        assert torch.equal(one["center_im"][:, 0], full["center_im"][:, b])
        assert torch.equal(one["features"][:, 0], full["features"][:, b])
        for k in ("pred_logits", "center_unnormalized", "ortho6d"):
            assert relerr(one[k][:, 0], full[k][:, b]) <= 1e-4, (k, b)


def test_cuda_graph_replay_matches_eager(dev):
    gold = load_golden("small")
    c = regenerate_case(gold)
    eng = DecoderEngine(c["sd"], dev)
    args = [c[k].to(dev) for k in ("tokens", "camera", "T_cp", "T_wp", "T_wl")]
    args[0] = args[0].bfloat16()
    eager = eng.forward(*args, c["H"], c["W"])
    torch.cuda.synchronize()
    eager = {k: v.clone() for k, v in eager.items()}
    for _ in range(3):                       # capture on first use, then pure replays
        replay = eng.forward(*args, c["H"], c["W"], graph=True)
    torch.cuda.synchronize()
    for k in OUT_KEYS:
        assert torch.equal(replay[k], eager[k]), k
    assert len(eng._graphs) == 1
    # programmatic dependent launch only reorders prologues: results are bit-identical with it switched off
    plain = eng.forward(*args, c["H"], c["W"], pdl=False)
    torch.cuda.synchronize()
    for k in OUT_KEYS:
        assert torch.equal(plain[k], eager[k]), k


# ------------------------------------------------- f-2: parse_pred + NMS on the device ----
@pytest.mark.parametrize("name", ["small", "ragged_wild", "white_noise"])
def test_parse_pred_kernel_against_reference_golden(dev, name):
    # the reference's own last-iteration tensors through the device kernel: detection set, NMS decision, boxes
    from parq_b200.decoder import parse_pred
    gold = load_golden(name)
    last = {k: torch.from_numpy(gold[k][-1]).to(dev) for k in ("center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob")}
    got = parse_pred(last)
    torch.cuda.synchronize()
    assert np.array_equal(got["pred_mask"].cpu().numpy(), gold["pred_mask"])
    assert np.array_equal(got["nms_mask"].cpu().numpy(), gold["nms_mask"])
    obb = got["obbs_pred"]._data.cpu().numpy()
    assert np.array_equal(obb[..., 18], gold["obbs_pred"][..., 18])                 # labels
    assert np.abs(obb - gold["obbs_pred"]).max() <= 1e-6 * max(1.0, np.abs(gold["obbs_pred"]).max())


@pytest.mark.parametrize("name", ["small", "ragged_wild", "white_noise"])
def test_post_nms_detection_sets_identical(dev, name):
    # north_star acceptance: teacher-forced decoder on the GPU -> device parse_pred == the reference's detection set
    gold = load_golden(name)
    c = regenerate_case(gold)
    gold_outs = [{k: torch.from_numpy(gold[k][i]) for k in OUT_KEYS} for i in range(8)]
    refs = O.refs_from_outputs(gold_outs, c["sd"])
    m = PARQDecoderB200(default_cfg(c["Nq"])).eval()
    m.load_state_dict(c["sd"], strict=True)
    eng = DecoderEngine(c["sd"], dev)
    got = _engine_forward(eng, c, dev, forced_refs=refs.to(dev))
    parsed = m.parse_pred([{k: got[k][i] for k in OUT_KEYS} for i in range(8)])
    torch.cuda.synchronize()
    assert np.array_equal(parsed["pred_mask"].cpu().numpy(), gold["pred_mask"]), "post-NMS detection set differs from the reference"
    # The NMS decision BEFORE the scope filter (not an output of the reference) is reported, not gated at zero: greedy
    # NMS cascades, so one rank swap between two boxes whose scores differ by less than the 1e-3 parity tolerance
    # (bf16 attention) can flip a chain of out-of-scope boxes.  It stays a small fraction of the boxes.
    flips = int((parsed["nms_mask"].cpu().numpy() != gold["nms_mask"]).sum())
    print("nms_mask flips before the scope filter: %d of %d boxes" % (flips, gold["nms_mask"].size))
    assert flips <= 0.08 * gold["nms_mask"].size


@pytest.mark.parametrize("name", ["small", "ragged_wild", "white_noise"])
def test_parse_pred_for_vis_branch_against_reference_golden(dev, name):
    # FOR_VIS=True (parq_decoder.py:407-421, utils/nms.py:182-224): same-class NMS at IoU 0.2, no track-scale filter --
    # the reference's own tensors and the reference's own answer (make_golden.py, `pred_mask_vis`)
    from parq_b200.decoder import parse_pred
    gold = load_golden(name)
    last = {k: torch.from_numpy(gold[k][-1]).to(dev) for k in ("center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob")}
    got = parse_pred(last, for_vis=True)
    torch.cuda.synchronize()
    assert np.array_equal(got["pred_mask"].cpu().numpy(), gold["pred_mask_vis"])
    m = PARQDecoderB200(default_cfg(256))
    m.for_vis = True
    assert np.array_equal(m.parse_pred([last])["pred_mask"].cpu().numpy(), gold["pred_mask_vis"])
    with pytest.raises(NotImplementedError):
        parse_pred(last, enable_nms=False)


def test_parse_pred_random_boxes_against_oracle(dev):
    from parq_b200.decoder import parse_pred
    g = torch.Generator().manual_seed(77)
    for B, K, for_vis in ((16, 256, False), (3, 512, False), (2, 100, False), (16, 256, True), (2, 100, True)):
        last = {"center_unnormalized": (torch.rand(B, K, 3, generator=g) - 0.5) * torch.tensor([4.0, 2.0, 3.0]) + torch.tensor([0.0, 0.0, 1.2]),
                "size_unnormalized": torch.rand(B, K, 3, generator=g) * 1.2 + 0.2,
                "ortho6d": torch.randn(B, K, 6, generator=g),
                "sem_cls_prob": torch.softmax(3 * torch.randn(B, K, 10, generator=g), -1)}
        want = O.parse_pred(last, for_vis=for_vis)
        got = parse_pred({k: v.to(dev) for k, v in last.items()}, for_vis=for_vis)
        torch.cuda.synchronize()
        assert torch.equal(got["labels"].cpu(), want["labels"]) and torch.equal(got["scores"].cpu(), want["scores"])
        assert torch.equal(got["nms_mask"].cpu(), want["nms_mask"]), (B, K)
        assert torch.equal(got["pred_mask"].cpu(), want["pred_mask"]), (B, K)
        assert 0 < int(want["nms_mask"].sum()) < B * K


def test_kv_cache_tile_contiguous_layout(dev):
    # hoisted K / V^T projection into the tile-contiguous caches the cross-attention streams; Nk = 576 leaves the
    # last key tile of every clip half empty (its V^T entries must be zero, its K entries are masked by the kernel)
    B, T, H, W, Nq, seed = 2, 3, 12, 16, 128, 4
    sd = I.make_weights(seed, Nq)
    eng = DecoderEngine(sd, dev)
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    eng.kv_project(tokens.to(dev).bfloat16(), B, T, H, W)
    torch.cuda.synchronize()
    Nk, heads = T * H * W, 4
    assert eng.workspace_value("kv_tiled", B, T, H, W) == 1
    ntile = eng.workspace_value("ntile", B, T, H, W)
    assert ntile == (Nk + 127) // 128
    Kc = eng.workspace_view("Kc", B, T, H, W, torch.bfloat16, (B, ntile, heads, 128, 256)).float().cpu()
    Vt = eng.workspace_view("Vt", B, T, H, W, torch.bfloat16, (B, ntile, heads, 256, 128)).float().cpu()
    L = "parq_module.decoder.layers.0.multihead_attn."
    Wi, bi = sd[L + "in_proj_weight"], sd[L + "in_proj_bias"]
    Kref = tokens @ Wi[1024:2048].t() + bi[1024:2048]            # (B, Nk, C)
    Vref = tokens @ Wi[2048:].t() + bi[2048:]
    pad = ntile * 128 - Nk
    Kp = torch.nn.functional.pad(Kref, (0, 0, 0, pad)).view(B, ntile, 128, heads, 256).permute(0, 1, 3, 2, 4)
    Vp = torch.nn.functional.pad(Vref, (0, 0, 0, pad)).view(B, ntile, 128, heads, 256).permute(0, 1, 3, 4, 2)
    keys_ok = (torch.arange(ntile * 128) < Nk).view(1, ntile, 1, 128, 1)
    assert relerr(torch.where(keys_ok, Kc, torch.zeros(())), Kp) <= 5e-3        # bf16 storage
    assert relerr(Vt, Vp) <= 5e-3
    assert Vt.permute(0, 1, 2, 4, 3)[~keys_ok.expand(B, ntile, heads, 128, 1).squeeze(-1)].abs().max() == 0


# ------------------------------------------------ other BASELINE.json configurations ----
def test_config4_geometry_against_oracle(dev):
    # long-clip stress geometry (BASELINE config 4: 32 views, 512 queries) at a feature-map size the CPU oracle
    # finishes in seconds (30x40 -> Nk = 38 400 keys): teacher-forced, first three iterations
    B, T, H, W, Nq, seed, iters = 1, 32, 30, 40, 512, 41, 3
    sd = I.make_weights(seed, Nq)
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    c = dict(tokens=tokens, camera=cam._data, T_cp=Tcp._data, T_wp=Twp._data, T_wl=Twl._data, H=H, W=W)
    outs, auxs = O.decoder_forward(tokens, cam._data, Tcp._data, Twp._data, Twl._data, sd, iters=iters, return_aux=True)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev, iters=iters)
    got = _engine_forward(eng, c, dev, forced_refs=refs.to(dev), debug=True)
    for i in range(iters):
        assert bit_equal(got["center_im"][i], auxs[i]["center_im"]) and torch.equal(got["center_valid"][i].cpu(), auxs[i]["center_valid"])
        assert relerr(got["features"][i].cpu(), auxs[i]["features"]) <= FEAT_TOL
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            assert relerr(got[k][i].cpu(), outs[i][k]) <= TOL, (k, i)


def test_config4_full_size_properties(dev):
    # BASELINE config 4 at full size: 1 clip, 32 views of 120x160 tokens (614 400 keys), 512 queries, 8 iterations
    B, T, H, W, Nq, seed = 1, 32, 120, 160, 512, 43
    sd = I.make_weights(seed, Nq)
    eng = DecoderEngine(sd, dev)
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randn(B, T * H * W, 1024, generator=g).to(dev).bfloat16()
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    refs = torch.rand(8, B, Nq, 3, generator=g).to(dev)
    args = (tokens, cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev), Twl._data.to(dev), H, W)
    full = eng.forward(*args, forced_refs=refs, debug=True)
    torch.cuda.synchronize()
    for k in OUT_KEYS:
        assert torch.isfinite(full[k]).all(), k
    assert (full["sem_cls_prob"].sum(-1) - 1).abs().max() <= 1e-5
    # projection against the numpy oracle at full size (bit-exact), features on a channel subset
    Tcl = O.camera_from_local(Tcp._data.numpy(), Twp._data.numpy(), Twl._data.numpy())
    lo, span = torch.tensor([-3.0, -2.0, 0.25]), torch.tensor([6.0, 2.5, 5.0])
    pc = O.transform_points(Tcl, (refs[3].cpu() * span + lo).numpy())
    cim, val = O.pinhole_project(cam._data.numpy(), pc)
    assert bit_equal(full["center_im"][3], cim) and np.array_equal(full["center_valid"][3].cpu().numpy(), val)
    # key-split invariance: the same cross-attention with a different split plan (cached K / V^T) agrees
    a = full["decoder_out"].clone()
    again = eng.forward(*args, forced_refs=refs, debug=True, skip_kv=True)
    torch.cuda.synchronize()
    assert torch.equal(again["decoder_out"], a)              # deterministic replay on the cached K / V^T


def test_streaming_window_shape(dev):
    # BASELINE config 5 shape: one clip, 8-view window sliding by one view; every window is an independent forward
    B, T, H, W, Nq, seed = 1, 8, 60, 80, 256, 51
    sd = I.make_weights(seed, Nq)
    eng = DecoderEngine(sd, dev)
    stream = I.make_tokens(1, T + 2, H, W, seed=seed)[0].view(T + 2, H * W, 1024).to(dev).bfloat16()
    cam, Tcp, Twp, _ = I.make_geometry(1, T + 2, H, W, seed=seed)
    outs = []
    for s in range(3):
        tok = stream[s:s + T].reshape(1, T * H * W, 1024).contiguous()
        Twl = Twp._data[:, s + T // 2: s + T // 2 + 1]
        o = eng.forward(tok, cam._data[:, s:s + T].to(dev), Tcp._data[:, s:s + T].to(dev), Twp._data[:, s:s + T].to(dev), Twl.to(dev), H, W, graph=True)
        outs.append({k: v.clone() for k, v in o.items()})
    torch.cuda.synchronize()
    assert all(torch.isfinite(o["center_unnormalized"]).all() for o in outs)
    assert not torch.equal(outs[0]["center_unnormalized"], outs[1]["center_unnormalized"])


# ------------------------------------------- f-1: AddRayPE + tokeniser fused producer ----
@pytest.mark.parametrize("name", ["raype_small", "raype_c1_view"])
def test_add_ray_pe_against_reference_golden(dev, name):
    from parq_b200.raype import AddRayPEB200
    gold = load_golden(name)
    B, T, H, W, seed = [int(x) for x in gold["shape"]]
    sd = I.make_raype_weights(seed)
    m = AddRayPEB200(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    feat = I.make_features(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    enc = m(feat.to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    tokens = m.tokens(feat.to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    torch.cuda.synchronize()
    assert enc.shape == (B, T, 1024, H, W) and enc.dtype == torch.float32
    # forward(): both encoder layers consume exact [hi|lo] splits of their inputs -> the 1e-3 bar of the decoder
    assert relerr(enc.cpu()[:, :, ::16], gold["encoding"]) <= TOL
    enc_o, tok_o = O.add_ray_pe(feat, cam._data, Tcp._data, Twp._data, Twl._data, sd)
    assert relerr(enc.cpu(), enc_o) <= TOL
    # tokens(): plain-bf16 hidden layer, features + encoding rounded once to bf16 (the decoder's input precision):
    # within one bf16 ulp of the value plus the bf16-hidden error of the encoding (3e-3 of its range)
    assert tokens.shape == (B, T * H * W, 1024) and tokens.dtype == torch.bfloat16
    d = (tokens.float().cpu() - tok_o).abs()
    assert (d <= tok_o.abs() * 2.0 ** -8 + 3e-3 * enc_o.abs().max()).all()
    with pytest.raises(NotImplementedError):
        m(feat, cam, Tcp, Twp, Twl)                       # CPU tensors: refuse, never fall back


def test_decoder_with_fp32_weights_against_oracle(dev):
    # released checkpoints are not bf16-representable: parq_pack_weights then keeps a low-order weight term and the
    # GEMMs of the fp32 residual stream run their third term (A_hi x W_lo)
    B, T, H, W, Nq, seed, iters = 2, 3, 12, 16, 256, 61, 3
    sd = I.make_weights(seed, Nq, bf16_exact=False)
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    c = dict(tokens=tokens, camera=cam._data, T_cp=Tcp._data, T_wp=Twp._data, T_wl=Twl._data, H=H, W=W)
    outs, auxs = O.decoder_forward(tokens, cam._data, Tcp._data, Twp._data, Twl._data, sd, iters=iters, return_aux=True)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev, iters=iters)
    assert eng.weight_lo
    got = _engine_forward(eng, c, dev, forced_refs=refs.to(dev), debug=True, chain=True)     # three-term stages in the chained kernel
    errs = {}
    for i in range(iters):
        assert bit_equal(got["center_im"][i], auxs[i]["center_im"])
        errs[("decoder_out", i)] = relerr(got["decoder_out"][i].cpu(), auxs[i]["decoder_out"])
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            errs[(k, i)] = relerr(got[k][i].cpu(), outs[i][k])
    print({k: "%.2e" % v for k, v in errs.items()})
    assert max(errs.values()) <= TOL, max(errs, key=errs.get)


# --------------------------------------------------- f-3: FPN upsample + concat writer ----
@pytest.mark.parametrize("name", ["fpn_small", "fpn_odd"])
def test_fpn_concat_against_reference_golden(dev, name):
    from parq_b200.fpn import fpn_concat
    gold = load_golden(name)
    B, T, H, W, seed = [int(x) for x in gold["shape"]]
    pyr = I.make_pyramid(B * T, H, W, seed=seed)
    got = fpn_concat({k: v.to(dev) for k, v in pyr.items()})
    torch.cuda.synchronize()
    assert got.shape == (B * T, 1024, H, W)
    got = got.view(B, T, 1024, H, W).cpu()
    assert bit_equal(got[:, :, :256], pyr["0"].view(B, T, 256, H, W))                      # the target level is an exact copy
    assert relerr(got[:, :, ::16], gold["all_features"]) <= 1e-6
    assert relerr(got, O.fpn_concat(pyr).view(B, T, 1024, H, W)) <= 1e-6
    other = fpn_concat({k: v.to(dev) for k, v in pyr.items()}, layer=1).cpu()                # a coarser target level (down- and up-sampling)
    assert relerr(other, O.fpn_concat(pyr, layer=1)) <= 1e-6
    # bf16 output = the same arithmetic rounded once; bf16 levels in = the fp32 kernel on the rounded levels
    half = fpn_concat({k: v.to(dev) for k, v in pyr.items()}, out_dtype=torch.bfloat16)
    assert half.dtype == torch.bfloat16 and torch.equal(half.float().cpu(), got.reshape(B * T, 1024, H, W).bfloat16().float())
    pyr16 = {k: v.bfloat16() for k, v in pyr.items()}
    both = fpn_concat({k: v.to(dev) for k, v in pyr16.items()}, out_dtype=torch.bfloat16)
    want = fpn_concat({k: v.float().to(dev) for k, v in pyr16.items()})
    torch.cuda.synchronize()
    assert torch.equal(both.float(), want.bfloat16().float())


@pytest.mark.parametrize("feat_dtype", [torch.float32, torch.bfloat16], ids=["fp32-features", "bf16-features"])
def test_pipeline_fpn_raype_decoder_nms_chain(dev, feat_dtype):
    # the hot path with its three "next" rows chained on the device, as PARQ.forward / update_metrics would run them:
    # pyramid -> fpn_concat (f-3) -> AddRayPEB200.tokens (f-1) -> PARQDecoderB200 -> parse_pred (f-2).
    # bf16-features: all_features written and read back in bf16 (half the bytes of f-3; one more rounding before the sum)
    from parq_b200.fpn import camera_feature, fpn_concat
    from parq_b200.raype import AddRayPEB200
    B, T, H, W, Nq, seed = 1, 2, 12, 16, 128, 71
    pyr = I.make_pyramid(B * T, H, W, seed=seed)
    cam_img, Tcp, Twp, Twl = I.make_geometry(B, T, 4 * H, 4 * W, seed=seed)
    cam = camera_feature(cam_img)                                      # image camera -> feature-map camera (1/4)
    feats = fpn_concat({k: v.to(dev) for k, v in pyr.items()}, out_dtype=feat_dtype).view(B, T, 1024, H, W)
    rpe = AddRayPEB200(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
    rsd = I.make_raype_weights(seed)
    rpe.load_state_dict(rsd, strict=True)
    tokens = rpe.to(dev).tokens(feats, cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    dec = PARQDecoderB200(default_cfg(Nq)).eval()
    sd = I.make_weights(seed, Nq)
    dec.load_state_dict(sd, strict=True)
    outs = dec.to(dev)(tokens, cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    parsed = dec.parse_pred(outs)
    torch.cuda.synchronize()
    # oracle chain up to the tokens, then the oracle decoder on OUR bf16 tokens (identical inputs) for iteration 0
    feat_o = O.fpn_concat(pyr).view(B, T, 1024, H, W)
    enc_o, tok_o = O.add_ray_pe(feat_o, cam._data, Tcp._data, Twp._data, Twl._data, rsd)
    d = (tokens.float().cpu() - tok_o).abs()
    pre = feat_o.permute(0, 1, 3, 4, 2).reshape(B, T * H * W, 1024).abs() * (2.0 ** -9 if feat_dtype == torch.bfloat16 else 0.0)
    assert (d <= tok_o.abs() * 2.0 ** -8 + pre + 3e-3 * enc_o.abs().max()).all()
    ref = O.decoder_forward(tokens.float().cpu(), cam._data, Tcp._data, Twp._data, Twl._data, sd, iters=1)
    for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
        assert relerr(outs[0][k].cpu(), ref[0][k]) <= TOL, k
    want = O.parse_pred({k: v.cpu() for k, v in outs[-1].items()})
    assert torch.equal(parsed["pred_mask"].cpu(), want["pred_mask"]) and torch.equal(parsed["nms_mask"].cpu(), want["nms_mask"])


def test_accelerate_hook_runs_the_library(dev):
    # accelerate() patches an instance of the REFERENCE's own PARQDecoder class in place.  The class comes
    # from /root/reference in the build container and from its bytecode tree oracle/_ref on the GPU box; only when neither
    # is present does the module with the reference's attribute names stand in for it.
    from oracle import ref_loader as RL
    from parq_b200.decoder import accelerate
    gold = load_golden("small")
    c = regenerate_case(gold)
    cam, Tcp, Twp, Twl = I.make_geometry(c["B"], c["T"], c["H"], c["W"], seed=c["seed"])
    real = RL.reference_available()
    if real:
        ns = RL.load_reference()
        m = RL.build_decoder(c["sd"], c["Nq"], device=dev)
        assert type(m).__module__ == "model.parq_decoder" and type(m).__name__ == "PARQDecoder"
        geo = (ns.Camera(cam._data.to(dev)), ns.Pose(Tcp._data.to(dev)), ns.Pose(Twp._data.to(dev)), ns.Pose(Twl._data.to(dev)))
    else:
        m = PARQDecoderB200(default_cfg(c["Nq"])).eval()
        m.load_state_dict(c["sd"], strict=True)
        m = m.to(dev)
        geo = (cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    m = accelerate(m, use_cuda_graph=True)           # H, W read from the Camera like the reference does (transformer_parq.py:301)
    # fp32 tokens, as parq_lightning.py:78-88 hands them over (the fixture's values are bf16-representable)
    tok32 = c["tokens"].to(dev).float()
    for _ in range(3):
        out = m(tok32.clone(), *geo)                  # a fresh tensor every call: the graph cache must still hit
    torch.cuda.synchronize()
    assert len(out) == 8 and set(out[0].keys()) == set(OUT_KEYS)
    for k in OUT_KEYS:                                # iteration 0 is free of the recurrence: compare with the fixture
        assert relerr(out[0][k].cpu(), gold[k][0]) <= TOL, k
    before = _lib.load().parq_kernel_launches()
    m.forward.use_cuda_graph = False
    out = m(tok32, *geo)
    assert _lib.load().parq_kernel_launches() - before > 100          # the CUDA library did the work
    if real:
        # everything else stays the reference's: its own parse_pred consumes our outputs (device Obb3D round trip and numpy NMS)
        parsed = m.parse_pred([dict(o) for o in out])
        ours = PARQDecoderB200(default_cfg(c["Nq"])).parse_pred([dict(o) for o in out])
        assert np.array_equal(parsed["pred_mask"].cpu().numpy(), ours["pred_mask"].cpu().numpy())


def test_graph_cache_is_keyed_by_shape_not_by_temporaries(dev):
    # fp32 tokens / non-contiguous pose slices create fresh temporaries per call; the graph must be reused
    gold = load_golden("small")
    c = regenerate_case(gold)
    eng = DecoderEngine(c["sd"], dev)
    base = None
    for i in range(4):
        tok = c["tokens"].to(dev).float().clone()                     # new address every call, needs the bf16 conversion
        poses = torch.stack([c["T_cp"], c["T_wp"]], 0).to(dev)         # slices of a stacked tensor
        out = eng.forward(tok, c["camera"].to(dev), poses[0], poses[1], c["T_wl"].to(dev), c["H"], c["W"], graph=True)
        torch.cuda.synchronize()
        if base is None:
            base = {k: v.clone() for k, v in out.items()}
        for k in OUT_KEYS:
            assert torch.equal(out[k], base[k]), k
    assert len(eng._graphs) == 1
    eager = eng.forward(c["tokens"].to(dev), c["camera"].to(dev), c["T_cp"].to(dev), c["T_wp"].to(dev), c["T_wl"].to(dev), c["H"], c["W"])
    torch.cuda.synchronize()
    for k in OUT_KEYS:
        assert torch.equal(eager[k], base[k]), k


def test_custom_mean_size_table_is_honoured(dev, tmp_path):
    # MEAN_SIZE_PATH / box_processor.mean_size_arr must reach the kernels (size = exp(s) * mean_size[argmax])
    from parq_b200.decoder import load_mean_size
    gold = load_golden("small")
    c = regenerate_case(gold)
    path = tmp_path / "sizes.txt"
    names = ["chair", "table", "cabinet", "trash can,trash bin", "bookshelf", "display,video display", "sofa,couch", "bathtub,tub"]
    path.write_text("".join("%s: [%.8f %.8f %.8f] \n" % (n, 0.5 + 0.1 * i, 1.0 + 0.2 * i, 0.25 * (i + 1)) for i, n in enumerate(names)))
    table = load_mean_size(str(path))
    assert table.shape == (10, 3) and table[3, 0].item() == pytest.approx(0.8) and table[8:].eq(1).all()
    cfg = default_cfg(c["Nq"])
    cfg.MEAN_SIZE_PATH = str(path)
    m = PARQDecoderB200(cfg).eval()
    m.load_state_dict(c["sd"], strict=True)
    m = m.to(dev)
    cam, Tcp, Twp, Twl = I.make_geometry(c["B"], c["T"], c["H"], c["W"], seed=c["seed"])
    out = m(c["tokens"].to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    torch.cuda.synchronize()
    cls = torch.from_numpy(gold["sem_cls_prob"][0]).argmax(-1)
    same = out[0]["sem_cls_prob"].cpu().argmax(-1) == cls
    want = torch.from_numpy(gold["size_unnormalized"][0]) / torch.tensor(O.MEAN_SIZE, dtype=torch.float32)[cls] * table.float()[cls]
    assert relerr(out[0]["size_unnormalized"].cpu()[same], want[same]) <= TOL


def test_decoder_ragged_clip_sizes(dev):
    # Nk = T*H*W = 105 keys per clip, two clips: no alignment of any kind (key tiles, clips and 32-token chunks all
    # straddle); the K / V^T projection takes the per-element path of the GEMM epilogue
    B, T, H, W, Nq, seed, iters = 2, 3, 5, 7, 128, 81, 2
    sd = I.make_weights(seed, Nq)
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    c = dict(tokens=tokens, camera=cam._data, T_cp=Tcp._data, T_wp=Twp._data, T_wl=Twl._data, H=H, W=W)
    outs, auxs = O.decoder_forward(tokens, cam._data, Tcp._data, Twp._data, Twl._data, sd, iters=iters, return_aux=True)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev, iters=iters)
    got = _engine_forward(eng, c, dev, forced_refs=refs.to(dev), debug=True)
    for i in range(iters):
        assert bit_equal(got["center_im"][i], auxs[i]["center_im"])
        assert relerr(got["features"][i].cpu(), auxs[i]["features"]) <= FEAT_TOL
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            assert relerr(got[k][i].cpu(), outs[i][k]) <= TOL, (k, i)


def test_unchained_launch_sequence_matches_chained(dev):
    # chain=False: one GEMM / LayerNorm launch per link (gemm_tc.cuh + add_ln_kernel) instead of the chained cluster kernel
    # (chain_tc.cuh).  Both against the fixture; against each other they differ only by the summation order of the
    # LayerNorm statistics (and a 1-ulp query change can flip a bf16 rounding inside the attention).
    gold = load_golden("small")
    c = regenerate_case(gold)
    gold_outs = [{k: torch.from_numpy(gold[k][i]) for k in OUT_KEYS} for i in range(8)]
    refs = O.refs_from_outputs(gold_outs, c["sd"]).to(dev)
    eng = DecoderEngine(c["sd"], dev)
    chained = {k: v.clone() for k, v in _engine_forward(eng, c, dev, forced_refs=refs, debug=True, chain=True).items()}
    plain = _engine_forward(eng, c, dev, forced_refs=refs, debug=True, chain=False)
    for i in range(8):
        assert relerr(plain["decoder_out"][i], chained["decoder_out"][i]) <= 2e-4
        for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
            assert relerr(plain[k][i].cpu(), gold[k][i]) <= TOL, (k, i)
            assert relerr(chained[k][i].cpu(), gold[k][i]) <= TOL, (k, i)
    # fp32 (checkpoint-style) weights: three-term GEMMs, plain 4-slot ring in the chained kernel
    sd = I.make_weights(61, 256, bf16_exact=False)
    eng = DecoderEngine(sd, dev, iters=2)
    a = {k: v.clone() for k, v in _engine_forward(eng, c, dev, debug=True, chain=True).items()}
    b = _engine_forward(eng, c, dev, debug=True, chain=False)
    assert relerr(a["decoder_out"][0], b["decoder_out"][0]) <= 2e-4
    assert relerr(a["pred_logits"][0], b["pred_logits"][0]) <= 2e-4


def test_hi_only_mask_on_both_launch_paths(dev):
    # opt-in "high-order activation term only" for the three GEMMs whose output is rounded to 16 bits (mask 7: self-attention
    # Q|K, V^T, cross-attention Q; include/parq_b200.h PARQ_FLAG_HI_ONLY_*): takes effect on the chained and the un-chained
    # path and stays inside the 1e-3 bar of the fixture (the default keeps every term: DESIGN.md 4.0)
    gold = load_golden("small")
    c = regenerate_case(gold)
    gold_outs = [{k: torch.from_numpy(gold[k][i]) for k in OUT_KEYS} for i in range(8)]
    refs = O.refs_from_outputs(gold_outs, c["sd"]).to(dev)
    eng = DecoderEngine(c["sd"], dev)
    for chain in (True, False):
        full = {k: v.clone() for k, v in _engine_forward(eng, c, dev, forced_refs=refs, chain=chain, hi_only=0).items()}
        part = _engine_forward(eng, c, dev, forced_refs=refs, chain=chain, hi_only=7)
        assert not torch.equal(full["pred_logits"], part["pred_logits"]), "the mask had no effect (chain=%s)" % chain
        for i in range(8):
            for k in ("pred_logits", "center_unnormalized", "ortho6d"):
                assert relerr(part[k][i].cpu(), gold[k][i]) <= TOL, (k, i, chain)
                assert relerr(full[k][i].cpu(), gold[k][i]) <= relerr(part[k][i].cpu(), gold[k][i]) + 1e-4, (k, i, chain)


def test_side_stream_branches_do_not_change_results(dev):
    # un-chained launch path (one clip): the reference-point MLP runs next to the sampler and V^T next to the Q|K projection on
    # side streams of the library (parq_api.cu SideStreams); the same kernels on one stream (fork=False) give the same bits, in
    # eager mode and inside a captured graph (parallel branches), free-running
    gold = load_golden("small")
    c = regenerate_case(gold)
    eng = DecoderEngine(c["sd"], dev)
    serial = {k: v.clone() for k, v in _engine_forward(eng, c, dev, chain=False, fork=False).items()}
    forked = {k: v.clone() for k, v in _engine_forward(eng, c, dev, chain=False).items()}
    graphed = _engine_forward(eng, c, dev, chain=False, graph=True)
    for k in OUT_KEYS:
        assert torch.equal(serial[k], forked[k]), k
        assert torch.equal(serial[k], graphed[k]), k


def test_launch_trace_counts_every_kernel(dev):
    # parq_trace: one stamp per kernel launch of the library (thread 0 of block 0 at the moment its dependency resolved)
    gold = load_golden("small")
    c = regenerate_case(gold)
    eng = DecoderEngine(c["sd"], dev)
    _engine_forward(eng, c, dev)
    lib = _lib.load()
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    assert lib.parq_trace(buf.data_ptr(), 4096) == 0
    n0 = int(lib.parq_kernel_launches())
    _engine_forward(eng, c, dev)
    n1 = int(lib.parq_kernel_launches())
    assert lib.parq_trace(None, 0) == 0
    h = buf.cpu()
    assert 100 < int(h[0]) <= n1 - n0, (int(h[0]), n1 - n0)       # kernels without a dependency wait do not stamp
    st = h[1:1 + int(h[0])]
    assert bool((st > 0).all()) and int(st.max() - st.min()) < 1_000_000_000      # nanoseconds of one small forward


def test_unshared_decoder_layers_against_reference_golden(dev):
    # SHARE_WEIGHTS False: one distinct layer per iteration, driven as single-iteration library calls (K / V^T re-projected
    # by every layer, as the reference does); module API included
    gold = load_golden("unshared")
    B, T, H, W, Nq, seed, layers = [int(x) for x in gold["shape"]]
    sd = I.make_weights(seed, Nq, n_layers=layers)
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    c = dict(tokens=tokens, camera=cam._data, T_cp=Tcp._data, T_wp=Twp._data, T_wl=Twl._data, H=H, W=W)
    gold_outs = [{k: torch.from_numpy(gold[k][i]) for k in OUT_KEYS} for i in range(layers)]
    refs = O.refs_from_outputs(gold_outs, sd)
    eng = DecoderEngine(sd, dev, iters=layers)
    assert eng.n_layers == layers
    for graph in (False, True):
        got = _engine_forward(eng, c, dev, forced_refs=refs.to(dev), graph=graph)
        for i in range(layers):
            assert bit_equal(got["coord_pos"][i], gold["coord_pos"][i])
            for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob"):
                assert relerr(got[k][i].cpu(), gold[k][i]) <= TOL, (k, i, graph)
    cfg = default_cfg(Nq, layers)
    cfg.TRANSFORMER.SHARE_WEIGHTS = False
    m = PARQDecoderB200(cfg).eval()
    m.load_state_dict(sd, strict=True)
    out = m.to(dev)(tokens.to(dev), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev))
    torch.cuda.synchronize()
    for k in OUT_KEYS:                                   # free running: iteration 0 is exact, iteration 1 starts from our own centres
        assert relerr(out[0][k].cpu(), gold[k][0]) <= TOL, k
    assert relerr(out[1]["center_unnormalized"].cpu(), gold["center_unnormalized"][1]) <= 5 * TOL


def test_free_running_equals_teacher_forcing_with_its_own_points(dev):
    # The free-running forward takes shortcuts the teacher-forced one does not (the heads kernel writes the next iteration's
    # sinusoidal embedding, the sampler projects after instead of before its dependency wait): feeding the free run's own
    # reference points back as forced points must reproduce it bit for bit.
    gold = load_golden("small")
    c = regenerate_case(gold)
    eng = DecoderEngine(c["sd"], dev)
    for chain in (None, True):
        free = {k: v.clone() for k, v in _engine_forward(eng, c, dev, chain=chain).items()}
        outs = [{k: free[k][i].cpu() for k in OUT_KEYS} for i in range(8)]
        refs = O.refs_from_outputs(outs, c["sd"])
        forced = _engine_forward(eng, c, dev, forced_refs=refs.to(dev), chain=chain)
        for k in OUT_KEYS:
            assert torch.equal(forced[k], free[k]), (k, chain)


def test_free_running_divergence_report(dev):
    """the free-running recurrence (no teacher forcing) is reported, not gated, next to the oracle's own
    sensitivity: the oracle re-run on tokens perturbed by a relative 1e-6 (fp32 rounding scale) and 2^-9 (bf16 operand
    rounding scale, what the tensor-core attention does to Q/K/V).  Written to gpurun_out/free_running.md."""
    import os
    B, T, H, W, Nq, seed = 1, 8, 30, 40, 256, 33
    sd = I.make_weights(seed, Nq)
    tokens = I.make_tokens(B, T, H, W, seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed)
    c = dict(tokens=tokens, camera=cam._data, T_cp=Tcp._data, T_wp=Twp._data, T_wl=Twl._data, H=H, W=W)
    run = lambda tk: O.decoder_forward(tk, cam._data, Tcp._data, Twp._data, Twl._data, sd)
    base = run(tokens)
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(tokens.shape, generator=g)
    pert = {"1e-6": run(tokens.float() * (1 + 1e-6 * noise)), "2^-9": run(tokens.float() * (1 + 2.0 ** -9 * noise))}
    free = _engine_forward(DecoderEngine(sd, dev), c, dev)
    keys = ("center_unnormalized", "pred_logits")
    lines = ["| iteration | " + " | ".join("GPU free-running %s" % k for k in keys) + " | " +
             " | ".join("oracle, tokens x(1+%s n) %s" % (p, k) for p in pert for k in keys) + " |",
             "|---|" + "---|" * (len(keys) * (1 + len(pert)))]
    for i in range(8):
        row = [relerr(free[k][i].cpu(), base[i][k]) for k in keys]
        row += [relerr(pert[p][i][k], base[i][k]) for p in pert for k in keys]
        assert all(np.isfinite(row))
        lines.append("| %d | " % i + " | ".join("%.2e" % v for v in row) + " |")
    assert relerr(free["center_unnormalized"][0].cpu(), base[0]["center_unnormalized"]) <= TOL
    report = "\n".join(lines)
    print("\n" + report)
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/free_running.md", "w") as f:
            f.write("max|d|/max|ref| per iteration vs the fp32 oracle, 1 clip x 8 views x 30x40, 256 queries (tests/test_gpu_parity.py::"
                    "test_free_running_divergence_report)\n\n" + report + "\n")
    except OSError:
        pass
