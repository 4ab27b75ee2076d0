"""Secondary BASELINE.json configurations (parity-test cases, not the headline bench line), measured on one B200:
  C4  long-clip stress: 1 clip, 32 views of 120x160 tokens (614 400 keys), 512 queries, 8 iterations
  C5  streaming shape: 1 clip, sliding 8-view window (60x80), 256 queries -- per-window latency p50 / p99
Prints one JSON line per configuration."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import _lib, inputs as I
from parq_b200.decoder import DecoderEngine

dev = torch.device("cuda:0")


def c4(steps=5):
    B, T, H, W, Nq = 1, 32, 120, 160, 512
    eng = DecoderEngine(I.make_weights(0, Nq), dev)
    g = torch.Generator().manual_seed(0)
    tokens = torch.randn(B, T * H * W, 1024, generator=g).to(dev).bfloat16()
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=0)
    args = (tokens, cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev), Twl._data.to(dev), H, W)
    for _ in range(3):
        eng.forward(*args, graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.forward(*args, graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    _lib.profile_enable(list(_lib.PROFILE_TAGS), 1024)
    eng.forward(*args)
    torch.cuda.synchronize()
    br = {k: round(v[0], 3) for k, v in _lib.profile_collect().items() if k != "_dropped"}
    Nk = T * H * W
    gf = (4.0 * Nk * 1024 * 1024 + 8 * 4.0 * Nq * Nk * 1024) / 1e9
    print(json.dumps({"config": "C4 long-clip stress: 1 clip x 32 views x 120x160, 512 queries, 8 iterations", "ms_per_clip": ms,
                      "clips_per_s": 1e3 / ms, "attention_plus_kv_GF": gf, "TFLOPs_on_those": gf / ms, "breakdown_ms": br,
                      "cross_attn_TFLOPs": 8 * 4.0 * Nq * Nk * 1024 / br["cross_attn"] / 1e9}), flush=True)


def c5(windows=200):
    T, H, W, Nq = 8, 60, 80, 256
    eng = DecoderEngine(I.make_weights(0, Nq), dev)
    n = T + 8
    stream = I.make_tokens(1, n, H, W, seed=5)[0].view(n, H * W, 1024).to(dev).bfloat16()
    cam, Tcp, Twp, _ = I.make_geometry(1, n, H, W, seed=5)
    cam, Tcp, Twp = cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev)
    tok = torch.empty(1, T * H * W, 1024, dtype=torch.bfloat16, device=dev)          # static window buffers -> one captured graph
    bc, bcp, bwp, bwl = (torch.empty(1, T, 6, device=dev), torch.empty(1, T, 12, device=dev), torch.empty(1, T, 12, device=dev),
                         torch.empty(1, 1, 12, device=dev))
    lat = []
    for i in range(windows + 10):
        s = i % 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tok.copy_(stream[s:s + T].reshape(1, T * H * W, 1024))                       # the window slides by one view
        bc.copy_(cam[:, s:s + T]); bcp.copy_(Tcp[:, s:s + T]); bwp.copy_(Twp[:, s:s + T]); bwl.copy_(Twp[:, s + T // 2: s + T // 2 + 1])
        out = eng.forward(tok, bc, bcp, bwp, bwl, H, W, graph=True)
        e1.record()
        torch.cuda.synchronize()
        if i >= 10:
            lat.append(e0.elapsed_time(e1))
    lat.sort()
    print(json.dumps({"config": "C5 streaming shape: 1 clip, sliding 8-view window 60x80, 256 queries, 8 iterations; window copy + K/V "
                                "projection + decoder per window, CUDA-graph replay", "windows": windows, "p50_ms": statistics.median(lat),
                      "p99_ms": lat[int(0.99 * len(lat)) - 1], "min_ms": lat[0], "max_ms": lat[-1]}), flush=True)


def c5_streaming(windows=200):
    """C5 through the streaming window cache (parq_b200/streaming.py, f-4): per step ONE new view is copied in and projected
    (K / V^T of 4 800 tokens instead of 38 400), then the 8 iterations run over the cached window."""
    from parq_b200.streaming import StreamingWindow
    T, H, W, Nq = 8, 60, 80, 256
    eng = DecoderEngine(I.make_weights(0, Nq), dev)
    n = T + 8
    stream = I.make_tokens(1, n, H, W, seed=5)[0].view(n, H * W, 1024).to(dev).bfloat16()
    cam, Tcp, Twp, _ = I.make_geometry(1, n, H, W, seed=5)
    cam, Tcp, Twp = cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev)
    sw = StreamingWindow(eng, T, H, W)
    for v in range(T - 1):
        sw.push(stream[v:v + 1], cam[:, v], Tcp[:, v], Twp[:, v])
    lat, lat_push = [], []
    for i in range(windows + 10):
        v = (T - 1 + i) % n
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        sw.push(stream[v:v + 1], cam[:, v], Tcp[:, v], Twp[:, v])               # the window slides by one view
        e1.record()
        out = sw.decode(Twp[:, (v - T // 2) % n].reshape(1, 1, 12))              # local frame = pseudo-camera of the window's middle view
        e2.record()
        torch.cuda.synchronize()
        if i >= 10:
            lat.append(e0.elapsed_time(e2))
            lat_push.append(e0.elapsed_time(e1))
    lat.sort()
    print(json.dumps({"config": "C5 streaming shape through the window cache (f-4): 1 clip, sliding 8-view window 60x80, 256 queries, 8 iterations; "
                                "per window: copy + K/V projection of ONE view, then the decoder over the cached K/V as one CUDA-graph replay",
                      "windows": windows, "p50_ms": statistics.median(lat), "p99_ms": lat[int(0.99 * len(lat)) - 1], "min_ms": lat[0], "max_ms": lat[-1],
                      "push_p50_ms": statistics.median(lat_push)}), flush=True)


def raype(steps=10):
    """f-1 producer at config-2 size: 16 clips x 8 views x 60x80 pixels -> bf16 tokens."""
    from parq_b200.raype import AddRayPEB200
    B, T, H, W = 16, 8, 60, 80
    m = AddRayPEB200(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
    m.load_state_dict(I.make_raype_weights(0), strict=True)
    m = m.to(dev)
    feat = torch.randn(B, T, 1024, H, W, device=dev)
    cam, Tcp, Twp, Twl = (t.to(dev) for t in I.make_geometry(B, T, H, W, seed=0))
    for fn, tag in ((m.tokens, "tokens (bf16 hidden)"), (m.forward, "encoding (split hidden)")):
        for _ in range(3):
            fn(feat, cam, Tcp, Twp, Twl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(feat, cam, Tcp, Twp, Twl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ntok = B * T * H * W
        gf = 2.0 * ntok * 1024 * (2 * 192 + (1024 if "tokens" in tag else 2048)) / 1e9
        print(json.dumps({"config": "f-1 AddRayPE producer, 16 clips x 8 views x 60x80 -> " + tag, "ms_per_batch": ms, "clips_per_s": B / ms * 1e3,
                          "tensor_GF": gf, "TFLOPs": gf / ms}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4", "c5", "raype"]
    if "raype" in which:
        raype()
    if "c4" in which:
        c4()
    if "c5" in which:
        c5()
        c5_streaming()
