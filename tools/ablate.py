"""A/B table of the library's environment switches: teacher-forced parity error against the reference (GPU fp32, TF32 off) at
config-1 geometry x 2 clips, and ms per decoder step at config 2 (graph replay).  Every row runs in its own process (the switches
are read when the library loads).
    python tools/ablate.py                 -> markdown table on stdout
    python tools/ablate.py --child         -> one JSON line for the current environment"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ROWS = [
    ("default", {}),
    ("hi-only: sa_qk", {"PARQ_HI_ONLY": "1"}),
    ("hi-only: sa_v", {"PARQ_HI_ONLY": "2"}),
    ("hi-only: ca_q", {"PARQ_HI_ONLY": "4"}),
    ("hi-only: sa_qk + sa_v + ca_q", {"PARQ_HI_ONLY": "7"}),
    ("self-attention V^T as its own GEMM (not stage 0 of chain P)", {"PARQ_NO_CHAIN_V": "1"}),
    ("no chain (separate GEMM + LayerNorm launches)", {"PARQ_NO_CHAIN": "1"}),
]


def child():
    import torch
    from conftest import relerr
    from oracle import parq_oracle as O
    from parq_b200 import inputs as I
    from parq_b200.decoder import DecoderEngine
    from test_gpu_fullsize import reference_outputs
    dev = torch.device("cuda:0")
    B, T, H, W, Nq, seed = 2, 8, 60, 80, 256, 21
    sd = I.make_weights(seed, Nq)
    tokens = torch.cat([I.make_tokens(1, T, H, W, seed=seed * 10 + b) for b in range(B)])
    cam, Tcp, Twp, Twl = (t._data for t in I.make_geometry(B, T, H, W, seed=seed))
    outs, how = reference_outputs(sd, Nq, tokens, cam, Tcp, Twp, Twl, dev)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev)
    got = eng.forward(tokens.to(dev).bfloat16(), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev))
    torch.cuda.synchronize()
    err = {k: max(relerr(got[k][i].cpu(), outs[i][k]) for i in range(8)) for k in ("pred_logits", "center_unnormalized", "ortho6d", "sem_cls_prob")}
    del eng
    # timing at config 2
    B = 16
    eng = DecoderEngine(I.make_weights(0, Nq), dev)
    g = torch.Generator().manual_seed(0)
    tok = torch.randn(B, T * H * W, 1024, generator=g).to(dev).bfloat16()
    geo = [t._data.to(dev) for t in I.make_geometry(B, T, H, W, seed=0)]
    for _ in range(5):
        eng.forward(tok, *geo, H, W, graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.forward(tok, *geo, H, W, graph=True)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"err": err, "ms_per_step": e0.elapsed_time(e1) / 20, "checker": how}))


def main():
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    rows = [r for r in ROWS if not only or any(o in r[0] for o in only)]
    print("| configuration | logits | centre | ortho6d | probabilities | ms / step (config 2) |\n|---|---|---|---|---|---|")
    for name, env in rows:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print("| %s | %.2e | %.2e | %.2e | %.2e | %.3f |" % (name, d["err"]["pred_logits"], d["err"]["center_unnormalized"], d["err"]["ortho6d"],
                                                                d["err"]["sem_cls_prob"], d["ms_per_step"]), flush=True)
        except Exception:
            print("| %s | failed: %s |" % (name, (r.stderr or r.stdout)[-300:].replace("\n", " ")), flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
