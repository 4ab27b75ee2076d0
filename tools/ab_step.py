"""In-process A/B timing of decoder-step variants at BASELINE config 2 (graph replay): the variants are run ALTERNATELY in short
blocks, so box, thermal state and the power-cap controller are shared and only the paired differences are read.  With --err the
teacher-forced parity error of every variant against the reference (GPU fp32, TF32 off; config-1 geometry x 2 clips) is printed too.
    python tools/ab_step.py [--err] [--rounds 12] [--steps 10] name=hi_only_mask[,chain=0|1] ...
    e.g. python tools/ab_step.py --err base=0 qkv=7 qkvpe=31"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from parq_b200 import inputs as I
from parq_b200.decoder import DecoderEngine

argv = sys.argv[1:]
want_err = "--err" in argv
rounds = int(argv[argv.index("--rounds") + 1]) if "--rounds" in argv else 12
steps = int(argv[argv.index("--steps") + 1]) if "--steps" in argv else 10
variants = []
for a in argv:
    if "=" in a and not a.startswith("--"):
        name, spec = a.split("=", 1)
        parts = spec.split(",")
        kw = {"hi_only": int(parts[0], 0)}
        for q in parts[1:]:
            k, v = q.split("=")
            kw[k] = bool(int(v))
        variants.append((name, kw))
if not variants:
    variants = [("base", {"hi_only": 0}), ("qkv", {"hi_only": 7})]
dev = torch.device("cuda:0")
T, H, W, Nq = 8, 60, 80, 256

if want_err:
    from conftest import relerr
    from oracle import parq_oracle as O
    from test_gpu_fullsize import reference_outputs
    B, seed = 2, 21
    sd = I.make_weights(seed, Nq)
    tokens = torch.cat([I.make_tokens(1, T, H, W, seed=seed * 10 + b) for b in range(B)])
    cam, Tcp, Twp, Twl = (t._data for t in I.make_geometry(B, T, H, W, seed=seed))
    outs, how = reference_outputs(sd, Nq, tokens, cam, Tcp, Twp, Twl, dev)
    refs = O.refs_from_outputs(outs, sd)
    eng = DecoderEngine(sd, dev)
    print("parity error (max over 8 teacher-forced iterations, max|d|/max|ref|), checker: %s" % how)
    for name, kw in variants:
        got = eng.forward(tokens.to(dev).bfloat16(), cam.to(dev), Tcp.to(dev), Twp.to(dev), Twl.to(dev), H, W, forced_refs=refs.to(dev),
                          chain=True, **{k: v for k, v in kw.items() if k != "chain"})
        torch.cuda.synchronize()
        err = {k: max(relerr(got[k][i].cpu(), outs[i][k]) for i in range(8))
               for k in ("pred_logits", "center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob")}
        print("  %-10s " % name + "  ".join("%s %.2e" % (k.split("_")[0], v) for k, v in err.items()), flush=True)
    del eng
    torch.cuda.empty_cache()

B = 16
eng = DecoderEngine(I.make_weights(0, Nq), dev)
tok = torch.empty(B, T * H * W, 1024, dtype=torch.bfloat16, device=dev)
for b in range(B):
    tok[b] = I.make_tokens(1, T, H, W, seed=1000 + b)[0].to(dev)
geo = [t._data.to(dev) for t in I.make_geometry(B, T, H, W, seed=2000)]
for name, kw in variants:
    for _ in range(3):
        eng.forward(tok, *geo, H, W, graph=True, **kw)
torch.cuda.synchronize()
# heat the chip into its power-capped steady state before anything is read
for _ in range(60):
    eng.forward(tok, *geo, H, W, graph=True, **variants[0][1])
torch.cuda.synchronize()
times = {name: [] for name, _ in variants}
for r in range(rounds):
    order = variants if r % 2 == 0 else variants[::-1]
    for name, kw in order:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.forward(tok, *geo, H, W, graph=True, **kw)
        e1.record()
        torch.cuda.synchronize()
        times[name].append(e0.elapsed_time(e1) / steps)
base = torch.tensor(times[variants[0][0]])
print("ms per step, %d alternating blocks of %d steps (mean, min; paired difference to '%s' with its standard error)" % (rounds, steps, variants[0][0]))
for name, _ in variants:
    t = torch.tensor(times[name])
    d = t - base
    print("  %-10s %.3f  %.3f   %+.3f +- %.3f" % (name, t.mean(), t.min(), d.mean(), d.std() / rounds ** 0.5))
