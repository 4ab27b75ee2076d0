"""Stress test of CUDA-graph capture on the un-chained launch path (side-stream branches inside a capture): re-captures the
forward many times per variant and counts failed captures.
    python tools/stress_capture.py [n]"""
import gc
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import inputs as I
from parq_b200.decoder import DecoderEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
B, T, H, W, Nq = 1, 2, 12, 16, 128
eng = DecoderEngine(I.make_weights(0, Nq), dev)
tok = I.make_tokens(B, T, H, W, seed=0).to(dev)
geo = [t._data.to(dev) for t in I.make_geometry(B, T, H, W, seed=0)]
for name, kw in (("fork + pdl", {}), ("fork, no pdl", {"pdl": False}), ("no fork, pdl", {"fork": False}), ("fork + pdl, gc off", {"gcoff": True})):
    fails = 0
    gcoff = kw.pop("gcoff", False)
    if gcoff:
        gc.disable()
    for i in range(n):
        eng._graphs.clear()
        try:
            eng.forward(tok, *geo, H, W, chain=False, graph=True, **kw)
            torch.cuda.synchronize()
        except Exception as e:
            fails += 1
            if fails == 1:
                print("   first failure:", repr(e)[:200])
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
    if gcoff:
        gc.enable()
    print("%-22s %d / %d captures failed" % (name, fails, n), flush=True)
