"""One decoder step at BASELINE config 2 (16 clips x 8 views x 60x80, 256 queries, 8 iterations) for ncu.
    ncu ... python tools/prof_step.py [iters] [clips]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import inputs as I
from parq_b200.decoder import DecoderEngine

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T, H, W, Nq = 8, 60, 80, 256
dev = torch.device("cuda:0")
eng = DecoderEngine(I.make_weights(0, Nq), dev, iters=iters)
g = torch.Generator().manual_seed(0)
tokens = torch.randn(B, T * H * W, 1024, generator=g).to(dev).bfloat16()
cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=0)
out = eng.forward(tokens, cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev), Twl._data.to(dev), H, W)
torch.cuda.synchronize()
print("ok", float(out["center_unnormalized"].abs().max()))
