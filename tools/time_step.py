"""Device-resident step time at BASELINE config 2 (graph replay), PDL on/off, plus the per-class event breakdown.
    python tools/time_step.py [steps] [clips] [views] [H] [W] [Nq]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import _lib, inputs as I
from parq_b200.decoder import DecoderEngine

a = [int(x) for x in sys.argv[1:]]
steps = a[0] if len(a) > 0 else 20
B = a[1] if len(a) > 1 else 16
T = a[2] if len(a) > 2 else 8
H = a[3] if len(a) > 3 else 60
W = a[4] if len(a) > 4 else 80
Nq = a[5] if len(a) > 5 else 256
dev = torch.device("cuda:0")
eng = DecoderEngine(I.make_weights(0, Nq, bf16_exact=not os.environ.get('PARQ_FP32_WEIGHTS')), dev)
print('weight_lo', eng.weight_lo)
tokens = torch.empty(B, T * H * W, 1024, dtype=torch.bfloat16, device=dev)
for b in range(B):
    tokens[b] = I.make_tokens(1, T, H, W, seed=1000 + b)[0].to(dev)
cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=2000)
args = (tokens, cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev), Twl._data.to(dev), H, W)


def timed(**kw):
    for _ in range(3):
        eng.forward(*args, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.forward(*args, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for rep in range(3):
    print("graph+pdl %.3f ms | graph no-pdl %.3f ms | eager+pdl %.3f ms || iterations only (K/V cached): pdl %.3f | no-pdl %.3f" % (
        timed(graph=True, pdl=True), timed(graph=True, pdl=False), timed(graph=False, pdl=True),
        timed(graph=True, pdl=True, skip_kv=True), timed(graph=True, pdl=False, skip_kv=True)), flush=True)
_lib.profile_enable(list(_lib.PROFILE_TAGS), 2048)
eng.forward(*args)
torch.cuda.synchronize()
print({k: (round(v[0], 4), v[1]) for k, v in _lib.profile_collect().items() if k != "_dropped"})
print("clips/s (graph+pdl): %.1f" % (B / timed(graph=True, pdl=True) * 1e3))
