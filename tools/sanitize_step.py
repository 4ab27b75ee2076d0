"""Small end-to-end pass over every round-2 kernel path for compute-sanitizer (memcheck / synccheck / racecheck):
chained and un-chained decoder (bf16 and fp32-pair tokens, fp32 weights), streaming window push + decode, FPN concat in bf16,
AddRayPE tokens from bf16 features, parse_pred.
    compute-sanitizer --tool memcheck python tools/sanitize_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import inputs as I
from parq_b200.decoder import DecoderEngine, parse_pred
from parq_b200.fpn import fpn_concat
from parq_b200.raype import AddRayPEB200
from parq_b200.streaming import StreamingWindow

dev = torch.device("cuda:0")
B, T, H, W, Nq = 2, 3, 8, 16, 256
for exact in (True, False):
    eng = DecoderEngine(I.make_weights(0, Nq, bf16_exact=exact), dev, iters=2)
    tok = I.make_tokens(B, T, H, W, seed=1)
    geo = [t._data.to(dev) for t in I.make_geometry(B, T, H, W, seed=1)]
    for chain in (True, False):
        out = eng.forward(tok.to(dev).bfloat16(), *geo, H, W, chain=chain, debug=True)
        out = eng.forward(tok.to(dev) * 1.001, *geo, H, W, chain=chain)          # fp32 tokens -> (hi, lo) pair
    parse_pred({k: v[-1] for k, v in out.items()})
eng = DecoderEngine(I.make_weights(0, Nq), dev, iters=2)
sw = StreamingWindow(eng, T, H, W)
tok1 = I.make_tokens(1, T + 1, H, W, seed=2)[0].view(T + 1, H * W, 1024).to(dev).bfloat16()
cam, Tcp, Twp, _ = (t._data.to(dev) for t in I.make_geometry(1, T + 1, H, W, seed=2))
for v in range(T + 1):
    sw.push(tok1[v:v + 1], cam[:, v], Tcp[:, v], Twp[:, v])
sw.decode(Twp[:, 1:2], graph=False)
pyr = {k: v.to(dev).bfloat16() for k, v in I.make_pyramid(B * T, H, W, seed=3).items()}
feats = fpn_concat(pyr, out_dtype=torch.bfloat16).view(B, T, 1024, H, W)
rpe = AddRayPEB200(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
rpe.load_state_dict(I.make_raype_weights(0), strict=True)
rpe.to(dev).tokens(feats, *[t.to(dev) for t in I.make_geometry(B, T, H, W, seed=3)])
torch.cuda.synchronize()
print("sanitize pass done")
