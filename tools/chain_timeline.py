"""Timeline of the chained GEMM kernel (chain_tc.cuh): clock64 stamps of CTA 0 of the three chain launches of one iteration.
    python tools/chain_timeline.py [clips]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import _lib, inputs as I
from parq_b200.decoder import DecoderEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T, H, W, Nq = 8, 60, 80, 256
dev = torch.device("cuda:0")
eng = DecoderEngine(I.make_weights(0, Nq), dev, iters=2)
g = torch.Generator().manual_seed(0)
tokens = torch.randn(B, T * H * W, 1024, generator=g).to(dev).bfloat16()
cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=0)
args = (tokens, cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev), Twl._data.to(dev), H, W)
eng.forward(*args)
torch.cuda.synchronize()
buf = torch.zeros(64, 64, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.parq_chain_debug(buf.data_ptr())
eng.forward(*args)
torch.cuda.synchronize()
lib.parq_chain_debug(None)
names = ["A ready", "1st operands", "last MMA issued", "acc complete", "LN pass1 done", "LN partials here", "epilogue done"]
buf = buf.cpu()
for launch in range(6):
    row = buf[launch]
    t0 = int(row[row > 0].min()) if (row > 0).any() else 0
    print("chain launch %d (%s)" % (launch, "PAB"[launch % 3]))
    for s in range(4):
        st = row[s * 8: s * 8 + 8]
        if not (st > 0).any():
            continue
        print("  stage %d: " % s + "  ".join("%s %+d" % (names[i], int(st[i]) - t0) for i in range(7) if int(st[i]) > 0))
        fine = row[32 + s * 8: 32 + s * 8 + 8]
        if (fine > 0).any():
            fn = ["p2 c0 start", "tmem ld done", "normalised", "out_cm issued", "out_f32 stored", "a_out stored", "p1 c1 start", "p1 c1 tmem st"]
            print("     fine: " + "  ".join("%s %+d" % (fn[i], int(fine[i]) - t0) for i in range(8) if int(fine[i]) > 0))
