"""Where a decoder step spends its time on the DEPENDENT LAUNCH CHAIN (graph replay, PDL on, power cap and all): every kernel
stamps the global timer when its stream dependency resolves (parq_trace); the difference of consecutive stamps is what each
launch costs end to end (execution + drain + hand-over to the next launch).
    python tools/launch_trace.py [clips] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import _lib, inputs as I
from parq_b200.decoder import DecoderEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T, H, W, Nq = 8, 60, 80, 256
dev = torch.device("cuda:0")
eng = DecoderEngine(I.make_weights(0, Nq), dev)
tokens = torch.empty(B, T * H * W, 1024, dtype=torch.bfloat16, device=dev)
for b in range(B):
    tokens[b] = I.make_tokens(1, T, H, W, seed=1000 + b)[0].to(dev)
cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=2000)
args = (tokens, cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev), Twl._data.to(dev), H, W)
lib = _lib.load()
cap = 1 << 16
buf = torch.zeros(cap, dtype=torch.int64, device=dev)
assert lib.parq_trace(buf.data_ptr(), cap) == 0      # before the graph is captured: the chain launches bake their stamp slots
for _ in range(5):
    eng.forward(*args, graph=True)
torch.cuda.synchronize()
buf[0] = 0
torch.cuda.synchronize()
for _ in range(reps):
    eng.forward(*args, graph=True)
torch.cuda.synchronize()
lib.parq_trace(None, 0)
h = buf.cpu()
n = int(h[0])
st = h[1:1 + n].double()
per = n // reps
print("stamps %d = %d per step" % (n, per))
d = (st[1:] - st[:-1]) / 1e3                      # us
steps = [d[r * per:(r + 1) * per - 1] for r in range(reps) if (r + 1) * per - 1 <= len(d)]
m = torch.stack(steps[1:]).mean(0) if len(steps) > 1 else steps[0]
names = ["posemb", "sample", "chain P", "self-attn", "chain A", "cross-attn", "combine", "chain B", "gn_apply", "gemm hd2", "heads"]
if (per - 3) % 8 == 0 and (per - 3) // 8 == 10:      # fused stream-K merge: no combine launch
    names.remove("combine")
NL = len(names)
head = per - NL * 8                                # launches before the first iteration (pose chain, K / V^T projection, ...)
print("prologue launches (us): " + " ".join("%.1f" % x for x in m[:head]))
if (per - head) % 8 == 0 and (per - head) // 8 == len(names):
    # stamp k marks the START (dependency resolved) of launch k: launch k costs stamp[k+1] - stamp[k]
    it = torch.cat([m[head:], m.new_zeros(1)])[:8 * NL].reshape(8, NL)
    print("per launch, mean over iterations 1..6 (us):")
    for j, nm in enumerate(names):
        print("  %-10s %7.1f" % (nm, it[1:7, j].mean()))
    print("  iteration  %7.1f" % it[1:7].sum(1).mean())
else:
    print("per-launch deltas (us):", " ".join("%.1f" % x for x in m))
print("step (first stamp to first stamp of the next replay): %.1f us" % float(((st[per::per] - st[:-per:per]) / 1e3).mean()))

# per-CTA stamps of the chain launches of the last replay: entry / dependency resolved / exit, relative to the launch's trace stamp
ch = h[cap // 2:][:(cap // 2) // 640 * 640].reshape(-1, 160, 4).double()
nct = 4 * (B * Nq // 128)
# the graph's launches carry the slots handed out while it was captured: they are the 24 slots with the latest stamps
order = sorted(range(ch.shape[0]), key=lambda i: float(ch[i, 0, 0]))[-24:]
ch = ch[order]
last = st[(reps - 1) * per:(reps) * per]
names_c = ["P", "A", "B"]
for it in (2, 5):
    for j, k in enumerate((names.index("chain P"), names.index("chain A"), names.index("chain B"))):
        slot = it * 3 + j
        t_dep = float(last[head + it * NL + k])
        t_next = float(last[head + it * NL + k + 1])
        c = ch[slot, :nct]
        e, w, x = (c[:, 0] - t_dep) / 1e3, (c[:, 1] - t_dep) / 1e3, (c[:, 2] - t_dep) / 1e3
        mhz = (c[:, 3] / (c[:, 2] - c[:, 0]) * 1e3).median()
        print("SM clock inside the kernel: %.0f MHz" % mhz)
        print("iteration %d chain %s: CTA entry %.1f .. %.1f us | wait returned %.1f .. %.1f | exit %.1f .. %.1f (median %.1f) | next launch's dependency resolved %.1f"
              % (it, names_c[j], e.min(), e.max(), w.min(), w.max(), x.min(), x.max(), x.median(), (t_next - t_dep) / 1e3))
