"""Where a decoder step spends its time on the DEPENDENT LAUNCH CHAIN (graph replay, PDL on, power cap and all): every kernel
stamps the global timer when its stream dependency resolves (parq_trace, parq_b200/tracing.py); the difference of consecutive
stamps is what each launch costs end to end (execution + drain + hand-over to the next launch).
    python tools/launch_trace.py [clips] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import inputs as I
from parq_b200.decoder import DecoderEngine
from parq_b200.tracing import launch_trace

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
for _ in range(5):
    eng.forward(*args, graph=True)
tr = launch_trace(eng, lambda: eng.forward(*args, graph=True), reps=reps)
per = tr["stamps_per_step"]
print("stamps %d per step" % per)
if "iteration_launches_us" in tr:
    print("prologue launches (us): " + " ".join("%.1f" % x for x in tr["prologue_us"]))
    print("per launch, mean over iterations 1..6 (us):")
    for nm, v in tr["iteration_launches_us"].items():
        print("  %-10s %7.1f" % (nm, v))
    print("  iteration  %7.1f" % tr["iteration_us"])
else:
    print("per-launch deltas (us):", " ".join("%.1f" % x for x in tr["per_launch_us"]))
print("step (first stamp to first stamp of the next replay): %.1f us" % tr["step_us"])
if "sm_mhz_in_chain_kernel" in tr:
    print("SM clock inside the chained kernel (median over its CTAs): %.0f MHz" % tr["sm_mhz_in_chain_kernel"])
if "names" in tr:
    # per-CTA stamps of the chain launches of the last replay: entry / dependency resolved / exit, relative to the launch's trace stamp
    names, head, cap, h, st = tr["names"], tr["head"], tr["cap"], tr["raw"], tr["stamps"]
    NL = len(names)
    ch = h[cap // 2:][:(cap // 2) // 640 * 640].reshape(-1, 160, 4).double()
    nct = 4 * (B * Nq // 128)
    order = sorted(range(ch.shape[0]), key=lambda i: float(ch[i, 0, 0]))[-24:]      # the graph's 24 chain launches: latest stamps
    ch = ch[order]
    last = st[(reps - 1) * per:reps * per]
    for it in (2, 5):
        for j, k in enumerate((names.index("chain P"), names.index("chain A"), names.index("chain B"))):
            t_dep, t_next = float(last[head + it * NL + k]), float(last[head + it * NL + k + 1])
            c = ch[it * 3 + j, :nct]
            e, w, x = (c[:, 0] - t_dep) / 1e3, (c[:, 1] - t_dep) / 1e3, (c[:, 2] - t_dep) / 1e3
            print("iteration %d chain %s: CTA entry %.1f .. %.1f us | wait returned %.1f .. %.1f | exit %.1f .. %.1f (median %.1f) | next launch's dependency resolved %.1f | %.0f MHz"
                  % (it, "PAB"[j], e.min(), e.max(), w.min(), w.max(), x.min(), x.max(), x.median(), (t_next - t_dep) / 1e3,
                     (c[:, 3] / (c[:, 2] - c[:, 0]) * 1e3).median()))
