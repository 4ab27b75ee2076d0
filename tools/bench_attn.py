"""Cross-attention kernel timing through parq_attention: (B, H) = (16, 4) against (64, 1) -- the latter reads a
fully contiguous K (512-byte rows at 512-byte pitch), which isolates the effect of the K layout on DRAM efficiency."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import _lib
from parq_b200.decoder import _ptr, _stream

dev = torch.device("cuda:0")
lib = _lib.load()


def run(B, H, Nq, Nk, nsplit=0, reps=12):
    Cc = H * 256
    Q = (torch.randn(B * Nq, Cc, device=dev) / 16).bfloat16()
    K = torch.randn(B * Nk, Cc, device=dev).bfloat16()
    ldv = (B * Nk + 63) // 64 * 64
    Vt = torch.randn(Cc, ldv, device=dev).bfloat16()
    nb = lib.parq_attention_scratch_bytes(B, H, Nq, Nk)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    out = torch.zeros(B * Nq, 2 * Cc, dtype=torch.bfloat16, device=dev)
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.parq_attention(_ptr(Q), Cc, _ptr(K), Cc, _ptr(Vt), ldv, B, H, Nq, Nk, 0, _ptr(scratch), nb, _ptr(out), nsplit, _stream()), "attn")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    fl = 4.0 * B * Nq * Nk * Cc
    print("B=%d H=%d Nq=%d Nk=%d nsplit=%d: median %.1f us (incl. combine) -> %.0f TFLOP/s, K+V %.2f GB -> %.2f TB/s" % (
        B, H, Nq, Nk, nsplit, ts[len(ts) // 2], fl / ts[len(ts) // 2] / 1e6, 4.0 * B * Nk * Cc / 1e9, 4.0 * B * Nk * Cc / ts[len(ts) // 2] / 1e6), flush=True)


run(16, 4, 256, 38400)
run(64, 1, 256, 38400)
run(16, 4, 256, 38400, nsplit=8)
run(64, 1, 256, 38400, nsplit=8)
