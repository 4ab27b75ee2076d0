"""Micro-benchmark of parq_project_sample (stand-alone entry point, features only) at config-2 size:
normal geometry, all-miss geometry (no texel traffic) and an L2-resident token map, to separate the
random-gather DRAM cost from the fixed per-query cost."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from parq_b200 import _lib, inputs as I
from parq_b200.decoder import _ptr, _stream, make_shape

dev = torch.device("cuda:0")
lib = _lib.load()
B, T, H, W, Nq, Cc = 16, 8, 60, 80, 256, 1024


def run(tag, tokens, ref, Tcl, cam, B, T, H, W, reps=50):
    shape = make_shape(B, T, H, W, Cc, Nq, 4, 768, 1, 10, (-3, 3, -2, 0.5, 0.25, 5.25))
    feat = torch.empty(B, Nq, Cc, device=dev)
    cim = torch.empty(B, T, Nq, 2, device=dev)
    val = torch.empty(B, T, Nq, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.parq_project_sample(C.byref(shape), _ptr(tokens), None, _ptr(ref), _ptr(Tcl), _ptr(cam), _ptr(feat), _ptr(cim), _ptr(val),
                                           None, _stream()), "ps")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    nin = float(val.float().mean())
    print("%-28s median %.1f us  min %.1f us  valid %.3f" % (tag, ts[len(ts) // 2], ts[0], nin), flush=True)


tokens = torch.randn(B, T * H * W, Cc, device=dev).bfloat16()
cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=2000)
from parq_b200.decoder import pose_chain
Tcl = pose_chain(Tcp._data.to(dev), Twp._data.to(dev), Twl._data.to(dev))
g = torch.Generator().manual_seed(0)
ref = torch.rand(B, Nq, 3, generator=g).to(dev)
run("normal (random refs)", tokens, ref, Tcl, cam._data.to(dev), B, T, H, W)
ref_c = (0.5 + 0.02 * torch.randn(B, Nq, 3, generator=g)).clamp(0, 1).to(dev)
run("clustered refs", tokens, ref_c, Tcl, cam._data.to(dev), B, T, H, W)
Tmiss = Tcl.clone()
Tmiss[..., 11] -= 100.0          # everything far behind the cameras
run("all-miss", tokens, ref, Tmiss, cam._data.to(dev), B, T, H, W)
Tid = torch.tensor([1., 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], device=dev).expand(B, T, 12).contiguous()
refz = ref.clone()
refz[..., 0] = 0.5 + (refz[..., 0] - 0.5) * 0.15
refz[..., 1] = 0.8 + (refz[..., 1] - 0.5) * 0.15
run("all-hit (8 views x 4 corners)", tokens, refz, Tid, cam._data.to(dev), B, T, H, W)
