"""One-clip (config 5) latency per window with the one-clip switches of the library turned off one at a time; every row in its
own process (the switches are read when the library loads).
    python tools/ablate_c5.py            -> markdown table on stdout
    python tools/ablate_c5.py --child    -> one JSON line for the current environment"""
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ROWS = [
    ("default", {}),
    ("no side-stream branches (`PARQ_NO_FORK`)", {"PARQ_NO_FORK": "1"}),
    ("no cluster split-K GEMM (`PARQ_NO_SPLITK`): 128 x 64 single-CTA tiles", {"PARQ_NO_SPLITK": "1"}),
    ("no narrow tiles either (`PARQ_NO_SPLITK` + `PARQ_NO_NARROW`): 128 x 256 tiles", {"PARQ_NO_SPLITK": "1", "PARQ_NO_NARROW": "1"}),
    ("all three off", {"PARQ_NO_FORK": "1", "PARQ_NO_SPLITK": "1", "PARQ_NO_NARROW": "1"}),
]


def child():
    import torch
    from parq_b200 import inputs as I
    from parq_b200.decoder import DecoderEngine
    from parq_b200.streaming import StreamingWindow
    dev = torch.device("cuda:0")
    T, H, W, Nq = 8, 60, 80, 256
    eng = DecoderEngine(I.make_weights(0, Nq), dev)
    n = T + 8
    stream = I.make_tokens(1, n, H, W, seed=5)[0].view(n, H * W, 1024).to(dev).bfloat16()
    cam, Tcp, Twp, _ = I.make_geometry(1, n, H, W, seed=5)
    cam, Tcp, Twp = cam._data.to(dev), Tcp._data.to(dev), Twp._data.to(dev)
    sw = StreamingWindow(eng, T, H, W)
    for v in range(T - 1):
        sw.push(stream[v:v + 1], cam[:, v], Tcp[:, v], Twp[:, v])
    lat = []
    for i in range(110):
        v = (T - 1 + i) % n
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sw.push(stream[v:v + 1], cam[:, v], Tcp[:, v], Twp[:, v])
        sw.decode(Twp[:, (v - T // 2) % n].reshape(1, 1, 12))
        e1.record()
        torch.cuda.synchronize()
        if i >= 10:
            lat.append(e0.elapsed_time(e1))
    lat.sort()
    print(json.dumps({"p50_ms": statistics.median(lat), "p99_ms": lat[int(0.99 * len(lat)) - 1]}))


def main():
    print("| configuration | p50 ms / window | p99 |")
    print("|---|---|---|")
    for name, env in ROWS:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print("| %s | %.3f | %.3f |" % (name, d["p50_ms"], d["p99_ms"]), flush=True)
        except Exception:
            print("| %s | failed | %s |" % (name, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""), flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
