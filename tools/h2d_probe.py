"""Pinned host -> device copy rate of the e2e token buffer (1.26 GB), one stream vs. chunked over several streams:
the PCIe ceiling the end-to-end number of bench.py sits on."""
import torch, time
dev = torch.device("cuda:0")
n = 1258291200
host = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
def run(chunks, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sz = n // chunks
        evs = []
        for c in range(chunks):
            s = ss[c % streams]
            s.wait_event(e0)
            with torch.cuda.stream(s):
                d[c * sz:(c + 1) * sz].copy_(host[c * sz:(c + 1) * sz], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s); evs.append(ev)
        for ev in evs:
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("chunks %d streams %d: %.2f ms -> %.1f GB/s" % (chunks, streams, best, n / best / 1e6), flush=True)
run(1, 1); run(4, 2); run(8, 2); run(8, 4); run(16, 1)
