"""Launch trace of a decoder forward (include/parq_b200.h: parq_trace).

While a trace buffer is set, thread 0 of block 0 of every kernel of the library appends the global timer at the moment its
stream dependency resolved; the difference of consecutive stamps is what each launch costs on the dependent chain of a graph
replay (execution + drain + hand-over), in the power-capped steady state.  The chained kernel additionally stamps entry /
dependency / exit and the elapsed ``clock64`` of every CTA (slots baked into the launch parameters, so the graph has to be
captured while the trace is on): SM clock inside the kernel = cycles / nanoseconds."""
import torch

from . import _lib

CHAINED_ITERATION = ["posemb", "sample", "chain P", "self-attn", "chain A", "cross-attn", "combine", "chain B", "gn_apply", "gemm hd2", "heads"]


def launch_trace(engine, forward, reps=8, cap=1 << 16, iters=8):
    """``forward()``: one forward through ``engine`` with graph=True (its cached graph is dropped and re-captured with the trace on).
    Returns a dict: stamps per step, mean cost per launch in us (over replays 2..), step time, SM clock inside the chained kernel."""
    lib = _lib.load()
    dev = engine.device
    buf = torch.zeros(cap, dtype=torch.int64, device=dev)
    _lib.check(lib.parq_trace(buf.data_ptr(), cap), "parq_trace")
    try:
        engine._graphs.clear()                      # re-capture: the chain launches bake their stamp slots
        for _ in range(3):
            forward()
        torch.cuda.synchronize(dev)
        buf[0] = 0
        torch.cuda.synchronize(dev)
        for _ in range(reps):
            forward()
        torch.cuda.synchronize(dev)
    finally:
        lib.parq_trace(None, 0)
        engine._graphs.clear()                      # the next capture is a plain one again
    h = buf.cpu()
    n = int(h[0])
    per = n // reps
    st = h[1:1 + n].double()
    d = (st[1:] - st[:-1]) / 1e3
    steps = [d[r * per:(r + 1) * per - 1] for r in range(reps) if (r + 1) * per - 1 <= len(d)]
    mean = torch.stack(steps[1:]).mean(0) if len(steps) > 1 else steps[0]
    out = {"stamps_per_step": per, "per_launch_us": [round(float(x), 2) for x in mean],
           "step_us": round(float(((st[per::per] - st[:-per:per]) / 1e3).mean()), 1), "stamps": st, "raw": h, "cap": cap, "reps": reps}
    names = list(CHAINED_ITERATION)
    if (per - 3) % iters == 0 and (per - 3) // iters == len(names) - 1:
        names.remove("combine")                     # fused stream-K merge: no combine launch
    head = per - len(names) * iters
    if head >= 0 and (per - head) // iters == len(names) and head <= 4:
        it = torch.cat([mean[head:], mean.new_zeros(1)])[:iters * len(names)].reshape(iters, len(names))
        mid = it[1:iters - 1]
        out["iteration_us"] = round(float(mid.sum(1).mean()), 1)
        out["iteration_launches_us"] = {nm: round(float(mid[:, j].mean()), 1) for j, nm in enumerate(names)}
        out["prologue_us"] = [round(float(x), 1) for x in mean[:head]]
        out["names"], out["head"] = names, head
    # SM clock inside the chained kernel: per-CTA (entry ns, dependency ns, exit ns, cycles) blocks in the upper half of the buffer
    ch = h[cap // 2:][:(cap // 2) // 640 * 640].reshape(-1, 160, 4).double()
    used = ch[:, 0, 0] >= (float(st[0]) if n > 0 else 0.0)      # launches of the measured replays only (the graph's slots)
    if n > 0 and bool(used.any()):
        c = ch[used][:, :128]
        ok = (c[..., 2] > c[..., 0]) & (c[..., 3] > 0)
        if bool(ok.any()):
            out["sm_mhz_in_chain_kernel"] = round(float((c[..., 3][ok] / (c[..., 2][ok] - c[..., 0][ok]) * 1e3).median()), 0)
    return out
