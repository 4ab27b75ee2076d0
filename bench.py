#!/usr/bin/env python
"""Benchmark of the PARQ decoder hot path (BASELINE.json config 2) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the whole decoder (hoisted K/V projection + 8 recurrent iterations) over one batch of
16 synthetic clips per GPU (8 views of 60x80 tokens x 1024 channels, 256 queries).  Prints ONE JSON line.
`--impl reference` times the UNMODIFIED reference PARQDecoder.forward (from /root/reference, or its bytecode tree
oracle/_ref on the GPU box) on the host cores for a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "decoder clips/sec"
UNIT = "clips/s"
CFG = dict(clips_per_gpu=16, views=8, H=60, W=80, C=1024, queries=256, iterations=8, heads=4, ffn=768)


def workload_config(n_gpus):
    c = dict(CFG)
    c["workload"] = ("PARQ decoder-only bf16: batch 16 clips x 8 views per GPU, 256 queries, 8 iterations, "
                     "precomputed synthetic FPN features (BASELINE.json configs[1])")
    c["global_clips"] = CFG["clips_per_gpu"] * n_gpus
    c["parallelism"] = "clips sharded over %d GPU(s), no collective on the hot path" % n_gpus
    c["l2"] = "inputs larger than L2: 1.26 GB tokens + 2.5 GB K/V streamed every step (L2 = 126 MB); no explicit flush"
    return c


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm --
# The UNMODIFIED reference (model/parq_decoder.py:30 PARQDecoder, imported by oracle/ref_loader.py from /root/reference
# or, on the GPU box, from the bytecode tree oracle/_ref that oracle/build_ref.py compiled from it) on the host cores.
def reference_clip_inputs(B=1, seed=100):
    from parq_b200 import inputs as I
    sd = I.make_weights(0, CFG["queries"])
    tokens = I.make_tokens(B, CFG["views"], CFG["H"], CFG["W"], seed=seed)
    cam, Tcp, Twp, Twl = I.make_geometry(B, CFG["views"], CFG["H"], CFG["W"], seed=seed)
    return sd, tokens, cam._data, Tcp._data, Twp._data, Twl._data


def reference_runner(dec_layers, device="cpu", B=1):
    """Returns (callable running one PARQDecoder.forward over B clips, kind).  kind = "reference" when the real module
    is importable, else "port" (oracle/parq_oracle.py in the reference's op order)."""
    sd, tokens, cam, Tcp, Twp, Twl = reference_clip_inputs(B)
    try:
        from oracle import ref_loader as RL
        ns = RL.load_reference()
        m = RL.build_decoder(sd, CFG["queries"], dec_layers, device=device)
        args = (tokens.to(device), ns.Camera(cam.to(device)), ns.Pose(Tcp.to(device)), ns.Pose(Twp.to(device)), ns.Pose(Twl.to(device)))

        def run():
            with torch.no_grad():
                return m(*args)
        return run, "reference"
    except Exception as e:                                   # no reference tree: fall back to the port, and say so
        print("reference module unavailable (%s): timing the oracle port" % e, file=sys.stderr)
        from oracle import parq_oracle as O

        def run():
            return O.decoder_forward(tokens, cam, Tcp, Twp, Twl, sd, iters=dec_layers, hoist_kv=False)
        return run, "port"


def time_cpu_reference(threads):
    """PARQDecoder.forward, fp32, eval, no_grad, all host threads, config-1 shape (1 clip, 8 views of
    60x80 tokens, 256 queries, 8 iterations); 1 warm-up + 3 timed runs, median.  Returns (seconds per clip, kind)."""
    torch.set_num_threads(threads)
    run, kind = reference_runner(CFG["iterations"])
    run()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), kind


def time_c1_pipeline(threads):
    """BASELINE.json configs[0]: random-init torchvision ResNet50-FPN + the reference's AddRayPE + PARQDecoder on the host
    cores for 1 clip of 8 views 240x320 (model/resnet_fpn.py:16-91, ray_positional_encoding.py:61, parq_lightning.py:68-88).
    The backbone is built with pretrained=False (no network).  1 warm-up + 3 timed, median seconds per clip, or None."""
    try:
        from oracle import ref_loader as RL
        from parq_b200 import inputs as I
        from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
        import einops
        ns = RL.load_reference()
        torch.set_num_threads(threads)
        T, H, W = CFG["views"], CFG["H"], CFG["W"]
        torch.manual_seed(0)
        try:
            body = resnet_fpn_backbone(backbone_name="resnet50", weights=None, trainable_layers=5).eval()
        except TypeError:
            body = resnet_fpn_backbone("resnet50", pretrained=False, trainable_layers=5).eval()
        rfpn = RL.load_resnet_fpn()
        fpn = rfpn.ResnetFPN.__new__(rfpn.ResnetFPN)          # the reference's forward (upsample + concat + camera scale) around
        torch.nn.Module.__init__(fpn)                         # a random-init backbone: its __init__ downloads weights
        from torchvision import transforms
        fpn.resnet_fpn, fpn.freeze, fpn.layer = body, False, "0"
        fpn.transform = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        rpe = RL.load_add_ray_pe()(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
        rpe.load_state_dict(I.make_raype_weights(0), strict=True)
        dec = RL.build_decoder(I.make_weights(0, CFG["queries"]), CFG["queries"], CFG["iterations"])
        cam, Tcp, Twp, Twl = I.make_geometry(1, T, 4 * H, 4 * W, seed=100)
        g = torch.Generator().manual_seed(100)
        batch0 = {"rgb_img": torch.rand(1, T, 3, 4 * H, 4 * W, generator=g), "camera": ns.Camera(cam._data)}
        poses = [ns.Pose(Tcp._data), ns.Pose(Twp._data), ns.Pose(Twl._data)]

        def run():
            with torch.no_grad():
                batch = fpn(dict(batch0))
                enc = rpe(batch["all_features"], batch["camera_feature"], *poses)
                feat = batch["all_features"] + enc
                tok = einops.rearrange(einops.rearrange(feat, "b t c h w -> b t h w c"), "b t h w c -> b (t h w) c")
                return dec(tok.contiguous(), batch["camera_feature"], *poses)
        run()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts)
    except Exception as e:
        print("config-1 pipeline baseline unavailable: %s" % e, file=sys.stderr)
        return None


def gpu_torch_baseline(dev, B):
    """"the bar is stock PyTorch (ATen/cuBLAS/cuDNN) on the same B200".  The unmodified reference module on
    the GPU, same batch as our arm (B clips, tokens resident), 1 warm-up + 3 timed forwards (CUDA events), median:
    fp32 with TF32 off (the parity-grade setting), fp32 with TF32 allowed, and bf16 autocast."""
    res = {}
    try:
        run, kind = reference_runner(CFG["iterations"], device=dev, B=B)
        if kind != "reference":
            return {"unavailable": "reference module not importable on this box"}
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        for name, tf32, amp in (("fp32_tf32_off", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            ts = []
            for i in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    run()
                e1.record()
                torch.cuda.synchronize()
                if i:
                    ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            res[name] = {"clips_per_s": B / (ms * 1e-3), "ms_per_step": ms}
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        res["what"] = ("unmodified reference PARQDecoder.forward (model/parq_decoder.py:134) through stock PyTorch on this GPU, %d clips "
                       "resident on the device, 1 warm-up + 3 timed, median" % B)
    except Exception as e:
        res["unavailable"] = "%s: %s" % (type(e).__name__, e)
    finally:
        torch.cuda.empty_cache()
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    # size the per-step sample so that (steps + warmup) steps end within ~4 minutes: the reference is built with
    # DEC_LAYERS = iters (every recurrent iteration is identical work) and the time is scaled to the 8 of the workload
    run1, kind = reference_runner(1)
    run1()
    t0 = time.perf_counter()
    run1()
    t_it = time.perf_counter() - t0
    total = args.steps + args.warmup
    iters = max(1, min(CFG["iterations"], int(240.0 / max(total * t_it, 1e-6))))
    run, kind = reference_runner(iters)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = (time.perf_counter() - t0) / args.steps
    sec_per_clip = dt * CFG["iterations"] / iters
    value = 1.0 / sec_per_clip
    what = "the unmodified reference PARQDecoder.forward (model/parq_decoder.py:134-163)" if kind == "reference" else \
        "the oracle port in the reference's op order (K/V projection repeated every iteration)"
    sample = ("1 clip per step (8 views x 4800 tokens x 1024 ch, 256 queries), %d of 8 recurrent iterations executed "
              "(DEC_LAYERS=%d) and scaled to 8; %s, fp32, torch CPU, %d threads" % (iters, iters, what, threads))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.gpus == 1:
        c1 = time_c1_pipeline(threads)
        if c1 is not None:
            line["config1_pipeline"] = {"seconds_per_clip": c1, "clips_per_s": 1.0 / c1,
                                        "what": "BASELINE.json configs[0]: random-init ResNet50-FPN + reference AddRayPE + PARQDecoder, "
                                                "1 clip x 8 views 240x320, fp32 CPU, 1 warm-up + 3 timed, median"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- our arm --
def run_ours(args):
    import torch.distributed as dist
    from parq_b200 import _lib, inputs as I
    from parq_b200.decoder import PARQDecoderB200, default_cfg
    from parq_b200.wrappers import Camera, Pose

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (args.gpus, world))
    if args.warmup < 3:
        raise SystemExit("--warmup must be >= 3")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, T, H, W, Nq, IT = CFG["clips_per_gpu"], CFG["views"], CFG["H"], CFG["W"], CFG["queries"], CFG["iterations"]
    Nk, Cc = T * H * W, CFG["C"]
    model = PARQDecoderB200(default_cfg(Nq, IT)).eval()
    model.load_state_dict(I.make_weights(0, Nq), strict=True)
    model = model.to(dev)
    model.feature_hw = (H, W)
    model.use_cuda_graph = True      # one graph launch per step instead of ~180 kernel launches from Python
    # this rank's clips: global clip ids rank*B .. rank*B+B-1 (block partition, parq_b200.shard.clip_range)
    tok_host = torch.empty(B, Nk, Cc, dtype=torch.bfloat16).pin_memory()
    for b in range(B):
        tok_host[b] = I.make_tokens(1, T, H, W, seed=1000 + rank * B + b)[0].to(torch.bfloat16)
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=2000 + rank)
    geo_host = [t._data.pin_memory() for t in (cam, Tcp, Twp, Twl)]
    tokens = tok_host.to(dev)
    geo = [Camera(geo_host[0].to(dev)), Pose(geo_host[1].to(dev)), Pose(geo_host[2].to(dev)), Pose(geo_host[3].to(dev))]
    lib = _lib.load()

    # ---- device-resident throughput -------------------------------------------------------------
    for _ in range(args.warmup):
        model(tokens, *geo)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = model(tokens, *geo)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)

    # ---- live kernel timing: the same steps launched eagerly (the graph replays them unchanged) with CUDA
    # events around every launch of the three roofline kernels, on the launching stream, sampler still running
    model.use_cuda_graph = False
    n0 = lib.parq_kernel_launches()
    _lib.profile_enable(["cross_attn", "project_sample", "kv_proj"], 64 * args.steps + 64)
    for _ in range(args.steps):
        model(tokens, *geo)
    torch.cuda.synchronize()
    launches_per_step = (lib.parq_kernel_launches() - n0) // args.steps
    launches = launches_per_step * args.steps          # kernels inside each timed graph replay x steps
    prof = _lib.profile_collect()
    clocks = sampler.stop() if rank == 0 else None
    _lib.profile_enable([], 0)

    # ---- full per-kernel-class breakdown (separate pass: events around every launch) ------------
    _lib.profile_enable(list(_lib.PROFILE_TAGS), 512)
    model(tokens, *geo)
    torch.cuda.synchronize()
    breakdown = _lib.profile_collect()
    _lib.profile_enable([], 0)
    model.use_cuda_graph = True

    # ---- algorithmic bytes of the gather: count the in-bounds bilinear corners of this very run ----
    dbg = model._engine.forward(tokens, geo[0]._data, geo[1]._data, geo[2]._data, geo[3]._data, H, W, debug=True)
    torch.cuda.synchronize()
    cu, cv = dbg["center_im"][..., 0], dbg["center_im"][..., 1]
    x0, y0 = torch.floor(cu), torch.floor(cv)
    nx = ((x0 >= 0) & (x0 <= W - 1)).int() + ((x0 + 1 >= 0) & (x0 + 1 <= W - 1)).int()
    ny = ((y0 >= 0) & (y0 <= H - 1)).int() + ((y0 + 1 >= 0) & (y0 + 1 <= H - 1)).int()
    n_inb = float((nx * ny).sum().item()) / IT          # in-bounds corner texels per launch (mean over the 8 iterations)
    valid_frac = float(dbg["center_valid"].float().mean().item())
    # ---- the sampling kernel alone, back to back: one launch per iteration's reference points (8 different gathers of
    # ~54 MB out of the 1.26 GB token map, cycled, so consecutive launches do not find their texels in L2), CUDA events
    # around the whole loop -> time per launch without the event/launch gaps that bracket a ~20 us kernel in the step
    import ctypes as C
    from parq_b200.decoder import make_shape, pose_chain, _ptr, _stream
    lo_ = torch.tensor([-3.0, -2.0, 0.25], device=dev)
    span_ = torch.tensor([6.0, 2.5, 5.0], device=dev)
    refs_it = ((dbg["coord_pos"] - lo_) / span_).contiguous()               # (IT, B, Nq, 3) normalised reference points
    Tcl = pose_chain(geo[1]._data, geo[2]._data, geo[3]._data)
    shp = make_shape(B, T, H, W, Cc, Nq, CFG["heads"], CFG["ffn"], 1, 10, (-3, 3, -2, 0.5, 0.25, 5.25))
    feat_out = torch.empty(B, Nq, Cc, dtype=torch.float32, device=dev)
    cam_d = geo[0]._data.contiguous()

    def sample_loop(n):
        for i in range(n):
            _lib.check(lib.parq_project_sample(C.byref(shp), _ptr(tokens), None, _ptr(refs_it[i % IT]), _ptr(Tcl), _ptr(cam_d), _ptr(feat_out),
                                               None, None, None, _stream()), "parq_project_sample")
    sample_loop(IT)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    sample_loop(IT * 8)
    s1.record()
    torch.cuda.synchronize()
    samp_b2b_ms = s0.elapsed_time(s1) / (IT * 8)
    del dbg, refs_it, feat_out

    # ---- end to end through the public module API with host buffers -----------------------------
    copy_stream = torch.cuda.Stream()
    tok_dev = [torch.empty_like(tokens), torch.empty_like(tokens)]
    geo_dev = [[torch.empty_like(g._data) for g in geo] for _ in range(2)]
    keys = ("pred_logits", "center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob")
    res_host = {k: torch.empty_like(out[-1][k], device="cpu").pin_memory() for k in keys}
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()

    def e2e_steps(n):
        for i in range(n):
            s = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])                       # buffer s free again
                tok_dev[s].copy_(tok_host, non_blocking=True)             # H2D of this step's inputs
                for d, h in zip(geo_dev[s], geo_host):
                    d.copy_(h, non_blocking=True)
                copied[s].record(copy_stream)
            main.wait_event(copied[s])
            o = model(tok_dev[s], Camera(geo_dev[s][0]), Pose(geo_dev[s][1]), Pose(geo_dev[s][2]), Pose(geo_dev[s][3]))
            consumed[s].record(main)
            for k in keys:
                res_host[k].copy_(o[-1][k], non_blocking=True)            # D2H of the step's detections
        torch.cuda.synchronize()

    for s in range(2):
        consumed[s].record(main)
    e2e_steps(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d = tok_host.numel() * 2 + sum(g.numel() * 4 for g in geo_host)
    d2h = sum(v.numel() * 4 for v in res_host.values())

    # ---- second end-to-end arm, at the PIPELINE boundary: what an evaluation host really ships over PCIe is the FPN
    # pyramid (bf16 levels, 3.3 MB per view) -- tokens (9.8 MB per view) are born on the device.  Per step: H2D of the four
    # levels + cameras / poses from pinned memory, fpn_concat (f-3) -> AddRayPEB200.tokens (f-1) -> decoder -> parse_pred + NMS
    # (f-2), D2H of the last-iteration detections and pred_mask.
    from parq_b200 import inputs as I2
    from parq_b200.fpn import camera_feature, fpn_concat
    from parq_b200.raype import AddRayPEB200
    rpe = AddRayPEB200(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
    rpe.load_state_dict(I2.make_raype_weights(0), strict=True)
    rpe = rpe.to(dev)
    pyr = I2.make_pyramid(B * T, H, W, seed=3000 + rank)
    pyr_host = {k: v.to(torch.bfloat16).pin_memory() for k, v in pyr.items()}
    del pyr
    cam_img = I2.make_geometry(B, T, 4 * H, 4 * W, seed=2000 + rank)[0]
    camf_host = camera_feature(cam_img)._data.contiguous().pin_memory()          # model/resnet_fpn.py:88-90 (host side, 6 floats per view)
    pyr_dev = [{k: torch.empty_like(v, device=dev) for k, v in pyr_host.items()} for _ in range(2)]
    pgeo_dev = [[torch.empty_like(g._data) for g in geo] for _ in range(2)]
    tok_buf = torch.empty(B, Nk, Cc, dtype=torch.bfloat16, device=dev)             # fixed address: the decoder replays one captured graph
    pkeys = keys + ("pred_mask",)
    pres_host = {k: torch.empty_like(out[-1][k], device="cpu").pin_memory() for k in keys}
    pres_host["pred_mask"] = torch.empty(B, Nq, dtype=torch.bool).pin_memory()

    def pipe_steps(n):
        for i in range(n):
            s = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[s])
                for k in pyr_host:
                    pyr_dev[s][k].copy_(pyr_host[k], non_blocking=True)
                pgeo_dev[s][0].copy_(camf_host, non_blocking=True)
                for d, h in zip(pgeo_dev[s][1:], geo_host[1:]):
                    d.copy_(h, non_blocking=True)
                copied[s].record(copy_stream)
            main.wait_event(copied[s])
            cam_s, Tcp_s, Twp_s, Twl_s = Camera(pgeo_dev[s][0]), Pose(pgeo_dev[s][1]), Pose(pgeo_dev[s][2]), Pose(pgeo_dev[s][3])
            feats = fpn_concat(pyr_dev[s], out_dtype=torch.bfloat16).view(B, T, Cc, H, W)
            toks = rpe.tokens(feats, cam_s, Tcp_s, Twp_s, Twl_s, out=tok_buf)
            o = model(toks, cam_s, Tcp_s, Twp_s, Twl_s)
            parsed_s = model.parse_pred(o)
            consumed[s].record(main)
            for k in pkeys:
                pres_host[k].copy_(parsed_s[k], non_blocking=True)
            del feats
        torch.cuda.synchronize()

    for s in range(2):
        consumed[s].record(main)
    pipe_steps(args.warmup)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    pipe_steps(args.steps)
    p1.record()
    barrier()
    ms_pipe = p0.elapsed_time(p1)
    h2d_pipe = sum(v.numel() * 2 for v in pyr_host.values()) + camf_host.numel() * 4 + sum(g.numel() * 4 for g in geo_host[1:])
    d2h_pipe = sum(v.numel() * v.element_size() for v in pres_host.values())
    del pyr_dev, tok_buf

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_pipe], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_pipe = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())

    # ---- the one collective of the path: device parse_pred + NMS on this rank's clips, then the
    # all_gather of the fixed-size detections into global clip order (what rank 0 feeds the order-dependent F1 fusion,
    # utils/f1_eval.py:293-352).  Verified against per-clip checksums that travel separately.
    from parq_b200 import shard
    model.use_cuda_graph = True
    out = model(tokens, *geo)
    shard.gather_detections({k: model.parse_pred(out)[k] for k in shard.DETECTION_KEYS + ("pred_mask",)}, world * B)   # untimed: NCCL channel set-up
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    parsed = model.parse_pred(out)
    local_det = {k: parsed[k] for k in shard.DETECTION_KEYS + ("pred_mask",)}
    gathered = shard.gather_detections(local_det, world * B)
    g1.record()
    barrier()
    gather_ms = g0.elapsed_time(g1)
    chk = torch.stack([local_det[k].double().flatten(1).sum(1) for k in shard.DETECTION_KEYS + ("pred_mask",)], 1)     # (B, 5) per-clip sums
    if world > 1:
        chks = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(chks, chk)
        chk_all = torch.cat(chks, 0)
        t = torch.tensor([gather_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = t.item()
    else:
        chk_all = chk
    chk_g = torch.stack([gathered[k].double().flatten(1).sum(1) for k in shard.DETECTION_KEYS + ("pred_mask",)], 1)
    gather_ok = bool(torch.equal(chk_g, chk_all)) and all(gathered[k].shape[0] == world * B for k in gathered) and \
        all(torch.equal(gathered[k][rank * B:(rank + 1) * B], local_det[k]) for k in gathered)
    gather_bytes = sum(v.numel() * v.element_size() for v in gathered.values())
    detection_gather = {"ms": gather_ms, "bytes": gather_bytes, "nranks": world, "verified_global_clip_order": gather_ok,
                        "kept_boxes_rank0": int(local_det["pred_mask"].sum().item()),
                        "what": "parse_pred + NMS on the device for this rank's %d clips, then shard.gather_detections (one NCCL all_gather "
                                "per tensor: centre, size, ortho6d, class probabilities, pred_mask) into global clip order" % B}

    # ---- launch trace (N = 1): what every launch costs on the dependent chain of the replayed graph, and the SM clock inside
    # the chained kernel (cycles / ns of its own CTAs) -- the evidence for "the step is power-capped" in DESIGN.md 4.0
    trace = None
    if world == 1 and rank == 0:
        try:
            from parq_b200.tracing import launch_trace
            eng = model._engine
            g4 = [g._data for g in geo]
            tr = launch_trace(eng, lambda: eng.forward(tokens, *g4, H, W, graph=True), reps=8, iters=IT)
            trace = {k: tr[k] for k in ("stamps_per_step", "step_us", "iteration_us", "iteration_launches_us", "prologue_us", "sm_mhz_in_chain_kernel") if k in tr}
            trace["how"] = ("parq_trace: global-timer stamp of every kernel when its stream dependency resolves, 8 graph replays after the timed region; "
                            "a launch costs the difference to the next stamp (execution + drain + hand-over); iterations 1..6; "
                            "sm_mhz_in_chain_kernel = clock64 / globaltimer of the chained kernel's own CTAs")
        except Exception as e:                       # instrumentation only: never fail the bench line
            trace = {"error": repr(e)}

    # ---- BASELINE.json configs[4] (streaming shape: one clip, sliding 8-view window) through the window cache, N = 1: latency per
    # window = copy + K / V^T projection of ONE new view + the 8 iterations over the cached window as one graph replay (secondary
    # record next to the headline: the one-clip launch path -- side streams, cluster split-K GEMM, few-row kernels -- has no other
    # number in this line)
    streaming = None
    if world == 1 and rank == 0:
        try:
            import statistics
            from parq_b200.decoder import DecoderEngine
            from parq_b200.streaming import StreamingWindow
            eng1 = DecoderEngine(I.make_weights(0, Nq), dev)
            nv = T + 8
            stream = I.make_tokens(1, nv, H, W, seed=5)[0].view(nv, H * W, Cc).to(dev).bfloat16()
            cam1, Tcp1, Twp1, _ = I.make_geometry(1, nv, H, W, seed=5)
            cam1, Tcp1, Twp1 = cam1._data.to(dev), Tcp1._data.to(dev), Twp1._data.to(dev)
            sw = StreamingWindow(eng1, T, H, W)
            for v in range(T - 1):
                sw.push(stream[v:v + 1], cam1[:, v], Tcp1[:, v], Twp1[:, v])
            lat = []
            for i in range(110):
                v = (T - 1 + i) % nv
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sw.push(stream[v:v + 1], cam1[:, v], Tcp1[:, v], Twp1[:, v])
                sw.decode(Twp1[:, (v - T // 2) % nv].reshape(1, 1, 12))
                e1.record()
                torch.cuda.synchronize()
                if i >= 10:
                    lat.append(e0.elapsed_time(e1))
            lat.sort()
            streaming = {"p50_ms": statistics.median(lat), "p99_ms": lat[int(0.99 * len(lat)) - 1], "windows": len(lat),
                         "what": "BASELINE.json configs[4]: 1 clip, sliding window of %d views %dx%d, %d queries, %d iterations; per window: copy + "
                                 "K / V^T projection of one view into the window cache (f-4), then one CUDA-graph replay of the decoder" % (T, H, W, Nq, IT)}
            del sw, eng1
        except Exception as e:                       # secondary record: never fail the bench line
            streaming = {"error": repr(e)}

    gpu_torch = gpu_torch_baseline(dev, B) if (world == 1 and rank == 0) else None

    if rank == 0:
        pk = peaks()
        step_ms = ms / args.steps
        value = world * B / (step_ms * 1e-3)
        ca_ms, ca_n = prof["cross_attn"]
        flops = 4.0 * B * Nq * Nk * Cc                                    # QK^T + PV over all heads, per launch
        ach = flops / (ca_ms / max(ca_n, 1) * 1e-3) / 1e12 if ca_n else None
        ps_ms, ps_n = prof["project_sample"]
        # texels actually fetched (bf16) + the sampled features (4 B/element) + center_im/valid + ref + poses/cameras.
        # The kernel moves exactly that, minus the optional center_im / center_valid outputs (debug only): the features leave
        # it once, as the bf16 [hi|lo] operand split (2 + 2 bytes per element).
        samp_bytes = n_inb * Cc * 2 + B * Nq * Cc * 4 + B * T * Nq * 9.0 + B * Nq * 12.0 + B * T * 72.0
        samp_moved = n_inb * Cc * 2 + B * Nq * Cc * 4 + B * Nq * 12.0 + B * T * 72.0
        kv_ms, kv_n = prof["kv_proj"]
        traffic, traffic_src, samp_traffic = None, None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("cross_attn_dram_bytes_per_launch")
            samp_traffic = tj.get("project_sample_dram_bytes_per_launch")
            traffic_src = "ncu --set full capture summarised in profiles/r2_ncu_full_iter.md, taken at commit %s" % tj.get("captured_at_commit")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": world * B / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "PARQDecoderB200.forward on pinned host tokens (bf16) + poses, double-buffered H2D on a copy stream, D2H of the last-iteration detections"},
            "e2e_pipeline": {"value": world * B / (ms_pipe / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_pipe, "d2h_bytes_per_step": d2h_pipe,
                             "ms_per_step": ms_pipe / args.steps,
                             "note": "pinned host FPN pyramid (4 bf16 levels) + cameras / poses -> fpn_concat (f-3, bf16 all_features) -> AddRayPEB200.tokens (f-1) -> "
                                     "PARQDecoderB200.forward -> parse_pred + NMS (f-2) -> D2H of the detections and pred_mask; double-buffered H2D"},
            "gpu_launches": launches,
            "roofline": {"kernel": "attn3_tc_kernel<bf16> (CTA-pair flash cross-attention, stream-K schedule, over %d image tokens)" % Nk, "bound": "tensor",
                         "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": (ach / pk["tf_sustained"]) if ach else None,
                         "frac_of_burst_peak": (ach / pk["tf_burst"]) if ach else None, "peak_source": pk["source"] + " (sustained cuBLAS bf16)",
                         "flops_per_launch": flops, "ms_per_launch": ca_ms / max(ca_n, 1), "launches_timed": ca_n, "traffic": traffic,
                         "traffic_source": traffic_src},
            # headline timing of the ~18 us sampling kernel = the back-to-back loop (CUDA events around 64 launches, live in this
            # run): bracketing ONE such launch with events adds ~8 us of event / launch gap to it (kept as `in_step`); the ncu
            # launch list of this same command (profiles/r2_launches_bench.md, gpu__time_duration) is the cross-check
            "roofline_sampling": {"kernel": "project_sample_kernel", "bound": "hbm",
                                  "achieved": samp_bytes / (samp_b2b_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                  "frac": samp_bytes / (samp_b2b_ms * 1e-3) / 1e9 / pk["hbm"],
                                  "algorithmic_bytes": samp_bytes, "moved_bytes": samp_moved, "traffic": samp_traffic, "bytes_per_launch": samp_bytes,
                                  "in_bounds_corners_per_launch": n_inb, "valid_view_fraction": valid_frac,
                                  "ms_per_launch": samp_b2b_ms, "launches_timed": IT * 8,
                                  "timing": "64 launches of the kernel back to back, cycling the 8 iterations' reference points (8 different ~54 MB gathers "
                                            "out of the 1.26 GB token map), CUDA events around the loop",
                                  "in_step": {"ms_per_launch": ps_ms / max(ps_n, 1), "launches_timed": ps_n,
                                              "frac": samp_bytes / (ps_ms / max(ps_n, 1) * 1e-3) / 1e9 / pk["hbm"] if ps_n else None,
                                              "how": "CUDA events around each single launch inside the timed steps (includes the event / launch gaps)"}},
            "roofline_kv_proj": {"kernel": "gemm2_tc_kernel (CTA-pair GEMM: K and V^T projection, 2 launches/step)", "bound": "tensor",
                                 "achieved": (4.0 * B * Nk * Cc * Cc) / (kv_ms / max(kv_n // 2, 1) * 1e-3) / 1e12 if kv_n else None,
                                 "peak": pk["tf_sustained"], "unit": "TFLOP/s"},
            "detection_gather": detection_gather,
            "breakdown_ms_per_step": {k: round(v[0], 4) for k, v in breakdown.items() if k != "_dropped"},
            "breakdown_launches": {k: v[1] for k, v in breakdown.items() if k != "_dropped"},
        }
        if world == 1:
            threads = os.cpu_count() or 1
            sec, kind = time_cpu_reference(threads)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": "1 clip (8 views x 4800 tokens x 1024 ch, 256 queries, all 8 iterations), 1 warm-up + 3 timed, median; "
                                              + ("the unmodified reference PARQDecoder.forward" if kind == "reference" else "oracle port, reference op order")
                                              + ", fp32 torch CPU"}
            line["gpu_torch_baseline"] = gpu_torch
            line["launch_trace"] = trace
            line["streaming_window"] = streaming
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
