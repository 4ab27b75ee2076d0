"""BASELINE.json configuration 3: the full eval pipeline on N GPUs, clips sharded by clip.

    python tools/eval_pipeline.py [--clips 64] [--batch 16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/eval_pipeline.py --clips 1024

Per batch of clips: stock torchvision ResNet50-FPN backbone (random init, fp32 torch/cuDNN -- out of scope of this
library, as north_star says) -> fpn_concat (f-3) -> AddRayPEB200.tokens (f-1) -> PARQDecoderB200 -> parse_pred + NMS
(f-2); finally one all_gather of the fixed-size detections in global clip order (the only collective).  Prints one
JSON line with whole-job clips/s and the per-stage device times.  Synthetic RGB clips (8 views of 240x320), random
ScanNet-like poses."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from parq_b200 import inputs as I, shard
from parq_b200.decoder import PARQDecoderB200, default_cfg
from parq_b200.fpn import camera_feature, fpn_concat
from parq_b200.raype import AddRayPEB200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64, help="global number of clips")
    ap.add_argument("--batch", type=int, default=16, help="clips per decoder batch per GPU")
    ap.add_argument("--views", type=int, default=8)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
    torch.manual_seed(0)
    backbone = resnet_fpn_backbone(backbone_name="resnet50", weights=None, trainable_layers=5).eval().to(dev)
    T, Himg, Wimg, H, W, Nq = args.views, 240, 320, 60, 80, 256
    rpe = AddRayPEB200(1024, [-3, 3, -2, 0.5, 0.25, 5.25], 64, 0.25, 5.25).eval()
    rpe.load_state_dict(I.make_raype_weights(0), strict=True)
    rpe = rpe.to(dev)
    dec = PARQDecoderB200(default_cfg(Nq)).eval()
    dec.load_state_dict(I.make_weights(0, Nq), strict=True)
    dec = dec.to(dev)
    dec.feature_hw = (H, W)
    dec.use_cuda_graph = True            # one graph launch per decoder call (static geometry buffers below)
    lo, hi = shard.clip_range(args.clips, rank, world)
    mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stages = ("backbone", "fpn_concat", "raype_tokens", "decoder", "parse_pred")
    ev = {s: [] for s in stages}
    dets = []

    def make_inputs(c0, c1):                     # synthetic data generation, outside the timed region
        n = c1 - c0
        g = torch.Generator(device="cpu").manual_seed(1000 + c0)
        rgb = torch.rand(n, T, 3, Himg, Wimg, generator=g).to(dev)
        return (rgb,) + tuple(t.to(dev) for t in I.make_geometry(n, T, Himg, Wimg, seed=c0))

    batches = [(c0, min(hi, c0 + args.batch)) for c0 in range(lo, hi, args.batch)]
    data = {b: make_inputs(*b) for b in batches}

    static = {}

    def run_batch(c0, c1, record):
        n = c1 - c0
        rgb, cam_img, Tcp, Twp, Twl = data[(c0, c1)]
        if n not in static:                      # fixed device buffers per batch size -> the captured graph is reused
            static[n] = tuple(type(t)(torch.empty_like(t._data)) for t in (cam_img, Tcp, Twp, Twl)) + \
                        (torch.empty(n, T * H * W, 1024, dtype=torch.bfloat16, device=dev),)
        for dst, src in zip(static[n][1:4], (Tcp, Twp, Twl)):
            dst._data.copy_(src._data)
        cam_img_in = cam_img
        cam_feat, Tcp, Twp, Twl, tok_buf = static[n]
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
        with torch.no_grad():
            marks[0].record()
            pyr = backbone(((rgb.view(n * T, 3, Himg, Wimg) - mean) / std))                 # model/resnet_fpn.py:64-71
            marks[1].record()
            feats = fpn_concat(pyr, out_dtype=torch.bfloat16).view(n, T, 1024, H, W)                                    # :73-85
            cam_feat._data.copy_(camera_feature(cam_img_in)._data)                             # :88-90 (into the static buffer)
            cam = cam_feat
            marks[2].record()
            tokens = rpe.tokens(feats, cam, Tcp, Twp, Twl, out=tok_buf)                                  # parq_lightning.py:72-85
            marks[3].record()
            outs = dec(tokens, cam, Tcp, Twp, Twl)                                            # :88
            marks[4].record()
            parsed = dec.parse_pred(outs)                                                     # parq_decoder.py:372-424
            marks[5].record()
        if record:
            torch.cuda.synchronize()
            for i, s in enumerate(stages):
                ev[s].append(marks[i].elapsed_time(marks[i + 1]))
            dets.append({k: parsed[k] for k in shard.DETECTION_KEYS} | {"pred_mask": parsed["pred_mask"]})

    run_batch(*batches[0], False)                            # warm-up: cuDNN autotune, weight packing, workspaces
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for b in batches:
        run_batch(*b, True)
    last = {k: torch.cat([d[k] for d in dets], 0) for k in shard.DETECTION_KEYS + ("pred_mask",)}
    gathered = shard.gather_detections(last, args.clips)     # the single collective: detections in global clip order
    t1.record()
    barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_local = hi - lo
        print(json.dumps({"config": "C3 full eval pipeline: backbone (stock torch) + fpn_concat + AddRayPE tokens + decoder + parse_pred/NMS, "
                                    "clips sharded by clip, one all_gather of detections",
                          "n_gpus": world, "clips": args.clips, "clips_per_gpu": n_local, "batch": args.batch, "ms_total": ms.item(),
                          "clips_per_s": args.clips / ms.item() * 1e3,
                          "stage_ms_per_batch_rank0": {s: round(sum(v) / max(len(v), 1), 3) for s, v in ev.items()},
                          "gathered": {k: list(v.shape) for k, v in gathered.items()},
                          "kept_boxes_rank0": int(sum(int(d["pred_mask"].sum()) for d in dets))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
