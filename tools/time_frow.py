"""Device times of the "next rows" kernels: loss tail (dense and sparse-target BCE, forward and
backward), GT packing and the post-NMS scale/clip tail, on BASELINE configs[1] (nc=1, b=16) and configs[2]
(nc=4300). CUDA events, warm, L2 flushed before every call. Writes gpurun_out/r2_frow_times.json.

    python tools/time_frow.py [--b3 8]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd  # noqa: E402
from kuzushiji_vision_b200 import loss as kloss  # noqa: E402
from kuzushiji_vision_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b3", type=int, default=8, help="batch of the nc=4300 case (16 needs ~150 GB with the dense form's buffers)")
args = ap.parse_args()
dev = torch.device("cuda:0")
PEAK = 6524.9
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3  # us


import collections  # noqa: E402

from torch.profiler import ProfilerActivity, profile  # noqa: E402


def loss_case(name, b, G, nc):
    imgsz, strides = 1280, synth.P2_P5
    A = synth.num_anchors(imgsz, strides)
    g = torch.Generator(device=dev).manual_seed(3)
    logits = torch.randn((b, A, nc), generator=g, device=dev)
    pred_distri = torch.randn((b, A, 64), generator=g, device=dev)
    gl, gb, mg = synth.gt_pages(b, G, imgsz, nc, seed=0, device=dev)
    feats = [torch.empty((1, 1, h, w), device=dev) for h, w in synth.level_shapes(imgsz, strides)]
    anchor_points, stride_tensor = kzd.make_anchors(feats, strides, 0.5)
    boxes, px = kzd.bbox_decode(anchor_points, pred_distri, stride_tensor)
    asg = kzd.TaskAlignedAssigner(10, nc, 0.5, 6.0)
    anc_px = anchor_points * stride_tensor
    sig = logits.sigmoid()
    tl, tb, ts, fg, gi = asg(sig, px, anc_px, gl, gb, mg)
    tss = asg.counters[0].float()
    del sig
    n = b * A * nc
    res = {"b": b, "G": G, "nc": nc, "A": A, "n_fg": int(fg.sum())}
    res["assign_us"] = timeit(lambda: asg(logits, px, anc_px, gl, gb, mg), iters=5)  # (logits stand in for scores: timing only)
    gout = torch.ones(3, device=dev)

    def run(sparse, backward):
        x = logits.detach().requires_grad_(backward)
        p = pred_distri.detach().requires_grad_(backward)
        out = kloss.loss_tail(x, p, anchor_points, stride_tensor, tb, ts, fg, tss, target_labels=tl if sparse else None)
        if backward:
            out.backward(gout)
        return out

    for sparse in (False, True):
        key = "sparse" if sparse else "dense"
        f = timeit(lambda: run(sparse, False), iters=5)
        fb = timeit(lambda: run(sparse, True), iters=5)
        # algorithmic bytes of the BCE part: forward reads x (+ t when dense); backward reads x (+ t) and writes g
        fwd_bytes = n * 4 * (1 if sparse else 2)
        bwd_bytes = n * 4 * (2 if sparse else 3)
        res[key] = {"fwd_us": f, "fwd_bwd_us": fb, "bce_fwd_bytes": fwd_bytes, "bce_bwd_bytes": bwd_bytes,
                    "fwd_gbs_on_bce_bytes": fwd_bytes / (f * 1e-6) / 1e9, "fwd_frac_of_peak": fwd_bytes / (f * 1e-6) / 1e9 / PEAK,
                    "bwd_gbs_on_bce_bytes": bwd_bytes / ((fb - f) * 1e-6) / 1e9,
                    "bwd_frac_of_peak": bwd_bytes / ((fb - f) * 1e-6) / 1e9 / PEAK}
    # AMP flow: fp16 logits and distances read natively by the sparse form (no fp32 copies)
    lh, ph = logits.half(), pred_distri.half()

    def run_half(backward):
        x = lh.detach().requires_grad_(backward)
        pp = ph.detach().requires_grad_(backward)
        out = kloss.loss_tail(x, pp, anchor_points, stride_tensor, tb, ts, fg, tss, target_labels=tl)
        if backward:
            out.backward(gout)
        return out

    fh = timeit(lambda: run_half(False), iters=5)
    fbh = timeit(lambda: run_half(True), iters=5)
    res["sparse_fp16"] = {"fwd_us": fh, "fwd_bwd_us": fbh, "bce_fwd_bytes": n * 2, "bce_bwd_bytes": n * 4,
                          "fwd_gbs_on_bce_bytes": n * 2 / (fh * 1e-6) / 1e9, "bwd_gbs_on_bce_bytes": n * 4 / ((fbh - fh) * 1e-6) / 1e9}
    with profile(activities=[ProfilerActivity.CUDA]) as prof_h:
        for _ in range(3):
            flush.zero_()
            run_half(True)
        torch.cuda.synchronize()
    aggh = collections.OrderedDict()
    for e in prof_h.events():
        if e.device_type.name == "CUDA" and "kzd::" in e.name:
            aggh.setdefault(e.name.replace("kzd::", "").split("(")[0][-40:], []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
    res["sparse_fp16_kernels_us"] = {k: round(sum(v) / len(v), 2) for k, v in aggh.items()}
    del lh, ph
    # per-kernel device times of one sparse-target forward + backward (CUPTI): the eager figures above include the
    # host side of the autograd Function and of the torch glue, which dominates at nc = 1
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            flush.zero_()
            run(True, True)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for e in prof.events():
        if e.device_type.name == "CUDA" and "kzd::" in e.name:
            agg.setdefault(e.name.replace("kzd::", "").split("(")[0][-40:], []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
    res["sparse_fwd_bwd_kernels_us"] = {k: round(sum(v) / len(v), 2) for k, v in agg.items()}
    a, bq = run(False, False), run(True, False)
    res["sparse_vs_dense_max_rel"] = float(((a - bq).abs() / a.abs().clamp(min=1e-6)).max())
    return res


out = {"peak_gbs": PEAK}
out["config2_nc1"] = loss_case("cfg2", 16, 600, 1)
torch.cuda.empty_cache()
out["config3_nc4300"] = loss_case("cfg3", args.b3, 600, 4300)
torch.cuda.empty_cache()

# GT packing (v8DetectionLoss.preprocess) at 16 x 600 labels and the post-NMS scale/clip tail at 16 x 1000 rows
g = torch.Generator().manual_seed(1)
rows = []
for j in range(16):
    nlab = 600
    rows.append(torch.cat((torch.full((nlab, 1), float(j)), torch.zeros(nlab, 1), torch.rand(nlab, 2, generator=g) * 0.8 + 0.1,
                           torch.rand(nlab, 2, generator=g) * 0.04 + 0.01), 1))
t = torch.cat(rows).to(dev)
scale = torch.tensor([1280.0, 1280.0, 1280.0, 1280.0], device=dev)
out["gt_pack_16x600_us"] = timeit(lambda: kloss.preprocess(t, 16, scale), iters=10)
det = torch.rand((16, 1000, 6), device=dev) * 1280
out["scale_boxes_16x1000_us"] = timeit(lambda: kzd.scale_boxes_batched(det, [1000] * 16, (1280, 1280), [(2000, 1500)] * 16), iters=10)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        kloss.preprocess(t, 16, scale)
        kzd.scale_boxes_batched(det, [1000] * 16, (1280, 1280), [(2000, 1500)] * 16)
    torch.cuda.synchronize()
kk = {}
for e in prof.events():
    if e.device_type.name == "CUDA" and "at::" not in e.name and "Memcpy" not in e.name and "Memset" not in e.name:
        kk.setdefault(e.name.replace("kzd::", "").split("(")[0][-40:], []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
out["gt_scale_kernels_us"] = {k: round(sum(v) / len(v), 2) for k, v in kk.items()}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_frow_times.json", "w"), indent=1)
print(json.dumps(out, indent=1))
