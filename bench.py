#!/usr/bin/env python
"""bench.py - pages/s of the dense-detection hot path (train decode + assign + inference decode + NMS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl kzd|reference]

One "step" = one pass of the hot path over one batch of synthetic kuzushiji pages per GPU:
    bbox_decode (utils/loss.py:197)  ->  x stride (caller glue, torch)  ->  TaskAlignedAssigner.forward
    Detect._inference decode (nn/modules/head.py:100)  ->  ops.non_max_suppression
Workload (BASELINE.json configs[1]): YOLOv12x P2-P5 head, 1280x1280 pages, batch 16 per GPU,
~600 GT boxes/page, single class; NMS at the predictor defaults conf 0.25 / iou 0.7 / max_det 300 on
inference-like class logits (~3000 candidates/page). Images are sharded over ranks
(weak scaling: 16 pages per GPU); the only collective is an all-reduce of 4 counters.

Prints ONE JSON line (rank 0). `value` = pages/s with inputs resident in HBM (CUDA events, max over
ranks); `e2e` = the same step fed from pinned HOST buffers with the H2D copy and the D2H of the results
inside the timed region; `roofline` = dominant kernel vs MEASURED_PEAKS.json; `cpu_baseline` = the
oracle's torch CPU port of the reference timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(imgsz=1280, batch=16, G=600, nc=1, strides=(4, 8, 16, 32), topk=10, alpha=0.5, beta=6.0,
                conf=0.25, iou=0.7, max_det=300, nms_candidates=3000)
METRIC = "pages/sec assign+decode+NMS @1280^2 P2-P5"
FALLBACK_HBM_GBS = 6650.0


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(b, A, G, nc, n_cand, n_keep):
    """SURVEY.md 8(d) per-page figures x pages per launch (all stated again in DESIGN.md)."""
    return {
        "decode_train": 272 * A * b,
        "assign": int((57 + 8 / b) * A + 24 * G) * b if nc == 1 else None,
        "decode_infer": (272 + 8 * nc) * A * b,
        "nms": int((32 + 4 * nc) * A * b + 24 * (n_cand + n_keep)),
    }


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's torch port of the reference on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_page_inputs(seed):
    from kuzushiji_vision_b200 import synth

    w = WORKLOAD
    d = synth.assigner_inputs(1, w["G"], w["imgsz"], w["nc"], w["strides"], seed=seed, device="cpu")
    lv = synth.head_outputs(1, w["nc"], w["imgsz"], w["strides"], seed=seed + 100, device="cpu",
                            sparse_cls=w["nms_candidates"])
    return d, lv


def cpu_step(d, lv):
    """One page through the CPU port: same four stages, same order, as the CUDA step."""
    from oracle import torch_oracle as O

    w = WORKLOAD
    boxes = O.decode_train(d["pred_distri"], d["anchor_points"])
    out = O.tal_assign(d["pd_scores"], boxes * d["stride_tensor"], d["anc_points"], d["gt_labels"], d["gt_bboxes"],
                       d["mask_gt"], w["topk"], w["nc"], w["alpha"], w["beta"])
    y = O.decode_infer(lv, w["strides"], w["nc"])
    det = O.non_max_suppression(y, w["conf"], w["iou"], max_det=w["max_det"])
    return out, det


def cpu_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return max(1, n)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = cpu_threads()
    torch.set_num_threads(threads)
    d, lv = cpu_page_inputs(0)
    for _ in range(args.warmup):
        cpu_step(d, lv)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(d, lv)
    dt = time.perf_counter() - t0
    val = args.steps / dt
    w = WORKLOAD
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 1280x1280 P2-P5 (A=136000), ~600 GT/page, nc=1; "
                               "one page per step on the host CPU", **{k: w[k] for k in ("imgsz", "G", "nc")}},
        "cpu_baseline": {"value": val, "unit": "pages/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x 1 page of the workload (oracle torch port of the reference, "
                                   f"{threads} threads)"},
        "e2e": {"value": val, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------

def run_kzd(args, rank, local_rank, world):
    import kuzushiji_vision_b200 as kzd
    from kuzushiji_vision_b200 import dist as kdist
    from kuzushiji_vision_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the kzd arm has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = WORKLOAD
    b, nc, G = w["batch"], w["nc"], w["G"]
    glo, ghi = kdist.shard_range(b * world, rank, world)  # this rank's pages of the global batch

    # synthetic inputs: training-side head outputs (N(0,1)) and inference-side ones (sparse objects)
    d = synth.assigner_inputs(b, G, w["imgsz"], nc, w["strides"], seed=glo, device="cpu")
    lv_cpu = synth.head_outputs(b, nc, w["imgsz"], w["strides"], seed=glo + 100, device="cpu",
                                sparse_cls=w["nms_candidates"])
    A = d["anc_points"].shape[0]
    host = {k: d[k].pin_memory() for k in ("pred_distri", "pd_scores", "gt_labels", "gt_bboxes", "mask_gt")}
    host_lv = [l.pin_memory() for l in lv_cpu]
    dv = {k: v.to(dev) for k, v in host.items()}
    dv_lv = [l.to(dev) for l in host_lv]
    anchor_points, stride_tensor = d["anchor_points"].to(dev), d["stride_tensor"].to(dev)
    anc_points = d["anc_points"].to(dev)
    assigner = kzd.TaskAlignedAssigner(w["topk"], nc, w["alpha"], w["beta"])
    counters = torch.zeros(4, dtype=torch.float64, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(t, lv, marks=None):
        def mark():
            if marks is not None:
                e = ev()
                e.record()
                marks.append(e)
        mark()
        boxes = kzd.bbox_decode(anchor_points, t["pred_distri"])
        mark()
        pd_bboxes = boxes * stride_tensor  # caller glue (utils/loss.py:236), plain torch
        mark()
        tgt = assigner(t["pd_scores"], pd_bboxes, anc_points, t["gt_labels"], t["gt_bboxes"], t["mask_gt"])
        mark()
        y = kzd.decode_inference(lv, w["strides"], nc)
        mark()
        rows, n_keep = kzd.nms_batched(y, w["conf"], w["iou"], max_det=w["max_det"])
        mark()
        return tgt, rows, n_keep

    # ---- device-resident timing ----
    for _ in range(max(args.warmup, 3)):
        tgt, rows, n_keep = step(dv, dv_lv)
    torch.cuda.synchronize()
    kdist.barrier()
    marks = []
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        t_start, t_end = ev(), ev()
        t_start.record()
        for _ in range(args.steps):
            tgt, rows, n_keep = step(dv, dv_lv, marks)
        counters[0:2] = assigner.counters
        counters[2] = float(b)
        counters[3] = float(sum(n_keep))
        kdist.all_reduce_counters(counters)  # the path's only collective: 32 bytes
        t_end.record()
        torch.cuda.synchronize()
    kdist.barrier()
    ms_total = kdist.max_over_ranks(t_start.elapsed_time(t_end), dev)
    pages = float(counters[2].item())
    value = pages * args.steps / (ms_total * 1e-3)

    stage_names = ["decode_train", "glue_mul_stride", "assign", "decode_infer", "nms"]
    stage_ms = {n: 0.0 for n in stage_names}
    for s in range(args.steps):
        m = marks[s * 6:(s + 1) * 6]
        for i, n in enumerate(stage_names):
            stage_ms[n] += m[i].elapsed_time(m[i + 1])
    stage_ms = {n: v / args.steps for n, v in stage_ms.items()}

    # ---- end to end from pinned host buffers ----
    def e2e_step():
        t = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        lv = [l.to(dev, non_blocking=True) for l in host_lv]
        tgt, rows, n_keep = step(t, lv)  # nms_batched reads its counters back (D2H + sync)
        res = rows.cpu()
        cnt = assigner.counters.cpu()
        return res, cnt

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    kdist.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(e2e_steps):
        res, cnt = e2e_step()
    t1.record()
    torch.cuda.synchronize()
    e2e_ms = kdist.max_over_ranks(t0.elapsed_time(t1), dev)
    e2e_value = pages * e2e_steps / (e2e_ms * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + sum(l.numel() * 4 for l in host_lv)
    d2h = rows.numel() * 4 + 2 * 8 + 2 * b * 4

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ----
    n_cand = int(w["nms_candidates"] * b * 0.98)
    ab = algorithmic_bytes(b, A, G, nc, n_cand, int(sum(n_keep)))
    single_kernel = {"decode_train": "decode_train_kernel", "decode_infer": "decode_infer_kernel<4>"}
    dom = max(single_kernel, key=lambda n: stage_ms[n])
    peak, peak_src = hbm_peak()
    achieved = ab[dom] / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": single_kernel[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab[dom],
                "stages": {n: {"ms": stage_ms[n], "algorithmic_bytes": ab.get(n),
                               "gbs": (ab[n] / (stage_ms[n] * 1e-3) / 1e9) if ab.get(n) else None}
                           for n in stage_names}}

    # ---- CPU baseline on a bounded sample (one page, repeated) ----
    threads = cpu_threads()
    torch.set_num_threads(threads)
    cd, clv = cpu_page_inputs(0)
    t0c = time.perf_counter()
    n_cpu = 0
    while n_cpu < 2 or (time.perf_counter() - t0c < 12.0 and n_cpu < 8):
        cpu_step(cd, clv)
        n_cpu += 1
    cpu_dt = time.perf_counter() - t0c
    cpu_baseline = {"value": n_cpu / cpu_dt, "unit": "pages/s", "cores": threads, "kind": "port",
                    "sample": f"{n_cpu} pages of the workload, one at a time (oracle torch port of the reference's "
                              f"CPU path, {threads} threads, no warm-up discarded)"}

    line = {
        "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: YOLOv12x P2-P5 head outputs, 1280x1280 pages, batch 16 per GPU, "
                               "~600 GT/page, nc=1; NMS conf .25 iou .7 max_det 300 on ~3000 candidates/page",
                   "pages_per_gpu": b, "anchors_per_page": A, "gt_slots": G, "nc": nc,
                   "l2_policy": "inputs larger than L2 (1.16 GB read per step vs 126 MB L2)",
                   "parallelism": f"image-sharded x{world}"},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": "pages/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": 10 * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "counters": {"target_scores_sum": float(counters[0]), "n_fg": float(counters[1]),
                     "kept_boxes": float(counters[3])},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kzd", choices=["kzd", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    from kuzushiji_vision_b200 import dist as kdist

    kdist.init()
    try:
        run_kzd(args, rank, local_rank, world)
    finally:
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
