#!/usr/bin/env python
"""bench.py - pages/s of the dense-detection hot path (train decode + assign + inference decode + NMS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl kzd|reference]

One "step" = one pass of the hot path over one batch of synthetic kuzushiji pages per GPU, from the
per-level head maps a model forward emits to assignment targets and final detections:
    training-side decode  bbox_decode (utils/loss.py:197) x stride  ->  TaskAlignedAssigner.forward
    inference decode      Detect._inference (nn/modules/head.py:100)  ->  ops.non_max_suppression
Both decodes consume the same head outputs (the validation-during-training flow,
engine/validator.py:180-185), so by default ONE kernel reads the maps once and emits both (PagePipeline
fuse_decodes); KZD_BENCH_FUSE=0 times the two separate decode kernels of the drop-in signatures.
Workload (BASELINE.json configs[1]): YOLOv12x P2-P5 head, 1280x1280 pages, batch 16 per GPU,
~600 GT boxes/page, single class; NMS at the reference's kuzushiji test-script settings (conf 0.25, iou 0.7,
max_det 1000: scripts/test_character_detection.py:237-243) on inference-like class logits (~3000
candidates/page). Images are sharded over ranks (weak scaling: 16 pages per GPU); the only
collective is an all-reduce of 4 counters.

Prints ONE JSON line (rank 0). `value` = pages/s with inputs resident in HBM (CUDA events, max over
ranks); `e2e` = the same step fed from pinned HOST buffers with the H2D copy and the D2H of the results
inside the timed region; `roofline` = dominant kernel vs MEASURED_PEAKS.json; `cpu_baseline` = the
oracle's torch CPU port of the reference timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(imgsz=1280, batch=16, G=600, nc=1, strides=(4, 8, 16, 32), topk=10, alpha=0.5, beta=6.0,
                conf=0.25, iou=0.7, max_det=1000, nms_candidates=3000)
METRIC = "pages/sec assign+decode+NMS @1280^2 P2-P5"
FALLBACK_HBM_GBS = 6650.0


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~5 ms while the timed region runs
    (nvidia-smi -lms 200 is too coarse for a region of a few tens of milliseconds)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self._stop, self.t = index, [], set(), None, False, None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.t = None
        return self

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in bits.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.005)

    def __exit__(self, *a):
        self._stop = True
        if self.t:
            self.t.join(timeout=1)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def algorithmic_bytes(b, A, G, nc, n_cand, n_keep):
    """SURVEY.md 8(d) per-page figures x pages per launch (all stated again in DESIGN.md)."""
    return {
        # 272*A for the decode + 16*A for the pixel-space copy (x stride, utils/loss.py:236) that the
        # same kernel writes (the reference's separate elementwise pass would read 16*A more)
        "decode_train": (272 + 16) * A * b,
        "assign": int((57 + 8 / b) * A + 24 * G) * b if nc == 1 else None,
        "decode_infer": (272 + 8 * nc) * A * b,
        # fused: one read of the maps 4(64+nc)A, writes y 4(4+nc)A + grid-unit boxes 16A + pixel boxes 16A
        "decode_dual": (4 * (64 + nc) + 4 * (4 + nc) + 32) * A * b,
        "nms": int((32 + 4 * nc) * A * b + 24 * (n_cand + n_keep)),
    }


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's torch port of the reference on the host cores
# ------------------------------------------------------------------------------------------------

def page_inputs(n_pages, seed, device="cpu"):
    """Synthetic head outputs + GT for `n_pages` pages of the workload. ONE set of per-level head maps
    feeds both halves of the step (as a model forward would): DFL logits ~ N(0,1); class logits like a
    freshly initialised Detect head (bias_init drives them strongly negative: background -6) with
    ~3000 object-like anchors per page at N(1,1). The training-side layout (pred_distri, pd_scores) is
    derived from the maps with the reference's own glue ops (utils/loss.py:210-215,235)."""
    from kuzushiji_vision_b200 import synth

    w = WORKLOAD
    lv = synth.head_outputs(n_pages, w["nc"], w["imgsz"], w["strides"], seed=seed, device=device,
                            sparse_cls=w["nms_candidates"])
    gl, gb, mg = synth.gt_pages(n_pages, w["G"], w["imgsz"], w["nc"], seed=seed, device=device)
    return lv, gl, gb, mg


def train_layout(lv, nc):
    """utils/loss.py:210-215 + :235: cat levels, split, permute, contiguous, sigmoid (caller glue)."""
    b = lv[0].shape[0]
    cat = torch.cat([x.view(b, 64 + nc, -1) for x in lv], 2)
    pred_distri = cat[:, :64].permute(0, 2, 1).contiguous()
    pd_scores = cat[:, 64:].permute(0, 2, 1).contiguous().sigmoid()
    return pred_distri, pd_scores


def cpu_step(lv, gl, gb, mg):
    """One page through the CPU port: same four stages, same order, as the CUDA step."""
    from oracle import torch_oracle as O

    w = WORKLOAD
    pred_distri, pd_scores = train_layout(lv, w["nc"])
    anchor_points, stride_tensor = O.make_anchors([x.shape[2:] for x in lv], w["strides"])
    boxes = O.decode_train(pred_distri, anchor_points)
    out = O.tal_assign(pd_scores, boxes * stride_tensor, anchor_points * stride_tensor, gl, gb, mg, w["topk"],
                       w["nc"], w["alpha"], w["beta"])
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
    """Reference arm: the reference's CPU implementation of the path (oracle torch port - the Python
    reference cannot travel to the GPU box) on all host cores; one page of the workload per step. If K
    full pages would not fit in ~4 minutes the page is cropped to a quarter (640x640, G/4) and each step
    counts as 0.25 page - the reference's cost grows faster than linearly with page size, so the crop
    only flatters the CPU arm."""
    if rank != 0:
        return
    threads = cpu_threads()
    torch.set_num_threads(threads)
    w = WORKLOAD
    page = page_inputs(1, 0)
    t0 = time.perf_counter()
    cpu_step(*page)
    t_page = time.perf_counter() - t0
    frac, note = 1.0, "1 full page per step"
    if t_page * (args.steps + args.warmup) > 240.0:
        keep = dict(w)
        w.update(imgsz=w["imgsz"] // 2, G=w["G"] // 4, nms_candidates=w["nms_candidates"] // 4)
        page = page_inputs(1, 0)
        frac, note = 0.25, "a quarter-page crop (640x640, G/4) per step, counted as 0.25 page"
        cpu_step(*page)
        w.update(keep)
    for _ in range(max(args.warmup - 1, 0)):
        cpu_step(*page)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(*page)
    dt = time.perf_counter() - t0
    val = frac * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 1280x1280 P2-P5 (A=136000), ~600 GT/page, nc=1; "
                               "NMS conf .25 iou .7 max_det 1000 on ~3000 candidates/page; host CPU, " + note,
                   **{k: w[k] for k in ("imgsz", "G", "nc")}},
        "cpu_baseline": {"value": val, "unit": "pages/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {note} (oracle torch port of the reference, "
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
    try:
        all_cpus = os.sched_getaffinity(0)
    except Exception:
        all_cpus = None
    numa_node = kdist.bind_to_gpu_numa(local_rank) if world > 1 else None  # before the pinned buffers are allocated
    w = WORKLOAD
    b, nc, G = w["batch"], w["nc"], w["G"]
    glo, ghi = kdist.shard_range(b * world, rank, world)  # this rank's pages of the global batch
    warmup = max(args.warmup, 3)

    lv_cpu, gl, gb, mg = page_inputs(b, seed=glo)
    host_lv = [l.pin_memory() for l in lv_cpu]
    host_gt = [t.pin_memory() for t in (gl, gb, mg)]
    A = synth.num_anchors(w["imgsz"], w["strides"])
    # one kernel decodes the head maps once for both halves (validation-during-training flow); set
    # KZD_BENCH_FUSE=0 to time the two separate decode kernels of the drop-in signatures instead
    fuse = os.environ.get("KZD_BENCH_FUSE", "1") != "0"
    # the inference decode also does the NMS filter stage on the values it has just computed (the Detect ->
    # non_max_suppression flow with a known conf_thres); KZD_BENCH_FILTER=0 keeps the separate filter pass
    fuse_filter = os.environ.get("KZD_BENCH_FILTER", "1") != "0"
    pipe_kw = dict(device=dev, topk=w["topk"], alpha=w["alpha"], beta=w["beta"], conf_thres=w["conf"],
                   iou_thres=w["iou"], max_det=w["max_det"], fuse_filter=fuse_filter)
    pipe = kzd.PagePipeline(b, synth.level_shapes(w["imgsz"], w["strides"]), w["strides"], nc, G, fuse_decodes=fuse,
                            **pipe_kw)
    dv_lv = [l.to(dev) for l in host_lv]
    pred_distri, pd_scores = train_layout(dv_lv, nc)
    pipe.load(levels=dv_lv, pred_distri=pred_distri, pd_scores=pd_scores, gt_labels=host_gt[0], gt_bboxes=host_gt[1],
              mask_gt=host_gt[2])
    del pred_distri, pd_scores
    counters = torch.zeros(4, dtype=torch.float64, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # ---- device-resident timing: inputs already in HBM; a step is one CUDA-graph replay. Two pipelines
    # (two input/output buffer sets, same page batch) are stepped alternately on two streams, so batch
    # i+1's HBM-bound decodes overlap batch i's issue-bound assign / NMS kernels (double buffering). ----
    depth = max(1, int(os.environ.get("KZD_BENCH_DEPTH", "4")))
    pipes = [pipe]
    for _ in range(depth - 1):
        pn = kzd.PagePipeline(b, synth.level_shapes(w["imgsz"], w["strides"]), w["strides"], nc, G,
                              fuse_decodes=fuse, **pipe_kw)
        pn.load(levels=pipe.levels, pred_distri=pipe.pred_distri, pd_scores=pipe.pd_scores, gt_labels=pipe.gt_labels,
                gt_bboxes=pipe.gt_bboxes, mask_gt=pipe.mask_gt)
        pipes.append(pn)
    streams = [torch.cuda.Stream(device=dev) for _ in range(max(depth, 2))]
    cur = torch.cuda.current_stream(dev)

    def run_steps(k):
        for s_ in streams:
            s_.wait_stream(cur)
        for i in range(k):
            pipes[i % depth].step(streams[i % depth])
        for s_ in streams:
            cur.wait_stream(s_)

    run_steps(warmup + (-warmup) % depth)
    out, n_keep = pipe.results()
    for pn in pipes[1:]:
        assert pn.results()[1] == n_keep
    kdist.all_reduce_counters(counters)  # warm the collective (communicator setup) outside the timed region
    torch.cuda.synchronize()
    kdist.barrier()
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        t_start, t_end = ev(), ev()
        t_start.record()
        run_steps(args.steps)
        counters[0:2] = out["counters"]
        counters[2] = float(b)
        counters[3] = float(sum(n_keep))
        kdist.all_reduce_counters(counters)  # the path's only collective: 32 bytes
        t_end.record()
        torch.cuda.synchronize()
    kdist.barrier()
    out, n_keep = pipe.results()
    ms_local = t_start.elapsed_time(t_end)
    ms_total = kdist.max_over_ranks(ms_local, dev)
    per_rank_ms = kdist.all_gather_summaries(torch.tensor([ms_local / args.steps], dtype=torch.float64, device=dev))
    per_rank_ms = [round(float(v), 4) for v in per_rank_ms.flatten().tolist()]
    pages = float(counters[2].item())
    value = pages * args.steps / (ms_total * 1e-3)
    # single pipeline, one stream (latency of one step)
    t0s, t1s = ev(), ev()
    t0s.record()
    for _ in range(50):
        pipe.step()
    t1s.record()
    torch.cuda.synchronize()
    ms_single = t0s.elapsed_time(t1s) / 50
    del pipes[1:]

    # ---- per-stage device times: each stage captured as its own CUDA graph and replayed back to back
    # (pure device time of that stage's kernels; decode stages are a single kernel launch each) ----
    import ctypes as C

    from kuzushiji_vision_b200 import _lib

    asg = pipe.assigner
    st_names = ["decode_dual", "assign", "nms"] if fuse else ["decode_train", "assign", "decode_infer", "nms"]
    y_static = kzd.decode_inference(pipe.levels, w["strides"], nc)
    px_static = pipe.out["pd_bboxes"]

    def nms_calls():
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if not fuse_filter:
            _lib.check(_lib.lib.kzd_nms_filter(y_static.data_ptr(), b, A, nc, w["conf"], 0, None, 0, 0,
                                               pipe.keys.data_ptr(), pipe.rows.data_ptr(), pipe.counts[0].data_ptr(),
                                               pipe.pool, st), "kzd_nms_filter")
        _lib.check(_lib.lib.kzd_nms_suppress(pipe.keys.data_ptr(), pipe.rows.data_ptr(), pipe.counts[0].data_ptr(), b,
                                             pipe.pool, pipe.n_cap, w["iou"], 0, w["max_det"], 30000, 7680.0,
                                             pipe.det_rows.data_ptr(), pipe.counts[1].data_ptr(),
                                             pipe.nms_ws.data_ptr(), pipe.ws_bytes, st), "kzd_nms_suppress")

    pool_arg = (w["conf"], pipe.keys, pipe.rows, pipe.counts[0]) if fuse_filter else None
    stage_fns = {
        "decode_dual": lambda: kzd.decode_inference(pipe.levels, w["strides"], nc, with_train_boxes=True,
                                                    nms_pool=pool_arg),
        "decode_train": lambda: kzd.bbox_decode(pipe.anchor_points, pipe.pred_distri, pipe.stride_tensor),
        "assign": lambda: asg(pipe.pd_scores, px_static, pipe.anc_points, pipe.gt_labels, pipe.gt_bboxes,
                              pipe.mask_gt),
        "decode_infer": lambda: kzd.decode_inference(pipe.levels, w["strides"], nc, nms_pool=pool_arg),
        "nms": nms_calls,
    }
    st_ms = {}
    n_prof = 20
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2: evict between stage replays
    for n in st_names:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            stage_fns[n]()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            keep_alive = stage_fns[n]()
        tot = 0.0
        for i in range(n_prof + 2):
            flush.zero_()
            e0, e1 = ev(), ev()
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                tot += e0.elapsed_time(e1)
        st_ms[n] = tot / n_prof
        del g, keep_alive
    del flush

    # ---- end to end: head maps + GT start in pinned HOST memory every step. H2D into a pipeline's static
    # buffers, the reference's caller glue (train layout + sigmoid, torch), one graph replay, and the
    # detections + counters copied back to the host. Two pipelines / streams: batch i+1's H2D overlaps
    # batch i's kernels; every step's results are read on the host before its buffers are reused. ----
    pipe_b = kzd.PagePipeline(b, synth.level_shapes(w["imgsz"], w["strides"]), w["strides"], nc, G,
                              fuse_decodes=fuse, **pipe_kw)
    e_pipes = [pipe, pipe_b]
    det_host = [torch.empty((b, w["max_det"], 6)).pin_memory() for _ in range(2)]
    cnt_host = [torch.empty(2, dtype=torch.float64).pin_memory() for _ in range(2)]

    def e2e_enqueue(k):
        pp, st_ = e_pipes[k], streams[k]
        with torch.cuda.stream(st_):
            pp.load(levels=host_lv, gt_labels=host_gt[0], gt_bboxes=host_gt[1], mask_gt=host_gt[2])  # H2D
            off = 0
            for lvl in pp.levels:  # utils/loss.py:210-215,235: split / permute / sigmoid (caller glue, torch)
                hw = lvl.shape[2] * lvl.shape[3]
                v = lvl.view(b, 64 + nc, hw)
                if not fuse:  # the fused decode reads the NCHW maps directly
                    pp.pred_distri[:, off:off + hw].copy_(v[:, :64].permute(0, 2, 1))
                pp.pd_scores[:, off:off + hw].copy_(v[:, 64:].permute(0, 2, 1).sigmoid())
                off += hw
            out_ = pp.step(st_)
            det_host[k].copy_(pp.det_rows, non_blocking=True)         # D2H
            cnt_host[k].copy_(out_["counters"], non_blocking=True)
            pp.done.record(st_)

    def e2e_collect(k):
        _, keep = e_pipes[k].results()  # waits for that step, reads the kept counts (pinned)
        return int(sum(keep)), float(cnt_host[k][1])

    def e2e_run(n_steps):
        for s_ in streams:
            s_.wait_stream(cur)
        got = None
        for i in range(n_steps):
            if i >= 2:
                got = e2e_collect(i % 2)
            e2e_enqueue(i % 2)
        for k in ((n_steps) % 2, (n_steps + 1) % 2):
            got = e2e_collect(k)
        for s_ in streams:
            cur.wait_stream(s_)
        return got

    e2e_run(4)
    torch.cuda.synchronize()
    kdist.barrier()
    e2e_steps = max(4, min(args.steps, 20))
    e2e_reps = []  # two timed repetitions, the faster one is reported (host-side hiccups on a shared box)
    for _ in range(2):
        kdist.barrier()
        t0, t1 = ev(), ev()
        t0.record()
        e2e_last = e2e_run(e2e_steps)
        t1.record()
        torch.cuda.synchronize()
        e2e_reps.append(kdist.max_over_ranks(t0.elapsed_time(t1), dev))
    e2e_ms = min(e2e_reps)
    e2e_value = pages * e2e_steps / (e2e_ms * 1e-3)
    h2d = sum(l.numel() * 4 for l in host_lv) + sum(t.numel() * 4 for t in host_gt)
    d2h = det_host[0].numel() * 4 + 2 * 8 + b * 4
    assert e2e_last[0] == int(sum(n_keep)), (e2e_last, sum(n_keep))
    del pipe_b, e_pipes

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ----
    n_cand = int(sum(pipe.candidate_counts()))
    ab = algorithmic_bytes(b, A, G, nc, n_cand, int(sum(n_keep)))
    ff = "true" if fuse_filter else "false"
    single_kernel = ({"decode_dual": f"decode_infer_kernel<4, float, true, {ff}>"} if fuse else
                     {"decode_train": "decode_train_kernel", "decode_infer": f"decode_infer_kernel<4, float, false, {ff}>"})
    dom = max(single_kernel, key=lambda n: st_ms[n])
    peak, peak_src = hbm_peak()
    achieved = ab[dom] / (st_ms[dom] * 1e-3) / 1e9
    total_ab = sum(ab[n] for n in st_names if ab.get(n))
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        traffic = next(v["traffic_bytes"] for k, v in tj.items() if k.replace(" ", "") == single_kernel[dom].replace(" ", ""))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": single_kernel[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ab[dom],
                "step": {"algorithmic_bytes": total_ab, "gbs": total_ab / (ms_total / args.steps * 1e-3) / 1e9,
                         "frac": total_ab / (ms_total / args.steps * 1e-3) / 1e9 / peak},
                "stages": {n: {"ms": st_ms[n], "algorithmic_bytes": ab.get(n),
                               "gbs": (ab[n] / (st_ms[n] * 1e-3) / 1e9) if ab.get(n) else None}
                           for n in st_names},
                "stage_note": "stage ms = device time of that stage's kernels (own CUDA graph, L2 flushed before each "
                              "replay, CUDA events); the timed region replays all four stages as one graph"}

    # ---- CPU baseline on a bounded sample (one page, repeated) ----
    if all_cpus is not None and numa_node is not None:
        os.sched_setaffinity(0, all_cpus)  # the CPU arm may use every host core again
    threads = cpu_threads()
    torch.set_num_threads(threads)
    page = page_inputs(1, 0)
    t0c = time.perf_counter()
    n_cpu = 0
    while n_cpu < 2 or (time.perf_counter() - t0c < 12.0 and n_cpu < 8):
        cpu_step(*page)
        n_cpu += 1
    cpu_dt = time.perf_counter() - t0c
    cpu_baseline = {"value": n_cpu / cpu_dt, "unit": "pages/s", "cores": threads, "kind": "port",
                    "sample": f"{n_cpu} pages of the workload, one at a time (oracle torch port of the reference's "
                              f"CPU path, {threads} threads, no warm-up discarded)"}

    line = {
        "metric": METRIC, "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: YOLOv12x P2-P5 head outputs, 1280x1280 pages, batch 16 per GPU, "
                               "~600 GT/page, nc=1; NMS conf .25 iou .7 max_det 1000 on ~3000 candidates/page",
                   "pages_per_gpu": b, "anchors_per_page": A, "gt_slots": G, "nc": nc,
                   "l2_policy": "inputs larger than L2 (0.6-1.2 GB read per step vs 126 MB L2)",
                   "decodes": ("fused: one kernel reads the head maps once and emits y + the training-side boxes"
                               if fuse else "separate decode_train (b,A,64) and decode_infer (NCHW maps) kernels"),
                   "nms_filter": ("inside the inference decode kernel (no second pass over y); the kernel's roofline "
                                  "figure counts the decode's algorithmic bytes only" if fuse_filter else
                                  "separate nms_filter_kernel pass over y"),
                   "step": "PagePipeline: one CUDA-graph replay of decode_train+assign+decode_infer+NMS per step; "
                           f"{depth} pipelines alternate on {depth} streams (multi-buffering)",
                   "ms_per_step_single_stream": ms_single,
                   "parallelism": f"image-sharded x{world}", "numa_node_rank0": numa_node},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_value, "unit": "pages/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step_repetitions": [round(v / e2e_steps, 3) for v in e2e_reps],
                "path": "pinned host head maps + GT -> H2D -> caller glue (torch) -> PagePipeline graph "
                                            "-> detections + counters D2H; two buffer sets / streams; faster of two repetitions"},
        "gpu_launches": pipe.kernels_per_step * args.steps,
        "per_rank_ms_per_step": per_rank_ms,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "counters": {"target_scores_sum": float(counters[0]), "n_fg": float(counters[1]),
                     "kept_boxes": float(counters[3]), "nms_candidates": n_cand},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
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
