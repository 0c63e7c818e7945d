#!/usr/bin/env python
"""bench.py - pages/s of the dense-detection hot path (train decode + assign + inference decode + NMS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl kzd|reference]

One "step" = one pass of the hot path over one GLOBAL batch of synthetic kuzushiji pages, from the per-level
head maps a model forward emits to assignment targets and final detections:
    training-side decode  bbox_decode (utils/loss.py:197) x stride  ->  TaskAlignedAssigner.forward
    inference decode      Detect._inference (nn/modules/head.py:100)  ->  ops.non_max_suppression
Both decodes consume the same head outputs (the validation-during-training flow,
engine/validator.py:180-185), so ONE kernel reads the maps once and emits both (PagePipeline fuse_decodes;
KZD_BENCH_FUSE=0 times the two separate decode kernels of the drop-in signatures).

Workload of the headline line = BASELINE.json configs[4], the sweep its metric ("... 1/2/4/8 B200") is quoted
on: 1280x1280 P2-P5 (A = 136,000), global batch 128 pages, ~1000 GT boxes/page, single class; NMS at the
reference's kuzushiji test-script settings (conf 0.25, iou 0.7, max_det 1000: scripts/test_character_detection.py
:237-243) on inference-like class logits (~3000 candidates/page). The global batch is
image-sharded over the ranks (strong scaling: 128/N pages per GPU, what DistributedSampler + batch //
world_size do: data/build.py:140, engine/trainer.py:286) and every rank runs its pages as micro-batches of
16 (one CUDA-graph replay each; 16 pages = the per-GPU batch of the 8-GPU point). No data-path collective;
one 32-byte all-reduce of counters per step, outside the kernels' critical path.
At N = 1 the line also carries `config1`: the same measurement on BASELINE configs[1] (batch 16, ~600
GT/page), the single-GPU configuration of round 1.

Prints ONE JSON line (rank 0). `value` = pages/s with inputs resident in HBM (CUDA events on every rank, ranks
aligned on the device before each window, max over ranks, median of several windows of exactly K steps);
`e2e` = the same step fed from pinned HOST buffers with the H2D copy and the D2H of the results inside the
timed region; `roofline` = dominant kernel vs MEASURED_PEAKS.json; `cpu_baseline` = the reference's CPU path
timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pages/sec assign+decode+NMS @1280^2 P2-P5"
FALLBACK_HBM_GBS = 6650.0
COMMON = dict(imgsz=1280, nc=1, strides=(4, 8, 16, 32), topk=10, alpha=0.5, beta=6.0, conf=0.25, iou=0.7, max_det=1000,
              nms_candidates=3000, micro=16)
CFG5 = dict(COMMON, name="configs[4]", G=1000, global_pages=128, scaling="strong")
CFG1 = dict(COMMON, name="configs[1]", G=600, pages_per_gpu=16, scaling="weak")
# context lines at N = 1 (stage times only): the multi-class head of configs[2] (batch 4 of its 16: the dense (b,A,nc)
# tensors of four pages already total ~70 GB with the timing buffers) and the inference path of configs[3]
CFG3 = dict(COMMON, name="configs[2]", nc=4300, G=600, micro=4, pages_per_gpu=4, scaling="weak")
CFG4 = dict(COMMON, name="configs[3]", imgsz=1536, G=600, micro=32, pages_per_gpu=32, max_det=2000, scaling="weak")
N_STREAMS = 4   # micro-batches in flight per GPU
N_WINDOWS = 5   # timed windows of exactly --steps steps; the median window is reported


def load_synth():
    """kuzushiji_vision_b200/synth.py by path: the reference arm must not import the package (it loads libkzd)."""
    spec = importlib.util.spec_from_file_location("kzd_synth", os.path.join(ROOT, "kuzushiji_vision_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload_config(cfg, world):
    """The `config` object both arms print (same keys, same workload string)."""
    if cfg["scaling"] == "strong":
        pages = f"global batch {cfg['global_pages']} pages image-sharded over the GPUs"
    else:
        pages = f"batch {cfg['pages_per_gpu']} pages per GPU"
    pages_global = cfg.get("global_pages", cfg.get("pages_per_gpu", 0) * world)
    return {
        "workload": f"BASELINE {cfg['name']}: YOLOv12x P2-P5 head outputs, {cfg['imgsz']}x{cfg['imgsz']} pages, {pages}, "
                    f"~{cfg['G']} GT/page, nc={cfg['nc']}; NMS conf {cfg['conf']} iou {cfg['iou']} max_det {cfg['max_det']} "
                    f"on ~{cfg['nms_candidates']} candidates/page",
        "imgsz": cfg["imgsz"], "gt_slots": cfg["G"], "nc": cfg["nc"], "micro_batch_pages": cfg["micro"],
        "global_pages": pages_global, "pages_per_gpu": pages_global // world, "parallelism": f"image-sharded x{world}",
        "l2_policy": "inputs larger than L2 (0.6-1.2 GB read per 16-page micro-batch vs 126 MB L2); CPU arm: not applicable",
    }


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def csrc_digest():
    """sha1 over the kernel sources: ties a committed ncu traffic figure to the code it was captured from."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "kuzushiji_vision_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~2 ms while the timed region runs
    (nvidia-smi -lms 200 is too coarse for a region of a few tens of milliseconds). start() is called before
    the ranks are aligned, so NVML initialisation never sits between the barrier and the timed region."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self._stop, self.t, self.on = index, [], set(), None, False, None, False

    def start(self):
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
            if self.on:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for n, bit in bits.items():
                        if r & bit:
                            self.reasons.add(n)
                except Exception:
                    pass
            time.sleep(0.002 if self.on else 0.001)

    def stop(self):
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
        # nc = 1: (57 + 8/b) A + 24 G; nc > 1: (49 + 4 nc + 8/b) A + 4 P + 24 G - the dense (b,A,nc)
        # target_scores has to be written; P = in-GT pairs whose score is gathered, ~90 per GT on these pages
        "assign": int((57 + 8 / b) * A + 24 * G) * b if nc == 1 else int((49 + 4 * nc + 8 / b) * A + 4 * 90 * G + 24 * G) * b,
        "decode_infer": (272 + 8 * nc) * A * b,
        # fused: one read of the maps 4(64+nc)A, writes y 4(4+nc)A + grid-unit boxes 16A + pixel boxes 16A
        "decode_dual": (4 * (64 + nc) + 4 * (4 + nc) + 32) * A * b,
        "nms": int((32 + 4 * nc) * A * b + 24 * (n_cand + n_keep)),
    }


def page_inputs(cfg, n_pages, seed, device="cpu", on_device=False, synth=None):
    """Synthetic head outputs + GT for `n_pages` pages. ONE set of per-level head maps feeds both halves of
    the step (as a model forward would): DFL logits ~ N(0,1); class logits like a freshly initialised Detect
    head (bias_init drives them strongly negative: background -6) with ~3000 object-like anchors per page at
    N(1,1). The training-side layout (pred_distri, pd_scores) is derived from the maps with the reference's
    own glue ops (utils/loss.py:210-215,235)."""
    synth = synth or load_synth()
    lv = synth.head_outputs(n_pages, cfg["nc"], cfg["imgsz"], cfg["strides"], seed=seed, device=device,
                            sparse_cls=cfg["nms_candidates"], on_device=on_device)
    gl, gb, mg = synth.gt_pages(n_pages, cfg["G"], cfg["imgsz"], cfg["nc"], seed=seed, device=device)
    return lv, gl, gb, mg


def train_layout(lv, nc):
    """utils/loss.py:210-215 + :235: cat levels, split, permute, contiguous, sigmoid (caller glue)."""
    b = lv[0].shape[0]
    cat = torch.cat([x.view(b, 64 + nc, -1) for x in lv], 2)
    pred_distri = cat[:, :64].permute(0, 2, 1).contiguous()
    pd_scores = cat[:, 64:].permute(0, 2, 1).contiguous().sigmoid()
    return pred_distri, pd_scores


# ------------------------------------------------------------------------------------------------
# the reference's own implementation of the path (CPU arm; also the same-GPU torch-eager context number)
# ------------------------------------------------------------------------------------------------

class RefPath:
    """The four stages of one step through the reference's own functions: the UNMODIFIED package staged under
    oracle/_ref by oracle/make_ref.py (kind "reference") or, when no copy is importable, the oracle's torch
    port of the same code (kind "port")."""

    def __init__(self, cfg):
        from oracle import ref_loader

        self.cfg = cfg
        self.kind = "reference" if ref_loader.available() else "port"
        if self.kind == "reference":
            self.R = ref_loader.load()
            self.assigner = self.R.TaskAlignedAssigner(topk=cfg["topk"], num_classes=cfg["nc"], alpha=cfg["alpha"],
                                                       beta=cfg["beta"])
            self.det = None
        else:
            from oracle import torch_oracle

            self.O = torch_oracle

    def step(self, lv, gl, gb, mg):
        c = self.cfg
        nc, strides = c["nc"], c["strides"]
        pred_distri, pd_scores = train_layout(lv, nc)
        if self.kind == "port":
            O = self.O
            anchor_points, stride_tensor = O.make_anchors([x.shape[2:] for x in lv], strides, device=lv[0].device)
            boxes = O.decode_train(pred_distri, anchor_points)
            out = O.tal_assign(pd_scores, boxes * stride_tensor, anchor_points * stride_tensor, gl, gb, mg, c["topk"], nc,
                               c["alpha"], c["beta"])
            y = O.decode_infer(lv, strides, nc)
            det = O.non_max_suppression(y, c["conf"], c["iou"], max_det=c["max_det"])
            return out, det
        R = self.R
        dev = lv[0].device
        anchor_points, stride_tensor = R.make_anchors(lv, strides, 0.5)                      # utils/loss.py:220
        proj = torch.arange(16, dtype=torch.float, device=dev)
        b, a, ch = pred_distri.shape
        dist = pred_distri.view(b, a, 4, ch // 4).softmax(3).matmul(proj)                    # utils/loss.py:199-201
        boxes = R.dist2bbox(dist, anchor_points, xywh=False)                                 # utils/loss.py:204
        with torch.no_grad():
            out = self.assigner.to(dev)(pd_scores, (boxes * stride_tensor).type(gb.dtype), anchor_points * stride_tensor,
                                        gl, gb, mg)                                          # utils/loss.py:233-241
            if self.det is None or self.det.stride.device != dev:
                self.det = R.Detect(nc=nc, ch=tuple(16 for _ in strides)).to(dev).eval()
                self.det.stride = torch.tensor([float(s) for s in strides], device=dev)
            y = self.det._inference([x for x in lv])                                         # nn/modules/head.py:100
            det = R.non_max_suppression(y, c["conf"], c["iou"], max_det=c["max_det"])        # detect/predict.py:25
        return out, det


def cpu_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return max(1, n)


def time_cpu_pages(ref, page, threads, budget_s, min_pages=1, max_pages=8):
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    n = 0
    while n < min_pages or (time.perf_counter() - t0 < budget_s and n < max_pages):
        ref.step(*page)
        n += 1
    return n, time.perf_counter() - t0


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path on this box's host cores - the
    unmodified package staged under oracle/_ref when it is importable, else the oracle's torch port - with all
    the host threads it can use, one page of the kzd arm's workload per step. If K full pages would not fit in
    ~4 minutes the page is cropped to a quarter (640x640, G/4) and each step counts as 0.25 page: the
    reference's cost grows faster than linearly with page size, so the crop only flatters the CPU arm."""
    if rank != 0:
        return
    cfg = dict(CFG5)
    threads = cpu_threads()
    torch.set_num_threads(threads)
    ref = RefPath(cfg)
    synth = load_synth()
    page = page_inputs(cfg, 1, 0, synth=synth)
    ref.step(*page)  # warm-up: imports, torchvision op registration, allocator
    t0 = time.perf_counter()
    ref.step(*page)
    t_page = time.perf_counter() - t0
    frac, note = 1.0, "1 full page per step"
    if t_page * (args.steps + max(args.warmup - 2, 0)) > 240.0:
        small = dict(cfg, imgsz=cfg["imgsz"] // 2, G=cfg["G"] // 4, nms_candidates=cfg["nms_candidates"] // 4)
        ref = RefPath(small)
        page = page_inputs(small, 1, 0, synth=synth)
        frac, note = 0.25, "a quarter-page crop (640x640, G/4) per step, counted as 0.25 page"
        ref.step(*page)
    for _ in range(max(args.warmup - 2, 0)):
        ref.step(*page)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step(*page)
    dt = time.perf_counter() - t0
    val = frac * args.steps / dt
    # the reference's own CPU thread policy (ultralytics/utils/__init__.py:43: min(8, cores - 1)), context row
    pol = max(1, min(8, (os.cpu_count() or 2) - 1))
    n_pol, dt_pol = time_cpu_pages(ref, page, pol, budget_s=10.0)
    what = ("the unmodified reference package (oracle/_ref)" if ref.kind == "reference"
            else "oracle torch port of the reference (no staged copy importable)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, world),
        "cpu_baseline": {"value": val, "unit": "pages/s", "cores": threads, "kind": ref.kind,
                         "sample": f"{args.steps} steps x {note}; {what}, torch CPU, {threads} threads",
                         "reference_thread_policy": {"threads": pol, "value": frac * n_pol / dt_pol, "unit": "pages/s",
                                                     "note": "min(8, cores-1), ultralytics/utils/__init__.py:43"}},
        "e2e": {"value": val, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------

def ev():
    return torch.cuda.Event(enable_timing=True)


class Measure:
    """One workload on this rank: `m` micro-batches of `micro` pages per step, max(m, N_STREAMS) graph-captured
    PagePipelines (own static buffers each), stepped round-robin on N_STREAMS streams so that one micro-batch's
    issue-bound assign / NMS kernels overlap the HBM-bound decode of the others."""

    def __init__(self, kzd, synth, cfg, m, dev, seed0, fuse, fuse_filter, levels_dtype=torch.float32, n_pipes=None):
        self.kzd, self.cfg, self.m, self.dev = kzd, cfg, m, dev
        self.b, self.nc, self.G = cfg["micro"], cfg["nc"], cfg["G"]
        self.shapes = synth.level_shapes(cfg["imgsz"], cfg["strides"])
        self.A = synth.num_anchors(cfg["imgsz"], cfg["strides"])
        self.fuse, self.fuse_filter = fuse, fuse_filter
        self.kw = dict(device=dev, topk=cfg["topk"], alpha=cfg["alpha"], beta=cfg["beta"], conf_thres=cfg["conf"],
                       iou_thres=cfg["iou"], max_det=cfg["max_det"], fuse_decodes=fuse, fuse_filter=fuse_filter,
                       levels_dtype=levels_dtype)
        self.n_pipes = n_pipes or max(m, N_STREAMS)
        self.inputs = []  # device-resident inputs of the m distinct micro-batches
        for j in range(m):
            lv, gl, gb, mg = page_inputs(cfg, self.b, seed0 + j * self.b, device=dev, on_device=True, synth=synth)
            self.inputs.append(([x.to(levels_dtype) for x in lv], gl, gb, mg))
        self.pipes = []
        for j in range(self.n_pipes):
            p = kzd.PagePipeline(self.b, self.shapes, cfg["strides"], self.nc, self.G, **self.kw)
            lv, gl, gb, mg = self.inputs[j % m]
            pred_distri, pd_scores = train_layout(lv, self.nc)
            p.load(levels=lv, pred_distri=pred_distri, pd_scores=pd_scores, gt_labels=gl, gt_bboxes=gb, mask_gt=mg)
            self.pipes.append(p)
        del pred_distri, pd_scores
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(N_STREAMS)]
        self.cur = torch.cuda.current_stream(dev)
        self.k = 0

    def run_steps(self, steps):
        for s_ in self.streams:
            s_.wait_stream(self.cur)
        for _ in range(steps * self.m):
            self.pipes[self.k % self.n_pipes].step(self.streams[self.k % N_STREAMS])
            self.k += 1
        for s_ in self.streams:
            self.cur.wait_stream(s_)

    @property
    def launches_per_step(self):
        return self.pipes[0].kernels_per_step * self.m

    # -- parity of the benchmarked configuration, before anything is timed ---------------------------
    def parity_check(self, pages=(0,)):
        """Pipeline outputs of micro-batch 0 against the oracle (torch restatement of the reference, same
        device, same inputs): all five assigner outputs and every NMS row of the checked pages."""
        from oracle import torch_oracle as O

        p = self.pipes[0]
        p.step()
        out, n_keep = p.results()
        lv, gl, gb, mg = self.inputs[0]
        anchor_points, stride_tensor = O.make_anchors(self.shapes, self.cfg["strides"], device=self.dev)
        anc = anchor_points * stride_tensor
        c = self.cfg
        for i in pages:
            sl = slice(i, i + 1)
            ref = O.tal_assign(p.pd_scores[sl].float(), out["pd_bboxes"][sl], anc, gl[sl], gb[sl], mg[sl], c["topk"], self.nc,
                               c["alpha"], c["beta"])
            mine = (out["target_labels"][sl], out["target_bboxes"][sl], out["target_scores"][sl], out["fg_mask"][sl],
                    out["target_gt_idx"][sl])
            for a_, r_ in zip((mine[0], mine[1], mine[3], mine[4]), (ref[0], ref[1], ref[3], ref[4])):
                if not torch.equal(a_, r_):
                    raise SystemExit(f"bench.py: parity check failed (assign, page {i})")
            torch.testing.assert_close(mine[2], ref[2], rtol=1e-5, atol=1e-9)
            y = self.kzd.decode_inference([x[sl] for x in lv], c["strides"], self.nc)
            det = O.non_max_suppression(y, c["conf"], c["iou"], max_det=c["max_det"])[0]
            got = out["det_rows"][i, :n_keep[i]]
            if det.shape != got.shape or not torch.equal(det, got):
                raise SystemExit(f"bench.py: parity check failed (NMS rows, page {i})")
        return {"pages_checked": list(pages), "against": "oracle/torch_oracle (torch restatement of the reference) on the "
                "same device and inputs: 5 assigner outputs + every NMS row", "workload": c["name"]}

    # -- device-resident timing ---------------------------------------------------------------------
    def timed_windows(self, kdist, steps, warmup, world):
        """N_WINDOWS windows of exactly `steps` steps. Before every window the ranks meet ON THE DEVICE (a tiny
        all-reduce enqueued on the timing stream: no rank's first kernel starts before the slowest rank's
        host got here), then start/stop events bracket the window on every rank."""
        dev = self.dev
        token = torch.zeros(1, device=dev)
        self.run_steps(warmup)
        torch.cuda.synchronize()
        kdist.all_reduce_counters(token)
        torch.cuda.synchronize()
        kdist.barrier()
        marks = []
        for _ in range(N_WINDOWS):
            kdist.all_reduce_counters(token)  # device-side rendezvous
            t0, t1 = ev(), ev()
            t0.record()
            self.run_steps(steps)
            t1.record()
            marks.append((t0, t1))
        torch.cuda.synchronize()
        kdist.barrier()
        local = torch.tensor([a.elapsed_time(b_) for a, b_ in marks], dtype=torch.float64, device=dev)
        allr = kdist.all_gather_summaries(local)          # (world, windows)
        per_window = allr.max(0).values.tolist()          # max over ranks, per window
        med = statistics.median(per_window)
        wi = min(range(len(per_window)), key=lambda i: abs(per_window[i] - med))
        return {"ms_window": per_window[wi], "windows_ms": [round(v, 4) for v in per_window],
                "per_rank_ms_per_step": [round(float(v) / steps, 5) for v in allr[:, wi].tolist()]}

    def single_stream_ms(self, reps=200):
        t0, t1 = ev(), ev()
        t0.record()
        for r in range(reps * self.m):
            self.pipes[r % self.n_pipes].step()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / reps

    # -- per-stage device times ----------------------------------------------------------------------
    def stage_times(self):
        """Each stage captured as its own CUDA graph and replayed with L2 flushed before every replay (pure
        device time of that stage's kernels; decode stages are a single kernel launch each)."""
        import ctypes as C

        from kuzushiji_vision_b200 import _lib

        kzd, dev, c = self.kzd, self.dev, self.cfg
        pipe = self.pipes[0]
        b, A, nc = self.b, self.A, self.nc
        asg = pipe.assigner
        names = ["decode_dual", "assign", "nms"] if self.fuse else ["decode_train", "assign", "decode_infer", "nms"]
        # the assigner as the eager drop-in runs it (a lone call: a warp per GT); "assign" is the form inside the timed
        # pipelines (two GTs per warp). Same results, reported side by side; only "assign" counts towards the step's bytes.
        asg_lone = kzd.TaskAlignedAssigner(c["topk"], nc, c["alpha"], c["beta"])
        asg_lone.per_stream = False  # like the pipeline's: the zeroed-once workspace of the warm-up call serves the captured one
        timed = names + ["assign_lone_call"]
        y_static = kzd.decode_inference(pipe.levels, c["strides"], nc)
        px_static = pipe.out["pd_bboxes"]

        def nms_calls():
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if not self.fuse_filter:
                _lib.check(_lib.lib.kzd_nms_filter(y_static.data_ptr(), b, A, nc, c["conf"], 0, None, 0, 0,
                                                   pipe.keys.data_ptr(), pipe.rows.data_ptr(), pipe.counts[0].data_ptr(),
                                                   pipe.pool, st), "kzd_nms_filter")
            _lib.check(_lib.lib.kzd_nms_suppress(pipe.keys.data_ptr(), pipe.rows.data_ptr(), pipe.counts[0].data_ptr(), b,
                                                 pipe.pool, pipe.n_cap, c["iou"], 0, c["max_det"], 30000, 7680.0,
                                                 pipe.det_rows.data_ptr(), pipe.counts[1].data_ptr(),
                                                 pipe.nms_ws.data_ptr(), pipe.ws_bytes, st), "kzd_nms_suppress")

        pool_arg = (c["conf"], pipe.keys, pipe.rows, pipe.counts[0]) if self.fuse_filter else None
        fns = {
            "decode_dual": lambda: kzd.decode_inference(pipe.levels, c["strides"], nc, with_train_boxes=True, nms_pool=pool_arg),
            "decode_train": lambda: kzd.bbox_decode(pipe.anchor_points, pipe.pred_distri, pipe.stride_tensor),
            "assign": lambda: asg(pipe.pd_scores, px_static, pipe.anc_points, pipe.gt_labels, pipe.gt_bboxes, pipe.mask_gt),
            "assign_lone_call": lambda: asg_lone(pipe.pd_scores, px_static, pipe.anc_points, pipe.gt_labels, pipe.gt_bboxes,
                                                 pipe.mask_gt),
            "decode_infer": lambda: kzd.decode_inference(pipe.levels, c["strides"], nc, nms_pool=pool_arg),
            "nms": nms_calls,
        }
        st_ms, n_prof = {}, 20
        # L2 flush between stage replays: write a buffer larger than L2, then READ another one, so the cache holds
        # clean lines only and the timed kernels do not pay for writing the flush pattern back to HBM
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
        for n in timed:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                fns[n]()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                keep_alive = fns[n]()
            tot = 0.0
            for i in range(n_prof + 2):
                flush.zero_()
                flush_rd.sum()
                e0, e1 = ev(), ev()
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    tot += e0.elapsed_time(e1)
            st_ms[n] = tot / n_prof
            del g, keep_alive
        del flush, flush_rd
        return names, st_ms

    # -- end to end ----------------------------------------------------------------------------------
    def e2e(self, kdist, steps, world):
        """Head maps + GT start in pinned HOST memory every step. Per micro-batch: H2D into a pipeline's static
        buffers, the reference's caller glue (train layout + sigmoid, torch), one graph replay, and the
        detections + counters copied back to the host. Two streams: micro-batch i+1's H2D overlaps micro-batch
        i's kernels; every micro-batch's results are read on the host before its buffers are reused. Also times
        the bare pinned H2D copies of the same buffers (the ceiling this path can reach on this host)."""
        dev, b, nc, m = self.dev, self.b, self.nc, self.m
        host = []
        for lv, gl, gb, mg in self.inputs:
            host.append(([x.cpu().pin_memory() for x in lv], [t.cpu().pin_memory() for t in (gl, gb, mg)]))
        n_slots = 2
        pipes = self.pipes[:n_slots]
        det_host = [torch.empty((b, self.cfg["max_det"], 6)).pin_memory() for _ in range(n_slots)]
        cnt_host = [torch.empty(2, dtype=torch.float64).pin_memory() for _ in range(n_slots)]
        streams = self.streams[:n_slots]
        fuse = self.fuse

        def enqueue(k, j):
            pp, st_ = pipes[k], streams[k]
            h_lv, h_gt = host[j]
            with torch.cuda.stream(st_):
                pp.load(levels=h_lv, gt_labels=h_gt[0], gt_bboxes=h_gt[1], mask_gt=h_gt[2])  # H2D
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

        def collect(k):
            _, keep = pipes[k].results()  # waits for that micro-batch, reads the kept counts (pinned)
            return int(sum(keep)), float(cnt_host[k][1])

        def run(n_steps):
            for s_ in streams:
                s_.wait_stream(self.cur)
            kept, total = 0, n_steps * m
            for i in range(total):
                if i >= n_slots:
                    kept += collect(i % n_slots)[0]
                enqueue(i % n_slots, i % m)
            for i in range(max(total - n_slots, 0), total):
                kept += collect(i % n_slots)[0]
            for s_ in streams:
                self.cur.wait_stream(s_)
            return kept

        def bare_copies(n_steps):
            for s_ in streams:
                s_.wait_stream(self.cur)
            for i in range(n_steps * m):
                with torch.cuda.stream(streams[i % n_slots]):
                    for dst, src in zip(pipes[i % n_slots].levels, host[i % m][0]):
                        dst.copy_(src, non_blocking=True)  # one cudaMemcpyAsync per level
            for s_ in streams:
                self.cur.wait_stream(s_)

        run(1)
        torch.cuda.synchronize()
        reps = []
        for _ in range(2):  # two timed repetitions, the faster one is reported (host-side hiccups on a shared box)
            kdist.barrier()
            t0, t1 = ev(), ev()
            t0.record()
            kept = run(steps)
            t1.record()
            torch.cuda.synchronize()
            reps.append(kdist.max_over_ranks(t0.elapsed_time(t1), dev))
        kdist.barrier()
        t0, t1 = ev(), ev()
        t0.record()
        bare_copies(steps)
        t1.record()
        torch.cuda.synchronize()
        copy_ms = kdist.max_over_ranks(t0.elapsed_time(t1), dev)
        esz = host[0][0][0].element_size()
        h2d_micro = sum(x.numel() * esz for x in host[0][0]) + sum(t.numel() * 4 for t in host[0][1])
        lv_bytes = sum(x.numel() * esz for x in host[0][0])
        d2h_micro = det_host[0].numel() * 4 + 2 * 8 + b * 4
        ms = min(reps)
        # restore the device-resident inputs the timed pipelines started with
        for k in range(n_slots):
            lv, gl, gb, mg = self.inputs[k % m]
            pd, ps = train_layout(lv, nc)
            pipes[k].load(levels=lv, pred_distri=pd, pd_scores=ps, gt_labels=gl, gt_bboxes=gb, mask_gt=mg)
        torch.cuda.synchronize()
        return {"ms": ms, "reps_ms_per_step": [round(v / steps, 3) for v in reps], "steps": steps, "kept_last": kept,
                "h2d_bytes_per_step": h2d_micro * m, "d2h_bytes_per_step": d2h_micro * m,
                "h2d_gbs_ceiling_per_gpu": lv_bytes * m * steps / (copy_ms * 1e-3) / 1e9,
                "h2d_gbs_achieved_per_gpu": h2d_micro * m * steps / (ms * 1e-3) / 1e9}


def stage_entry(name, ms, nbytes, peak, fuse_filter):
    gbs = nbytes / (ms * 1e-3) / 1e9 if nbytes else None
    e = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": gbs, "frac": gbs / peak if gbs else None}
    if name == "nms" and fuse_filter:
        e["note"] = ("SURVEY 8d bytes of the whole non_max_suppression call, (32+4nc) B per anchor of which belong to the "
                     "filter pass over y - done inside the decode kernel here; the kernels timed in this stage (order, mask, "
                     "walk) move ~24 B per candidate and are latency-bound")
        if e["frac"] and e["frac"] > 1.5:  # multi-class head: the filter pass is nearly all of the call's bytes
            e["frac"] = None
    return e


def roofline_block(meas, names, st_ms, ms_per_step, n_cand, n_keep):
    b, A, G, nc, m = meas.b, meas.A, meas.G, meas.nc, meas.m
    ab = algorithmic_bytes(b, A, G, nc, n_cand, n_keep)
    ff = "true" if meas.fuse_filter else "false"
    wide = "true" if nc >= 32 else "false"  # csrc/kzd_decode.cu: WIDE_NC
    single = ({"decode_dual": f"decode_infer_kernel<4, float, true, {ff}, {wide}>"} if meas.fuse else
              {"decode_train": "decode_train_kernel", "decode_infer": f"decode_infer_kernel<4, float, false, {ff}, {wide}>"})
    dom = max(single, key=lambda n: st_ms[n])
    peak, peak_src = hbm_peak()
    achieved = ab[dom] / (st_ms[dom] * 1e-3) / 1e9
    total_ab = sum(ab[n] for n in names if ab.get(n)) * m
    traffic, traffic_note = None, "no committed ncu capture"
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if tj.get("_csrc_digest") != csrc_digest():
            traffic_note = (f"profiles/r2_traffic.json was captured from other kernel sources "
                            f"({tj.get('_csrc_digest')} vs {csrc_digest()}): refused as stale")
        else:
            traffic = next(v["traffic_bytes"] for k, v in tj.items()
                           if not k.startswith("_") and k.replace(" ", "") == single[dom].replace(" ", ""))
            traffic_note = "profiles/r2_traffic.json (ncu --set full, same kernel sources)"
    except Exception:
        pass
    step_gbs = total_ab / (ms_per_step * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": single[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ab[dom], "pages_per_launch": b,
            "step": {"algorithmic_bytes": total_ab, "gbs": step_gbs, "frac": step_gbs / peak},
            "stages": {n: stage_entry(n, st_ms[n], ab.get("assign" if n == "assign_lone_call" else n), peak, meas.fuse_filter)
                       for n in list(names) + [k for k in st_ms if k not in names]},
            "stage_note": "stage ms = device time of that stage's kernels on one micro-batch (own CUDA graph; L2 flushed "
                          "before each replay by writing 256 MB and then reading 256 MB, so no dirty lines are left for "
                          "the stage to evict; CUDA events); the timed region replays all stages as one graph"}


def measure_workload(kzd, kdist, synth, cfg, m, dev, rank, world, steps, warmup, fuse, fuse_filter, clk=None):
    """Everything the line reports for one workload; returns a dict (rank 0 uses it)."""
    pages_rank = m * cfg["micro"]
    glo = rank * pages_rank
    meas = Measure(kzd, synth, cfg, m, dev, seed0=glo, fuse=fuse, fuse_filter=fuse_filter)
    parity = meas.parity_check()
    if clk is not None:  # clocks / throttle reasons are sampled while this step loop keeps the device busy: the warm-up,
        clk.on = True    # the timed windows and the single-stream repetition that follows them (NVML answers in ms)
    win = meas.timed_windows(kdist, steps, warmup, world)
    single_ms = meas.single_stream_ms()
    if clk is not None:
        clk.on = False
    # the path's only collective: 4 counters (target_scores_sum, n_fg, pages, kept boxes), once per step in a
    # training loop; timed on its own, after the step windows (it is not on the kernels' critical path)
    out, n_keep = meas.pipes[0].results()
    counters = torch.zeros(4, dtype=torch.float64, device=dev)
    counters[0:2] = out["counters"]
    counters[2] = float(pages_rank)
    counters[3] = float(sum(n_keep))
    kdist.all_reduce_counters(counters.clone())
    torch.cuda.synchronize()
    kdist.barrier()
    c0, c1 = ev(), ev()
    c0.record()
    for _ in range(10):
        kdist.all_reduce_counters(counters.clone())
    c1.record()
    torch.cuda.synchronize()
    coll_us = c0.elapsed_time(c1) / 10 * 1e3
    kdist.all_reduce_counters(counters)
    pages_global = float(counters[2].item())
    ms_step = win["ms_window"] / steps
    res = {"value": pages_global * steps / (win["ms_window"] * 1e-3), "ms_per_step": ms_step, "pages_global": pages_global,
           "windows_ms_per_step": [round(v / steps, 5) for v in win["windows_ms"]],
           "per_rank_ms_per_step": win["per_rank_ms_per_step"], "parity_checked": parity,
           "counter_all_reduce_us": round(coll_us, 2), "launches_per_step": meas.launches_per_step,
           "ms_per_step_single_stream": single_ms,
           "counters": {"target_scores_sum": float(counters[0]), "n_fg": float(counters[1]), "kept_boxes": float(counters[3])}}
    names, st_ms = meas.stage_times()
    e2e_steps = max(2, min(steps, 20 if m == 1 else 6))
    e = meas.e2e(kdist, e2e_steps, world)
    res["e2e"] = {"value": pages_global * e2e_steps / (e["ms"] * 1e-3), "unit": "pages/s",
                  "h2d_bytes_per_step": e["h2d_bytes_per_step"], "d2h_bytes_per_step": e["d2h_bytes_per_step"],
                  "steps": e2e_steps, "ms_per_step_repetitions": e["reps_ms_per_step"],
                  "h2d_gbs_achieved_per_gpu": round(e["h2d_gbs_achieved_per_gpu"], 2),
                  "h2d_gbs_ceiling_per_gpu": round(e["h2d_gbs_ceiling_per_gpu"], 2),
                  "h2d_ceiling_note": "bare pinned cudaMemcpyAsync of the same head-map buffers (one copy per level), same "
                                      "streams, all ranks at once, timed in this run; e2e within "
                                      f"{100 * (1 - e['h2d_gbs_achieved_per_gpu'] / e['h2d_gbs_ceiling_per_gpu']):.1f} % of it",
                  "path": "pinned host head maps + GT -> H2D -> caller glue (torch) -> PagePipeline graph -> detections + "
                          "counters D2H; two buffer sets / streams; faster of two repetitions"}
    n_cand = int(sum(meas.pipes[0].candidate_counts()))
    res["roofline"] = roofline_block(meas, names, st_ms, ms_step, n_cand, int(sum(n_keep)))
    res["nms_candidates_micro0"] = n_cand
    res["_meas"] = meas
    return res


def context_stages(kzd, synth, cfg, dev, fuse, fuse_filter, only=None):
    """Stage times + roofline fractions of one more BASELINE config (one pipeline, one micro-batch; no step timing)."""
    meas = Measure(kzd, synth, cfg, 1, dev, seed0=0, fuse=fuse, fuse_filter=fuse_filter, n_pipes=1)
    pipe = meas.pipes[0]
    pipe.step()
    _, n_keep = pipe.results()
    n_cand = int(sum(pipe.candidate_counts()))
    names, st_ms = meas.stage_times()
    rb = roofline_block(meas, names, st_ms, sum(st_ms.values()), n_cand, int(sum(n_keep)))
    stages = {n: v for n, v in rb["stages"].items() if only is None or n in only}
    out = {"config": dict(workload_config(cfg, 1), max_det=cfg["max_det"]), "pages_per_launch": cfg["micro"], "stages": stages,
           "nms_candidates": n_cand, "kept_boxes": int(sum(n_keep)), "note": rb["stage_note"]}
    del meas, pipe
    torch.cuda.empty_cache()
    return out


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
    clk = ClockSampler(local_rank).start()  # NVML up before any rank alignment
    warmup = max(args.warmup, 3)
    fuse = os.environ.get("KZD_BENCH_FUSE", "1") != "0"
    fuse_filter = os.environ.get("KZD_BENCH_FILTER", "1") != "0"
    cfg = CFG5
    micro_total = cfg["global_pages"] // cfg["micro"]
    if micro_total % world:
        raise SystemExit(f"bench.py: {micro_total} micro-batches do not divide over {world} GPUs")
    m = micro_total // world
    kdist.all_reduce_counters(torch.zeros(1, device=dev))  # communicator setup, long before anything is timed
    res = measure_workload(kzd, kdist, synth, cfg, m, dev, rank, world, args.steps, warmup, fuse, fuse_filter, clk)
    clk.stop()
    meas = res.pop("_meas")

    extra = {}
    if world == 1:
        # ---- context: fp16 head maps end to end (the AMP / half flow the reference really runs:
        # engine/trainer.py:379, engine/validator.py:115-117); the fp32 line stays the headline ----
        try:
            half = Measure(kzd, synth, cfg, 1, dev, seed0=0, fuse=fuse, fuse_filter=fuse_filter, levels_dtype=torch.float16,
                           n_pipes=2)
            eh = half.e2e(kdist, 6, world)
            extra["e2e_amp_fp16_context"] = {
                "value": cfg["micro"] * 6 / (eh["ms"] * 1e-3), "unit": "pages/s", "h2d_bytes_per_step": eh["h2d_bytes_per_step"],
                "note": "one micro-batch (16 pages) per step, fp16 head maps read natively by the decode kernel; context "
                        "only - the fp32 line is the headline"}
            del half
        except Exception as exc:  # context line only
            extra["e2e_amp_fp16_context"] = {"error": str(exc)[:200]}
        del meas
        torch.cuda.empty_cache()
        # ---- BASELINE configs[1] (round 1's single-GPU configuration), same measurement ----
        r1 = measure_workload(kzd, kdist, synth, CFG1, 1, dev, rank, world, args.steps, warmup, fuse, fuse_filter)
        r1.pop("_meas")
        extra["config1"] = dict(r1, config=workload_config(CFG1, world), unit="pages/s", metric=METRIC)
        torch.cuda.empty_cache()
        # ---- stage lines of BASELINE configs[2] (nc = 4300) and configs[3] (inference path, 1536^2, batch 32) ----
        for key, c_, only in (("config3_stages", CFG3, None), ("config4_stages", CFG4, ("decode_dual", "decode_infer", "nms"))):
            try:
                extra[key] = context_stages(kzd, synth, c_, dev, fuse, fuse_filter, only)
            except Exception as exc:  # context only: never costs the headline line
                extra[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                torch.cuda.empty_cache()
    if rank != 0:
        return

    cpu_baseline = None
    if world == 1:
        # ---- CPU baseline on a bounded sample (pages of the workload, one at a time) ----
        if all_cpus is not None and numa_node is not None:
            os.sched_setaffinity(0, all_cpus)
        threads = cpu_threads()
        ref = RefPath(cfg)
        page = page_inputs(cfg, 1, 0, synth=synth)
        torch.set_num_threads(threads)
        ref.step(*page)  # imports / op registration
        n_cpu, cpu_dt = time_cpu_pages(ref, page, threads, budget_s=10.0)
        pol = max(1, min(8, (os.cpu_count() or 2) - 1))
        n_pol, dt_pol = time_cpu_pages(ref, page, pol, budget_s=8.0)
        what = "the unmodified reference package staged under oracle/_ref" if ref.kind == "reference" else "oracle torch port"
        cpu_baseline = {"value": n_cpu / cpu_dt, "unit": "pages/s", "cores": threads, "kind": ref.kind,
                        "sample": f"{n_cpu} pages of the workload, one at a time after one warm-up page ({what}, torch "
                                  f"CPU, {threads} threads)",
                        "reference_thread_policy": {"threads": pol, "value": n_pol / dt_pol, "unit": "pages/s",
                                                    "note": "min(8, cores-1), ultralytics/utils/__init__.py:43"}}
        # ---- context: the reference's torch-eager path on this same GPU (it is the bit-exactness oracle) ----
        try:
            pg = page_inputs(cfg, 2, 0, device=dev, on_device=True, synth=synth)
            ref.step(*pg)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                ref.step(*pg)
            torch.cuda.synchronize()
            extra["torch_eager_same_gpu"] = {"value": 2 * 3 / (time.perf_counter() - t0), "unit": "pages/s", "kind": ref.kind,
                                             "note": "reference functions (torch eager, CUDA) on 2 pages per call; context only"}
        except Exception as exc:
            extra["torch_eager_same_gpu"] = {"error": str(exc)[:200]}

    depth_note = (f"{m} micro-batch(es) of {cfg['micro']} pages per step and GPU; {max(m, N_STREAMS)} graph-captured "
                  f"PagePipelines alternate on {N_STREAMS} streams")
    line = {
        "metric": METRIC, "value": res["value"], "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, world),  # the same object in both arms
        "config_detail": dict(anchors_per_page=synth.num_anchors(cfg["imgsz"], cfg["strides"]),
                              decodes=("fused: one kernel reads the head maps once and emits y + the training-side boxes"
                                       if fuse else "separate decode_train (b,A,64) and decode_infer (NCHW maps) kernels"),
                              nms_filter=("inside the inference decode kernel (no second pass over y); the kernel's roofline "
                                          "figure counts the decode's algorithmic bytes only" if fuse_filter else
                                          "separate nms_filter_kernel pass over y"),
                              step="PagePipeline: one CUDA-graph replay of decode_train+assign+decode_infer+NMS per micro-batch; " + depth_note,
                              timing=f"median of {N_WINDOWS} windows of exactly {args.steps} steps; ranks aligned by a device-side "
                                     "all-reduce before each window; CUDA events per rank, max over ranks",
                              ms_per_step_single_stream=res["ms_per_step_single_stream"], numa_node_rank0=numa_node),
        "clocks": clk.summary(),
        "e2e": res["e2e"],
        "gpu_launches": res["launches_per_step"] * args.steps,
        "windows_ms_per_step": res["windows_ms_per_step"],
        "per_rank_ms_per_step": res["per_rank_ms_per_step"],
        "counter_all_reduce_us": res["counter_all_reduce_us"],
        "parity_checked": res["parity_checked"],
        "roofline": res["roofline"],
        "counters": dict(res["counters"], nms_candidates_micro0=res["nms_candidates_micro0"]),
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
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
