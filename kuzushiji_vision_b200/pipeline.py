"""The whole hot path of one page batch as ONE CUDA graph.

Eagerly, a step is ~15 kernel launches issued from Python through four drop-in calls, plus the host
read-back inside ``non_max_suppression``; on a B200 the kernels take a few hundred microseconds, so
launch latency and the host round trip dominate. ``PagePipeline`` fixes the shapes, captures

    bbox_decode (+ fused x stride)  ->  TaskAlignedAssigner.forward
    Detect._inference decode        ->  NMS filter -> order -> IoU bitmask -> greedy scan

once (``torch.cuda.CUDAGraph`` around the same libkzd calls the drop-ins make) and replays it per batch;
the per-image counters come back through one asynchronous copy into pinned memory. Results are the same
bits as the eager drop-ins (tests/test_gpu_parity.py::test_pipeline_graph_matches_eager).

Reference call sites being replayed: utils/loss.py:229-241 (decode + assign), nn/modules/head.py:73
(inference decode), models/yolo/detect/predict.py:25 (NMS).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .head import REG_MAX, bbox_decode, decode_inference
from .tal import TaskAlignedAssigner, make_anchors


class PagePipeline:
    def __init__(self, batch, level_hw, strides, nc, n_max_boxes, device="cuda:0", topk=10, alpha=0.5, beta=6.0,
                 conf_thres=0.25, iou_thres=0.7, max_det=300, max_nms=30000, agnostic=False, max_wh=7680.0,
                 pool=32768, n_cap=4096, use_graph=True, overlap=True, fuse_decodes=False, fuse_filter=False,
                 cluster_order=False, assign_priority=True, levels_dtype=torch.float32):
        dev = torch.device(device)
        self.device, self.b, self.nc, self.G = dev, int(batch), int(nc), int(n_max_boxes)
        self.level_hw = [(int(h), int(w)) for h, w in level_hw]
        self.strides = [float(s) for s in strides]
        self.A = sum(h * w for h, w in self.level_hw)
        self.conf, self.iou, self.max_det, self.max_nms = float(conf_thres), float(iou_thres), int(max_det), int(max_nms)
        self.agnostic, self.max_wh = bool(agnostic), float(max_wh)
        b, A, G, f32 = self.b, self.A, self.G, torch.float32
        no = 4 * REG_MAX + nc
        # static inputs (fill with load() or write into them directly)
        # levels_dtype fp16 / bf16: the AMP / half-inference flow (engine/trainer.py:379, engine/validator.py:115-117):
        # the decode kernels read the half maps natively and pd_scores keeps the half dtype, so the assigner
        # reproduces the reference's half-precision rounding of pd_scores.pow(alpha) (utils/tal.py:137,150)
        self.levels = [torch.zeros((b, no, h, w), dtype=levels_dtype, device=dev) for h, w in self.level_hw]
        # the fused decode reads the NCHW maps directly: no (b,A,64) transposed copy is needed then
        self.pred_distri = None if fuse_decodes else torch.zeros((b, A, 4 * REG_MAX), dtype=f32, device=dev)
        self.pd_scores = torch.zeros((b, A, nc), dtype=levels_dtype, device=dev)
        self.gt_labels = torch.zeros((b, G, 1), dtype=f32, device=dev)
        self.gt_bboxes = torch.zeros((b, G, 4), dtype=f32, device=dev)
        self.mask_gt = torch.zeros((b, G, 1), dtype=f32, device=dev)
        feats = [torch.empty((1, 1, h, w), device=dev) for h, w in self.level_hw]
        self.anchor_points, self.stride_tensor = make_anchors(feats, self.strides, 0.5)
        self.anc_points = self.anchor_points * self.stride_tensor
        self.assigner = TaskAlignedAssigner(topk, nc, alpha, beta)
        self.assigner.per_stream = False  # one graph = one call in flight: warm-up and capture share the workspace
        self.assigner.overlapped = True   # the NMS chain and other pipelines' decodes run next to the assigner
        # NMS buffers (fixed capacities; an overflow is reported by results())
        self.pool = min(A, int(pool))
        self.n_cap = max(64, (min(int(n_cap), self.max_nms, self.pool) + 63) // 64 * 64)
        if cluster_order and self.max_nms >= self.n_cap:
            # A pool of at most 4096 lets one thread-block cluster per image order the candidates in a single
            # launch (more candidates than n_cap overflow the suppression stage anyway): ~4 us less latency
            # for a single pipeline, but the gang-scheduled clusters cost ~10 us per step when several
            # pipelines are in flight and the decode kernels of the others fill the SMs (measured), so it is
            # opt-in here; the eager drop-in (ops.nms_batched) uses it by default.
            self.pool = min(self.pool, self.n_cap)
        self.keys = torch.empty((b, self.pool), dtype=torch.int64, device=dev)
        self.rows = torch.empty((b, self.pool, 6), dtype=f32, device=dev)
        self.counts = torch.zeros((2, b), dtype=torch.int32, device=dev)
        self.det_rows = torch.zeros((b, self.max_det, 6), dtype=f32, device=dev)
        self.ws_bytes = _lib.lib.kzd_nms_workspace_bytes(b, self.pool, self.n_cap)
        self.nms_ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        # kept-per-image counts are written by the scan kernel straight into pinned host memory (UVA:
        # the pinned pointer is valid on the device), so a step needs no D2H copy node at all
        self.keep_host = torch.full((b,), 0, dtype=torch.int32).pin_memory()
        self.counts_host = torch.zeros((2, b), dtype=torch.int32).pin_memory()
        self.counters_host = torch.zeros(2, dtype=torch.float64).pin_memory()
        self.done = torch.cuda.Event()
        self.overlap = bool(overlap)
        # one kernel reads the head maps once and emits both y and the training-side boxes (the
        # validation-during-training flow, engine/validator.py:180-185); pred_distri is then not read
        self.fuse_decodes = bool(fuse_decodes)
        # the inference decode also does the NMS filter stage (conf threshold, xyxy, candidate pool) on the
        # values it has just computed, so the NMS branch starts at the ordering kernels
        self.fuse_filter = bool(fuse_filter)
        self._side = torch.cuda.Stream(device=dev, priority=-1)  # the small NMS kernels go first
        # the assigner's kernels also go on a high-priority branch: with several pipelines in flight they are
        # then placed ahead of the other pipelines' bulky decode CTAs (measured: -2 % per step pipelined,
        # +3 % for a single pipeline)
        self._side2 = torch.cuda.Stream(device=dev, priority=-1) if (assign_priority and overlap) else None
        self.out = None
        self.graph = None
        # kernels per step: decode(s), assign x3 (candidates, patch, normalise), [nms filter], ordering (one bucket-rank kernel up to
        # n_cap 4096, else chunk sort + rank), (mask, walk) x column ranges - mirrors kzd_nms_suppress
        order = 1 if (self.n_cap <= 4096 and self.pool <= 65535) or (self.pool + 511) // 512 <= 8 else 2
        t0 = (self.max_det + self.max_det // 4 + 127) // 64
        phases = 1 if t0 >= (self.n_cap + 63) // 64 else 2
        self.kernels_per_step = (1 if self.fuse_decodes else 2) + 3 + (0 if self.fuse_filter else 1) + order + 2 * phases
        if use_graph:
            self._capture()

    # -- the step body: the same libkzd calls as the drop-ins ------------------------------------
    def _assign(self, boxes_px):
        with torch.no_grad():
            return self.assigner(self.pd_scores, boxes_px, self.anc_points, self.gt_labels, self.gt_bboxes, self.mask_gt)

    def _nms(self, y):
        dev = self.device
        lib = _lib.lib
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if not self.fuse_filter:
            _lib.check(lib.kzd_nms_filter(y.data_ptr(), self.b, self.A, self.nc, self.conf, 0, None, 0, 1,
                                          self.keys.data_ptr(), self.rows.data_ptr(), self.counts[0].data_ptr(),
                                          self.pool, st), "kzd_nms_filter")
        _lib.check(lib.kzd_nms_suppress(self.keys.data_ptr(), self.rows.data_ptr(), self.counts[0].data_ptr(), self.b,
                                        self.pool, self.n_cap, self.iou, int(self.agnostic), self.max_det,
                                        self.max_nms, self.max_wh, self.det_rows.data_ptr(),
                                        self.keep_host.data_ptr(), self.nms_ws.data_ptr(), self.ws_bytes, st),
                   "kzd_nms_suppress")

    def _pool(self):
        return (self.conf, self.keys, self.rows, self.counts[0]) if self.fuse_filter else None

    # -- the step body: the same libkzd calls as the drop-ins. The training half (decode -> assign) and
    # the inference half (decode -> NMS) do not depend on each other, so they are forked onto two
    # streams (two branches of the captured graph). The two HBM-bound decodes are kept back to back
    # (running them concurrently only splits the bandwidth); the issue/latency-bound NMS chain then
    # overlaps the training decode, and the assigner runs last. With fuse_decodes one kernel produces
    # both decodes from a single read of the maps, then NMS and assign run side by side.
    def _body(self):
        main = torch.cuda.current_stream(self.device)
        side = self._side if self.overlap else main
        if self.fuse_decodes:
            with torch.no_grad():
                y, boxes, boxes_px = decode_inference(self.levels, self.strides, self.nc, with_train_boxes=True,
                                                      nms_pool=self._pool())
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._nms(y)
            if self._side2 is not None:
                self._side2.wait_stream(main)
                with torch.cuda.stream(self._side2):
                    tgt = self._assign(boxes_px)
                main.wait_stream(self._side2)
            else:
                tgt = self._assign(boxes_px)
            main.wait_stream(side)
        else:
            side.wait_stream(main)
            decoded = torch.cuda.Event()
            with torch.cuda.stream(side):
                with torch.no_grad():
                    y = decode_inference(self.levels, self.strides, self.nc, nms_pool=self._pool())
                decoded.record(side)
                self._nms(y)
            main.wait_event(decoded)
            with torch.no_grad():
                boxes, boxes_px = bbox_decode(self.anchor_points, self.pred_distri, self.stride_tensor)
            tgt = self._assign(boxes_px)
            main.wait_stream(side)
        return dict(pred_bboxes=boxes, pd_bboxes=boxes_px, target_labels=tgt[0], target_bboxes=tgt[1],
                    target_scores=tgt[2], fg_mask=tgt[3], target_gt_idx=tgt[4], counters=self.assigner.counters, y=y,
                    det_rows=self.det_rows, counts=self.counts)

    def _capture(self):
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up: level-table inference (one D2H), workspaces, lazy module loads
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._body()

    # -- public API ------------------------------------------------------------------------------
    def load(self, levels=None, pred_distri=None, pd_scores=None, gt_labels=None, gt_bboxes=None, mask_gt=None):
        """Copy new inputs (host or device tensors) into the static buffers, asynchronously."""
        if levels is not None:
            for dst, src in zip(self.levels, levels):
                dst.copy_(src, non_blocking=True)
        for name, src in (("pred_distri", pred_distri), ("pd_scores", pd_scores), ("gt_labels", gt_labels),
                          ("gt_bboxes", gt_bboxes), ("mask_gt", mask_gt)):
            if src is not None and getattr(self, name) is not None:
                getattr(self, name).copy_(src.reshape(getattr(self, name).shape), non_blocking=True)

    def step(self, stream=None):
        """Enqueue one pass over the current inputs (graph replay, or the eager body) on ``stream``
        (default: the current stream). Two pipelines stepped alternately on two streams overlap one
        batch's issue-bound kernels with the other batch's HBM-bound decodes."""
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self.out = self._body()
            self.done.record(stream)
        return self.out

    def results(self):
        """Wait for the last step; returns (outputs dict, kept-per-image list). If the fixed NMS capacities
        of the graph overflowed (more candidates per page than ``pool`` / ``n_cap``), the NMS of this step is
        redone eagerly with capacities sized from the actual counts (the graph itself is left as it is)."""
        self.done.synchronize()
        n_keep = self.keep_host.tolist()
        if min(n_keep) < 0:
            from .ops import nms_batched

            # the graph's filter already rewrote y's box rows as xyxy; capacities from the counts this step measured
            worst = max(self.counts[0].tolist())
            rows, n_keep = nms_batched(self.out["y"], self.conf, self.iou, agnostic=self.agnostic, max_det=self.max_det,
                                       max_nms=self.max_nms, max_wh=self.max_wh, in_place=False, xyxy_input=True,
                                       pool=min(self.A, (worst + 511) // 512 * 512),
                                       n_cap=(min(worst, self.max_nms) + 63) // 64 * 64)
            self.det_rows.copy_(rows)
            self.keep_host.copy_(torch.as_tensor(n_keep, dtype=torch.int32))
        return self.out, n_keep

    def counters(self):
        """[sum(target_scores), count(fg_mask)] of the last step as Python floats (one small D2H)."""
        self.done.synchronize()
        return self.out["counters"].tolist()

    def candidate_counts(self):
        self.done.synchronize()
        return self.counts[0].tolist()

    def detections(self):
        out, n_keep = self.results()
        empty = torch.zeros((0, 6), device=self.device)
        return [out["det_rows"][i, :n] if n else empty for i, n in enumerate(n_keep)]
