"""Drop-in for ``ultralytics.utils.tal`` on the hot path: ``TaskAlignedAssigner``, ``make_anchors``,
``dist2bbox``.

Reference: ``yolov12/ultralytics/utils/tal.py`` - ``TaskAlignedAssigner`` :14-295 (constructor :29-37,
``forward`` :39-80), ``make_anchors`` :333-345, ``dist2bbox`` :348-357. Same names, argument order,
defaults, return dtypes and early-outs; the computation is one call into ``libkzd`` (CUDA, sm_100a).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib

# How torch.topk's zero-metric padding is reproduced (see include/kzd.h, kzd_tal_assign):
# 1 = "lowest anchor index first", which is what torch's CUDA radix-select emits; 0 = ignore padding.
ZERO_FILL_DEFAULT = 1


def infer_levels(anc_points: torch.Tensor):
    """Recover ``[(h, w, stride), ...]`` from anchor centres laid out by ``make_anchors`` (levels
    concatenated, row-major, centre = (index + 0.5) * stride). Returns ``None`` when the points are
    not such a grid. Host-side, run once per anchor count (the device re-verifies on every call).
    """
    pts = anc_points.detach().to("cpu", torch.float32).numpy()
    A = pts.shape[0]
    xs, ys = pts[:, 0], pts[:, 1]
    levels, off = [], 0
    half = np.float32(0.5)
    while off < A:
        s = np.float32(2.0) * xs[off]
        if not (s > 0) or ys[off] != xs[off]:
            return None
        same_row = np.nonzero(ys[off:] != ys[off])[0]
        w = int(same_row[0]) if same_row.size else A - off
        h = 0
        while off + h * w < A:
            k = off + h * w
            if xs[k] != half * s or ys[k] != (np.float32(h) + half) * s or k + w > A:
                break
            h += 1
        if h == 0:
            return None
        ex = np.tile((np.arange(w, dtype=np.float32) + half) * s, h)
        ey = np.repeat((np.arange(h, dtype=np.float32) + half) * s, w)
        if not (np.array_equal(xs[off:off + h * w], ex) and np.array_equal(ys[off:off + h * w], ey)):
            return None
        levels.append((h, w, float(s)))
        off += h * w
        if len(levels) > _lib.KZD_MAX_LEVELS:
            return None
    return levels


class TaskAlignedAssigner(nn.Module):
    """Task-aligned assigner with the reference's interface (utils/tal.py:14-80).

    ``forward(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt)`` returns
    ``(target_labels int64 (b,A), target_bboxes (b,A,4), target_scores (b,A,nc), fg_mask bool (b,A),
    target_gt_idx int64 (b,A))``. After a call ``self.counters`` holds a device ``float64[2]`` =
    ``[target_scores.sum(), fg_mask.sum()]`` (what utils/loss.py:243,250 reduce again).

    Host-side state (none of it changes results):
      * one persistent libkzd workspace per ``(b, A, device, stream)``, grown to the largest ``G`` seen - the
        reference's ``G`` is the batch's largest label count (utils/loss.py:188) and changes every
        iteration. Calls on different streams get different workspaces, so one instance may be used from
        several streams; calls on ONE stream are ordered by the stream.
      * ``overlapped``: a scheduling hint. A lone call (the eager drop-in in a training step) gives every GT a whole
        warp; ``PagePipeline`` sets it, because two GTs per warp leave more of the machine to the kernels running
        next to the assigner. Same results either way.
      * the level table of the anchor grid, inferred on the host from ``anc_points`` once per anchor tensor.
        The same tensor object (same storage, same version counter) is vouched for on later calls, so the
        device-side check is skipped; any other tensor is checked on the device against the cached table and,
        should it not match (same A, different H x W), the next call rebuilds the table.
    """

    def __init__(self, topk=13, num_classes=80, alpha=1.0, beta=6.0, eps=1e-9):
        super().__init__()
        self.topk = topk
        self.num_classes = num_classes
        self.bg_idx = num_classes
        self.alpha = alpha
        self.beta = beta
        self.eps = eps
        self.zero_fill = ZERO_FILL_DEFAULT
        self.stateless = False  # True: libkzd zeroes the per-anchor state on every call (no persistent workspace)
        self.per_stream = True  # False: one workspace per (b, A, device) whatever the stream (PagePipeline: one graph)
        self.overlapped = False  # True: other kernels share the SMs with the call (PagePipeline): two GTs per warp (same results)
        self.counters = None
        self._levels = {}    # (A, device) -> (ctypes level array, n_levels)
        self._vouched = {}   # (A, device) -> (anchor tensor, its version): verified against the table on the host
        self._ws = {}        # (b, A, device, stream) -> [zeroed-once uint8 workspace, G it was sized for]
        self._report = {}    # (A, device) -> (pinned int32[1], event): did the last device check fail?

    def _level_table(self, anc_points):
        """-> (ctypes level array, n_levels, trusted). ``trusted``: this very tensor was checked on the host."""
        key = (anc_points.shape[0], str(anc_points.device))
        hit = self._levels.get(key)
        v = self._vouched.get(key)
        if hit is not None and v is not None and v[0] is anc_points and v[1] == anc_points._version:
            return hit[0], hit[1], True
        rep = self._report.pop(key, None)
        if rep is not None and rep[1].query() and int(rep[0][0]) != 0:
            hit = None  # the cached table did not describe the previous call's anchors: rebuild it
        if hit is None:
            lv = infer_levels(anc_points)  # one D2H of A*8 bytes; checks every anchor against the grid
            arr, _ = _lib.levels_array(lv or [])
            hit = (arr, len(lv) if lv else 0)
            self._levels[key] = hit
            # holding the tensor keeps its storage from being recycled under the same address
            self._vouched[key] = (anc_points, anc_points._version)
            return hit[0], hit[1], hit[1] > 0
        return hit[0], hit[1], False

    def _workspace(self, b, A, G, dev, stream):
        key = (b, A, str(dev), stream if self.per_stream else 0)
        ent = self._ws.get(key)
        if ent is None or ent[1] < G:
            if ent is None and len(self._ws) >= 8:  # train / val batch shapes alternate: keep a few, drop the oldest
                self._ws.pop(next(iter(self._ws)))
            g_cap = G if ent is None else max(G, ent[1] * 2)
            ws = torch.zeros(_lib.lib.kzd_assign_workspace_bytes(b, A, g_cap), dtype=torch.uint8, device=dev)
            ent = [ws, g_cap]
            self._ws[key] = ent
        return ent[0]

    @torch.no_grad()
    def forward(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt):
        self.bs = pd_scores.shape[0]
        self.n_max_boxes = gt_bboxes.shape[1]
        if self.n_max_boxes == 0:  # utils/tal.py:64-71 (all float, labels = background index)
            self.counters = torch.zeros(2, dtype=torch.float64, device=pd_scores.device)
            return (
                torch.full_like(pd_scores[..., 0], self.bg_idx),
                torch.zeros_like(pd_bboxes),
                torch.zeros_like(pd_scores),
                torch.zeros_like(pd_scores[..., 0]),
                torch.zeros_like(pd_scores[..., 0]),
            )
        _lib.require_cuda(pd_scores, "pd_scores")
        if self.topk > _lib.KZD_MAX_TOPK:
            raise _lib.KzdError(f"topk={self.topk} > {_lib.KZD_MAX_TOPK} is not supported by libkzd")
        dev = pd_scores.device
        b, A, nc = pd_scores.shape
        G = self.n_max_boxes
        f32 = torch.float32
        # AMP: the reference keeps bbox_scores in pd_scores' dtype (utils/tal.py:137); libkzd reproduces that one
        # rounding of pd_scores.pow(alpha)
        score_dtype = _lib.dtype_code(pd_scores.dtype)
        native = score_dtype != 0  # fp16 / bf16 scores are read by the kernels as they are: no up-cast pass over (b,A,nc)
        pd_scores_c = pd_scores.contiguous() if native else pd_scores.to(f32).contiguous()
        pd_bboxes_c = pd_bboxes.to(f32).contiguous()
        anc_c = anc_points if (anc_points.dtype == f32 and anc_points.is_contiguous()) else anc_points.to(f32).contiguous()
        gl = gt_labels.reshape(b, G).to(f32).contiguous()
        gb = gt_bboxes.to(f32).contiguous()
        mg = mask_gt.reshape(b, G).to(f32).contiguous()
        levels, n_levels, trusted = self._level_table(anc_c)

        target_labels = torch.empty((b, A), dtype=torch.int64, device=dev)
        target_bboxes = torch.empty((b, A, 4), dtype=f32, device=dev)
        target_scores = torch.empty((b, A, nc), dtype=f32, device=dev)
        fg_u8 = torch.empty((b, A), dtype=torch.uint8, device=dev)
        target_gt_idx = torch.empty((b, A), dtype=torch.int64, device=dev)
        counters = torch.empty(2, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev)
        # persistent, zeroed-once workspace: every call leaves the per-anchor state all-zero again, so no
        # per-call memset is needed (include/kzd.h, KZD_ASSIGN_PERSISTENT)
        if self.stateless:
            ws = torch.empty(_lib.lib.kzd_assign_workspace_bytes(b, A, G), dtype=torch.uint8, device=dev)
        else:
            ws = self._workspace(b, A, G, dev, stream.cuda_stream)
        flags = ((0 if self.stateless else _lib.KZD_ASSIGN_PERSISTENT) | (_lib.KZD_ASSIGN_TRUSTED_ANCHORS if trusted else 0)
                 | (_lib.KZD_ASSIGN_NATIVE_SCORES if native else 0) | (0 if self.overlapped else _lib.KZD_ASSIGN_WARP_PER_GT))
        with torch.cuda.device(dev):
            rc = _lib.lib.kzd_tal_assign(
                pd_scores_c.data_ptr(), pd_bboxes_c.data_ptr(), anc_c.data_ptr(), gl.data_ptr(), gb.data_ptr(),
                mg.data_ptr(), b, A, G, nc, int(self.topk), float(self.alpha), float(self.beta), float(self.eps),
                levels, n_levels, int(self.zero_fill), score_dtype,
                target_labels.data_ptr(), target_bboxes.data_ptr(), target_scores.data_ptr(), fg_u8.data_ptr(),
                target_gt_idx.data_ptr(), counters.data_ptr(), ws.data_ptr(), ws.numel(), flags,
                C.c_void_p(stream.cuda_stream),
            )
        _lib.check(rc, "kzd_tal_assign")
        if not trusted and n_levels > 0 and not torch.cuda.is_current_stream_capturing():
            # asynchronous report of the device-side anchor check, looked at by the next call
            off = _lib.lib.kzd_assign_report_offset(b, A)
            host = torch.zeros(1, dtype=torch.int32).pin_memory()
            host.copy_(ws[off:off + 4].view(torch.int32), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            self._report[(A, str(dev))] = (host, ev)
        self.counters = counters
        return target_labels, target_bboxes, target_scores, fg_u8.view(torch.bool), target_gt_idx


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Anchor centres (grid units) and per-anchor stride; reference utils/tal.py:333-345.

    Host-side torch glue (exact half-integers, A*12 bytes): kept in torch on purpose.
    """
    assert feats is not None
    dtype, device = feats[0].dtype, feats[0].device
    pts, strs = [], []
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2:] if isinstance(feats, list) else (int(feats[i][0]), int(feats[i][1]))
        sx = torch.arange(end=w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(end=h, device=device, dtype=dtype) + grid_cell_offset
        pts.append(torch.stack((sx.repeat(h), sy.repeat_interleave(w)), -1))
        strs.append(torch.full((h * w, 1), float(stride), dtype=dtype, device=device))
    return torch.cat(pts), torch.cat(strs)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """ltrb distances -> xywh / xyxy boxes; reference utils/tal.py:348-357 (plain torch glue; the fused
    CUDA forms are ``head.detect_inference`` and ``loss.bbox_decode``)."""
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)
