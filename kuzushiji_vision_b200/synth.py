"""Seeded synthetic kuzushiji-like pages and random-init head outputs.

Used by the tests, ``bench.py`` and ``__graft_entry__.smoke()``. Everything is generated with
``torch.Generator`` on the CPU (so a seed gives the same page on every box) and then moved to the
requested device.
"""
from __future__ import annotations

import torch

REG_MAX = 16
P2_P5 = (4, 8, 16, 32)
P3_P5 = (8, 16, 32)


def level_shapes(imgsz, strides):
    """[(H_l, W_l)] for a square ``imgsz`` page."""
    return [(imgsz // s, imgsz // s) for s in strides]


def num_anchors(imgsz, strides):
    return sum(h * w for h, w in level_shapes(imgsz, strides))


def head_outputs(b, nc, imgsz, strides=P2_P5, seed=0, device="cpu", sparse_cls=None, on_device=False):
    """Per-level raw Detect maps ``(b, 64+nc, H_l, W_l)`` ~ N(0,1), as ``Detect.forward`` emits
    them in training (nn/modules/head.py:69-72 of the reference).

    ``sparse_cls = n`` switches the class logits to the inference-like distribution of
    SURVEY.md section 8d: background logit -6 and, per page, ``n`` uniformly chosen anchors with
    logit N(1,1) on one random class (about 0.98 n candidates above conf 0.25).

    ``on_device=True`` draws the numbers with a generator on ``device`` (seconds faster for the bench's
    hundreds of megabytes per batch; the values then depend on the device type, not only on the seed).
    """
    gdev = device if on_device else "cpu"
    g = torch.Generator(device=gdev).manual_seed(seed)
    no = 4 * REG_MAX + nc
    shapes = level_shapes(imgsz, strides)
    A = sum(h * w for h, w in shapes)
    flat = torch.randn((b, no, A), generator=g, device=gdev)
    if sparse_cls is not None:
        flat[:, 4 * REG_MAX:] = -6.0
        for i in range(b):
            idx = torch.randperm(A, generator=g, device=gdev)[:sparse_cls]
            c = torch.randint(0, nc, (sparse_cls,), generator=g, device=gdev)
            flat[i, 4 * REG_MAX + c, idx] = 1.0 + torch.randn(sparse_cls, generator=g, device=gdev)
    out, o = [], 0
    for h, w in shapes:
        out.append(flat[:, :, o:o + h * w].reshape(b, no, h, w).contiguous().to(device))
        o += h * w
    return out


def gt_pages(b, G, imgsz, nc=1, seed=0, device="cpu", layout="columns", wh_range=(16.0, 56.0), border=60.0):
    """Padded ground truth as ``v8DetectionLoss.preprocess`` produces it (utils/loss.py:180-195):
    ``gt_labels (b,G,1)``, ``gt_bboxes (b,G,4)`` xyxy px with zero rows as padding, and
    ``mask_gt (b,G,1) = gt_bboxes.sum(2) > 0``.

    ``layout="columns"``: characters stacked in vertical columns (pitch about 1.2 x mean width,
    0-15 % vertical overlap), centres at least ``border`` px from the page edge, box sides
    U(wh_range) scaled by imgsz/1280, per-page count U(0.7 G, G).
    ``layout="uniform"``: the stress variant with uniformly random centres (heavy overlap).
    """
    g = torch.Generator().manual_seed(seed + 7919)
    scale = imgsz / 1280.0
    lo, hi = wh_range[0] * scale, wh_range[1] * scale
    brd = border * scale
    boxes = torch.zeros((b, G, 4))
    labels = torch.zeros((b, G, 1))
    for i in range(b):
        n = int(torch.randint(int(0.7 * G), G + 1, (1,), generator=g)) if G > 0 else 0
        w = lo + (hi - lo) * torch.rand(n, generator=g)
        h = lo + (hi - lo) * torch.rand(n, generator=g)
        if layout == "uniform":
            cx = brd + (imgsz - 2 * brd) * torch.rand(n, generator=g)
            cy = brd + (imgsz - 2 * brd) * torch.rand(n, generator=g)
        else:
            pitch = 1.2 * 0.5 * (lo + hi)
            ncol = max(1, int((imgsz - 2 * brd) // pitch))
            cx = torch.empty(n)
            cy = torch.empty(n)
            col_y = torch.full((ncol,), brd)
            col = 0
            for k in range(n):
                # next column when this one is full
                tries = 0
                while col_y[col] + float(h[k]) > imgsz - brd and tries < ncol:
                    col = (col + 1) % ncol
                    tries += 1
                if tries == ncol:  # page full: restart columns (boxes will overlap, as in mosaics)
                    col_y[:] = brd
                cx[k] = imgsz - brd - (col + 0.5) * pitch + float(torch.rand(1, generator=g) - 0.5) * 4 * scale
                cy[k] = col_y[col] + 0.5 * float(h[k])
                col_y[col] += float(h[k]) * (1.0 - 0.15 * float(torch.rand(1, generator=g)))
                if float(torch.rand(1, generator=g)) < 0.03:
                    col = (col + 1) % ncol
        bx = torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), 1)
        perm = torch.randperm(n, generator=g)
        boxes[i, :n] = bx[perm]
        labels[i, :n, 0] = torch.randint(0, nc, (n,), generator=g).float()
    mask = (boxes.sum(2, keepdim=True) > 0).float()
    return labels.to(device), boxes.to(device), mask.to(device)


def assigner_inputs(b, G, imgsz, nc=1, strides=P2_P5, seed=0, device="cpu", layout="columns"):
    """Everything ``TaskAlignedAssigner.forward`` takes, built the way ``v8DetectionLoss.__call__``
    builds it (utils/loss.py:210-241) from random-init head outputs, using plain torch ops
    (this is input preparation, not the path under test).

    Returns a dict with ``pd_scores (b,A,nc)``, ``pd_bboxes (b,A,4)`` px, ``anc_points (A,2)`` px,
    ``gt_labels``, ``gt_bboxes``, ``mask_gt``, plus ``pred_distri (b,A,64)``, ``anchor_points``
    (grid units), ``stride_tensor (A,1)`` and ``levels``.
    """
    levels = head_outputs(b, nc, imgsz, strides, seed, "cpu")
    no = 4 * REG_MAX + nc
    cat = torch.cat([l.view(b, no, -1) for l in levels], 2)
    pred_distri = cat[:, : 4 * REG_MAX].permute(0, 2, 1).contiguous()
    pred_scores = cat[:, 4 * REG_MAX:].permute(0, 2, 1).contiguous()
    pts, strs = [], []
    for (h, w), s in zip(level_shapes(imgsz, strides), strides):
        xs = torch.arange(w, dtype=torch.float32) + 0.5
        ys = torch.arange(h, dtype=torch.float32) + 0.5
        pts.append(torch.stack((xs.repeat(h), ys.repeat_interleave(w)), 1))
        strs.append(torch.full((h * w, 1), float(s)))
    anchor_points, stride_tensor = torch.cat(pts), torch.cat(strs)
    bins = torch.arange(REG_MAX, dtype=torch.float32)
    d = pred_distri.view(b, -1, 4, REG_MAX).softmax(3).matmul(bins)
    xyxy = torch.cat((anchor_points - d[..., :2], anchor_points + d[..., 2:]), -1)
    gl, gb, mg = gt_pages(b, G, imgsz, nc, seed, "cpu", layout)
    out = dict(
        pd_scores=pred_scores.sigmoid(), pd_bboxes=xyxy * stride_tensor, anc_points=anchor_points * stride_tensor,
        gt_labels=gl, gt_bboxes=gb, mask_gt=mg, pred_distri=pred_distri, anchor_points=anchor_points,
        stride_tensor=stride_tensor, levels=levels,
    )
    return {k: ([t.to(device) for t in v] if isinstance(v, list) else v.to(device)) for k, v in out.items()}
