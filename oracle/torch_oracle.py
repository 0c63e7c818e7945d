"""TEST INFRASTRUCTURE ONLY - torch restatement of the reference's dense-detection hot path.

This is the parity oracle (see ``oracle/__init__.py``): an independent fp32 restatement, in
plain torch ops, of what the reference computes on this path. It runs on CPU and - because it
issues the same aten arithmetic in the same order as the reference (one rounding per
``+ - * /``, ``torch.topk`` / ``argmax`` for the selections, ``torchvision.ops.nms`` for the
suppression) - on CUDA it is the "reference's own torch implementation on the same inputs"
that BASELINE.json's north_star asks the CUDA path to match.

Pinned by: ``tests/golden/*.npz`` (outputs of the real reference, generated in the build
container by ``tests/golden/make_golden.py``) and, when ``/root/reference`` exists, by a live
comparison in ``tests/test_oracle_vs_reference.py``.

All file:line citations are relative to ``/root/reference/yolov12/ultralytics/``.
"""
from __future__ import annotations

import math

import torch

# --------------------------------------------------------------------------------------
# anchors  (utils/tal.py:333-345)
# --------------------------------------------------------------------------------------


def make_anchors(level_hw, strides, offset=0.5, dtype=torch.float32, device="cpu"):
    """Grid-cell centres in grid units and the per-anchor stride.

    Follows ``make_anchors`` (utils/tal.py:333-345): levels in the given order, row-major
    (y outer, x inner) inside a level, centre = index + ``offset``.
    Returns ``anchor_points (A,2)`` as (x, y) and ``stride_tensor (A,1)``.
    """
    pts, strs = [], []
    for (h, w), s in zip(level_hw, strides):
        xs = torch.arange(w, dtype=dtype, device=device) + offset
        ys = torch.arange(h, dtype=dtype, device=device) + offset
        gx = xs.repeat(h)  # x is the fast index
        gy = ys.repeat_interleave(w)
        pts.append(torch.stack((gx, gy), dim=1))
        strs.append(torch.full((h * w, 1), float(s), dtype=dtype, device=device))
    return torch.cat(pts, 0), torch.cat(strs, 0)


# --------------------------------------------------------------------------------------
# CIoU of (GT, prediction) pairs  (utils/metrics.py:96-130, called as bbox_iou(gt, pd,
# xywh=False, CIoU=True) from utils/tal.py:153-155)
# --------------------------------------------------------------------------------------

_CIOU_EPS = 1e-7


def ciou_gt_pd(gt, pd):
    """Complete-IoU of xyxy box pairs, ``gt`` as box1 and ``pd`` as box2, shape (...,4) -> (...).

    Every arithmetic step is its own torch op so that each rounds to fp32 separately,
    as in the eager reference (utils/metrics.py:103-130).
    """
    e = _CIOU_EPS
    gx1, gy1, gx2, gy2 = gt.unbind(-1)
    px1, py1, px2, py2 = pd.unbind(-1)
    gw = gx2 - gx1
    gh = (gy2 - gy1) + e
    pw = px2 - px1
    ph = (py2 - py1) + e
    iw = (torch.minimum(gx2, px2) - torch.maximum(gx1, px1)).clamp(min=0)
    ih = (torch.minimum(gy2, py2) - torch.maximum(gy1, py1)).clamp(min=0)
    inter = iw * ih
    union = gw * gh + pw * ph - inter + e
    iou = inter / union
    cw = torch.maximum(gx2, px2) - torch.minimum(gx1, px1)
    ch = torch.maximum(gy2, py2) - torch.minimum(gy1, py1)
    c2 = cw * cw + ch * ch + e  # x.pow(2) is x*x in aten
    dx = px1 + px2 - gx1 - gx2
    dy = py1 + py2 - gy1 - gy2
    rho2 = (dx * dx + dy * dy) / 4
    dat = torch.atan(pw / ph) - torch.atan(gw / gh)
    v = (4 / math.pi**2) * (dat * dat)
    alpha = v / (v - iou + (1 + e))
    return iou - (rho2 / c2 + v * alpha)


# --------------------------------------------------------------------------------------
# TaskAlignedAssigner  (utils/tal.py:39-295)
# --------------------------------------------------------------------------------------


def _pow_scalar(x, p):
    # aten pow(tensor, scalar): 0.5 -> sqrt, 2 -> x*x, 1 -> copy, else ::pow  
    return x.pow(p)


def _assign_one_batch(pd_scores, pd_bboxes, anc, gt_labels, gt_bboxes, mask_gt, topk, nc, alpha, beta, eps):
    """Dense (b,G,A) evaluation for a batch slice; see ``tal_assign`` for the contract."""
    b, A, _ = pd_scores.shape
    G = gt_bboxes.shape[1]
    dev = pd_scores.device

    # -- anchors strictly inside a GT (utils/tal.py:241-262): min of the 4 signed distances > 1e-9
    ax = anc[:, 0].view(1, 1, A)
    ay = anc[:, 1].view(1, 1, A)
    x1, y1, x2, y2 = (gt_bboxes[..., k].unsqueeze(-1) for k in range(4))
    dmin = torch.minimum(torch.minimum(ax - x1, ay - y1), torch.minimum(x2 - ax, y2 - ay))
    in_gt = (dmin > 1e-9).to(gt_bboxes.dtype)  # (b,G,A) float 0/1

    # -- per-pair class score, CIoU and alignment metric (utils/tal.py:132-155)
    valid = (in_gt * mask_gt).bool()  # (b,G,A)
    labels_long = gt_labels.squeeze(-1).long()  # (b,G)
    overlaps = torch.zeros((b, G, A), dtype=pd_bboxes.dtype, device=dev)
    cls_score = torch.zeros((b, G, A), dtype=pd_scores.dtype, device=dev)
    bi, gi, ai = valid.nonzero(as_tuple=True)
    if bi.numel():
        cls_score[bi, gi, ai] = pd_scores[bi, ai, labels_long[bi, gi]]
        overlaps[bi, gi, ai] = ciou_gt_pd(gt_bboxes[bi, gi], pd_bboxes[bi, ai]).clamp_(min=0)
    align = _pow_scalar(cls_score, alpha) * _pow_scalar(overlaps, beta)

    # -- top-k anchors per GT (utils/tal.py:157-190)
    k = min(topk, A)
    top_idx = torch.topk(align, k, dim=-1, largest=True).indices  # (b,G,k)
    gt_on = mask_gt.bool().expand(-1, -1, k)
    top_idx = torch.where(gt_on, top_idx, torch.zeros_like(top_idx))
    hits = torch.zeros((b, G, A), dtype=torch.int8, device=dev)
    hits.scatter_add_(-1, top_idx, torch.ones_like(top_idx, dtype=torch.int8))
    hits = torch.where(hits > 1, torch.zeros_like(hits), hits)  # padded GTs hit index 0 k times
    mask_pos = hits.to(align.dtype) * in_gt * mask_gt  # (utils/tal.py:128)

    # -- one GT per anchor (utils/tal.py:264-295)
    n_claims = mask_pos.sum(-2)  # (b,A)
    if n_claims.max() > 1:
        contested = (n_claims.unsqueeze(1) > 1).expand(-1, G, -1)
        best_g = overlaps.argmax(1)  # first max over ALL G rows
        onehot = torch.zeros_like(mask_pos)
        onehot.scatter_(1, best_g.unsqueeze(1), 1)
        mask_pos = torch.where(contested, onehot, mask_pos).float()
        n_claims = mask_pos.sum(-2)
    target_gt_idx = mask_pos.argmax(-2)  # (b,A) first 1, or 0

    # -- targets gathered for EVERY anchor (utils/tal.py:192-238): background rows carry GT 0
    flat = target_gt_idx + torch.arange(b, dtype=torch.int64, device=dev).unsqueeze(-1) * G
    target_labels = gt_labels.long().flatten()[flat].clamp_(min=0)
    target_bboxes = gt_bboxes.reshape(-1, gt_bboxes.shape[-1])[flat]
    one_hot = torch.zeros((b, A, nc), dtype=torch.int64, device=dev)
    one_hot.scatter_(2, target_labels.unsqueeze(-1), 1)
    one_hot = torch.where((n_claims > 0).unsqueeze(-1), one_hot, torch.zeros_like(one_hot))

    # -- normalise by the per-GT maxima (utils/tal.py:111-116)
    align = align * mask_pos
    pos_align = align.amax(dim=-1, keepdim=True)
    pos_ov = (overlaps * mask_pos).amax(dim=-1, keepdim=True)
    norm = (align * pos_ov / (pos_align + eps)).amax(-2).unsqueeze(-1)
    target_scores = one_hot * norm
    return target_labels, target_bboxes, target_scores, n_claims.bool(), target_gt_idx


@torch.no_grad()
def tal_assign(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt,
               topk=10, num_classes=1, alpha=0.5, beta=6.0, eps=1e-9, images_per_chunk=1):
    """Restatement of ``TaskAlignedAssigner.forward`` (utils/tal.py:39-118).

    Inputs (fp32): ``pd_scores (b,A,nc)`` sigmoided, ``pd_bboxes (b,A,4)`` xyxy px,
    ``anc_points (A,2)`` px, ``gt_labels (b,G,1)``, ``gt_bboxes (b,G,4)`` xyxy px,
    ``mask_gt (b,G,1)`` 0/1.
    Returns ``(target_labels int64 (b,A), target_bboxes fp32 (b,A,4), target_scores fp32
    (b,A,nc), fg_mask bool (b,A), target_gt_idx int64 (b,A))``.

    The dense ``(b,G,A)`` tensors are evaluated ``images_per_chunk`` images at a time; images
    are independent, so this does not change any value (only bounds memory).
    """
    b = pd_scores.shape[0]
    G = gt_bboxes.shape[1]
    if G == 0:  # utils/tal.py:64-71 - all float32, labels filled with the background index
        return (
            torch.full_like(pd_scores[..., 0], num_classes),
            torch.zeros_like(pd_bboxes),
            torch.zeros_like(pd_scores),
            torch.zeros_like(pd_scores[..., 0]),
            torch.zeros_like(pd_scores[..., 0]),
        )
    outs = []
    for s in range(0, b, images_per_chunk):
        sl = slice(s, min(b, s + images_per_chunk))
        outs.append(_assign_one_batch(pd_scores[sl], pd_bboxes[sl], anc_points, gt_labels[sl],
                                      gt_bboxes[sl], mask_gt[sl], topk, num_classes, alpha, beta, eps))
    return tuple(torch.cat([o[i] for o in outs], 0) for i in range(5))


# --------------------------------------------------------------------------------------
# decode  (utils/loss.py:197-204; nn/modules/head.py:100-131; nn/modules/block.py:71-74;
#          utils/tal.py:348-357)
# --------------------------------------------------------------------------------------

REG_MAX = 16


def decode_train(pred_dist, anchor_points):
    """Training-side decode: ``pred_dist (b,A,64)`` (side-major, 16 bins per side) and
    ``anchor_points (A,2)`` in grid units -> xyxy ``(b,A,4)`` in grid units.
    Follows ``v8DetectionLoss.bbox_decode`` (utils/loss.py:197-204) + ``dist2bbox(xywh=False)``.
    """
    b, A, c = pred_dist.shape
    bins = torch.arange(c // 4, dtype=pred_dist.dtype, device=pred_dist.device)
    d = pred_dist.view(b, A, 4, c // 4).softmax(3).matmul(bins)  # (b,A,4) = l,t,r,b
    lt, rb = d[..., :2], d[..., 2:]
    return torch.cat((anchor_points - lt, anchor_points + rb), -1)


def decode_infer(levels, strides, nc, reg_max=REG_MAX):
    """Inference decode of per-level head maps ``(b, 4*reg_max+nc, H_l, W_l)`` ->
    ``y (b, 4+nc, A)``: xywh boxes in pixels followed by sigmoided class scores.
    Follows ``Detect._inference`` (nn/modules/head.py:100-131), ``DFL.forward``
    (nn/modules/block.py:71-74; the frozen 1x1 conv with weights arange(16) is written as the
    explicit expectation, so no TF32 conv path is involved) and ``dist2bbox(xywh=True, dim=1)``.
    """
    b = levels[0].shape[0]
    no = 4 * reg_max + nc
    x = torch.cat([l.reshape(b, no, -1) for l in levels], 2)
    anc, st = make_anchors([l.shape[2:] for l in levels], strides, 0.5, x.dtype, x.device)
    anc, st = anc.t(), st.t()  # (2,A), (1,A)
    box, cls = x[:, : 4 * reg_max], x[:, 4 * reg_max:]
    A = x.shape[2]
    p = box.view(b, 4, reg_max, A).softmax(2)
    bins = torch.arange(reg_max, dtype=x.dtype, device=x.device).view(1, 1, reg_max, 1)
    d = (p * bins).sum(2)  # (b,4,A)
    x1y1 = anc.unsqueeze(0) - d[:, :2]
    x2y2 = anc.unsqueeze(0) + d[:, 2:]
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


# --------------------------------------------------------------------------------------
# NMS  (utils/ops.py:167-316, 416-433; torchvision.ops.nms)
# --------------------------------------------------------------------------------------


def greedy_nms(boxes, scores, iou_thres):
    """Plain-loop restatement of ``torchvision.ops.nms`` semantics (pinned torchvision 0.21.0,
    ``uv.lock``; the same arithmetic is restated in-tree by the reference at
    ``examples/YOLOv8-LibTorch-CPP-Inference/main.cc:80-136``): stable sort by score descending,
    suppress j when ``inter / (area_i + area_j - inter) > thr``. Small inputs only (O(N^2) Python).
    """
    order = torch.sort(scores, descending=True, stable=True).indices.tolist()
    bx = boxes.tolist()
    import numpy as np

    f = np.float32
    keep, dead = [], set()
    for ii, i in enumerate(order):
        if i in dead:
            continue
        keep.append(i)
        a = [f(v) for v in bx[i]]
        sa = f(f(a[2] - a[0]) * f(a[3] - a[1]))
        for j in order[ii + 1:]:
            if j in dead:
                continue
            c = [f(v) for v in bx[j]]
            w = max(f(min(a[2], c[2]) - max(a[0], c[0])), f(0))
            h = max(f(min(a[3], c[3]) - max(a[1], c[1])), f(0))
            inter = f(w * h)
            sb = f(f(c[2] - c[0]) * f(c[3] - c[1]))
            den = f(f(sa + sb) - inter)
            with np.errstate(all="ignore"):
                if f(inter / den) > f(iou_thres):
                    dead.add(j)
    return torch.tensor(keep, dtype=torch.int64, device=boxes.device)


def xywh2xyxy(x):
    """utils/ops.py:416-433."""
    half = x[..., 2:] / 2
    return torch.cat((x[..., :2] - half, x[..., :2] + half), -1)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300, nc=0, max_nms=30000, max_wh=7680, in_place=True,
                        nms_fn=None, labels=()):
    """Restatement of ``ops.non_max_suppression`` (utils/ops.py:167-316) for the detection task
    (no masks, no rotated boxes, no wall-clock limit). ``labels``: apriori labels per image,
    ``(n_i, 5) [cls, cx, cy, w, h]`` rows appended with class score 1 (utils/ops.py:256-261).

    ``prediction (b, 4+nc, A)``: xywh px + class scores. Returns a list of ``(n_i, 6)`` tensors
    ``[x1,y1,x2,y2,conf,cls]`` in descending-score order; with ``in_place`` the caller's box rows
    are overwritten with xyxy (utils/ops.py:243-244).
    """
    import torchvision

    nms_fn = nms_fn or torchvision.ops.nms
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    cand = prediction[:, 4:4 + nc].amax(1) > conf_thres
    multi_label = multi_label and nc > 1
    pt = prediction.transpose(-1, -2)  # (b,A,4+nc) view
    if in_place:
        pt[..., :4] = xywh2xyxy(pt[..., :4])
    else:
        pt = torch.cat((xywh2xyxy(pt[..., :4]), pt[..., 4:]), -1)
    cls_filter = None if classes is None else torch.tensor(classes, device=prediction.device)
    out = [torch.zeros((0, 6), device=prediction.device)] * bs
    for i in range(bs):
        x = pt[i][cand[i]]
        if labels is not None and len(labels) and len(labels[i]):  # utils/ops.py:256-261
            lb = labels[i]
            v = torch.zeros((len(lb), nc + 4), device=x.device)
            v[:, :4] = xywh2xyxy(lb[:, 1:5])
            v[range(len(lb)), lb[:, 0].long() + 4] = 1.0
            x = torch.cat((x, v), 0)
        if not x.shape[0]:
            continue
        box, cls = x[:, :4], x[:, 4:4 + nc]
        if multi_label:
            r, c = torch.where(cls > conf_thres)
            x = torch.cat((box[r], x[r, 4 + c, None], c[:, None].float()), 1)
        else:
            conf, j = cls.max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if cls_filter is not None:
            x = x[(x[:, 5:6] == cls_filter).any(1)]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        off = x[:, 5:6] * (0 if agnostic else max_wh)
        keep = nms_fn(x[:, :4] + off, x[:, 4], iou_thres)[:max_det]
        out[i] = x[keep]
    return out


# --------------------------------------------------------------------------------------
# GT packing  (utils/loss.py:180-195)
# --------------------------------------------------------------------------------------


def preprocess(targets, batch_size, scale_tensor):
    """Restatement of ``v8DetectionLoss.preprocess`` (utils/loss.py:180-195): targets (n,6) =
    [image, class, cx, cy, w, h] -> (batch_size, G, 5) = [class, x1, y1, x2, y2] px, zero padded."""
    n, ne = targets.shape
    if n == 0:
        return torch.zeros(batch_size, 0, ne - 1, device=targets.device)
    img = targets[:, 0]
    counts = torch.stack([(img == j).sum() for j in range(batch_size)])
    G = int(img.unique(return_counts=True)[1].max())
    out = torch.zeros(batch_size, G, ne - 1, device=targets.device)
    for j in range(batch_size):
        c = int(counts[j])
        if c:
            out[j, :c] = targets[img == j, 1:]
    out[..., 1:5] = xywh2xyxy(out[..., 1:5].mul_(scale_tensor))
    return out


# --------------------------------------------------------------------------------------
# loss tail  (utils/loss.py:65-113, 243-258)
# --------------------------------------------------------------------------------------


def ciou_generic(box1, box2):
    """bbox_iou(box1, box2, xywh=False, CIoU=True) for autograd use (utils/metrics.py:96-130); the
    weighting factor alpha is computed under no_grad as in the reference."""
    e = _CIOU_EPS
    b1x1, b1y1, b1x2, b1y2 = box1.chunk(4, -1)
    b2x1, b2y1, b2x2, b2y2 = box2.chunk(4, -1)
    w1, h1 = b1x2 - b1x1, b1y2 - b1y1 + e
    w2, h2 = b2x2 - b2x1, b2y2 - b2y1 + e
    inter = (torch.minimum(b1x2, b2x2) - torch.maximum(b1x1, b2x1)).clamp(min=0) * (
        torch.minimum(b1y2, b2y2) - torch.maximum(b1y1, b2y1)).clamp(min=0)
    union = w1 * h1 + w2 * h2 - inter + e
    iou = inter / union
    cw = torch.maximum(b1x2, b2x2) - torch.minimum(b1x1, b2x1)
    ch = torch.maximum(b1y2, b2y2) - torch.minimum(b1y1, b2y1)
    c2 = cw.pow(2) + ch.pow(2) + e
    rho2 = ((b2x1 + b2x2 - b1x1 - b1x2).pow(2) + (b2y1 + b2y2 - b1y1 - b1y2).pow(2)) / 4
    v = (4 / math.pi**2) * ((w2 / h2).atan() - (w1 / h1).atan()).pow(2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + e))
    return iou - (rho2 / c2 + v * alpha)


def loss_tail(pred_scores, pred_distri, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask,
              target_scores_sum=None):
    """[box, cls, dfl] before the hyp gains, restating utils/loss.py:243-254 with BboxLoss.forward
    (:99-113), DFLoss (:72-88) and bbox2dist (utils/tal.py:360-363). ``target_bboxes`` in pixels.
    ``target_scores_sum``: the normaliser when it is not ``target_scores.sum()`` (under AMP the reference sums
    the fp32 scores, utils/loss.py:243, but feeds the BCE their rounded copy, :247)."""
    import torch.nn.functional as F

    reg_max = REG_MAX
    pred_bboxes = decode_train(pred_distri, anchor_points)
    tss = (target_scores.sum() if target_scores_sum is None else target_scores_sum).clamp(min=1)
    cls = F.binary_cross_entropy_with_logits(pred_scores, target_scores, reduction="none").sum() / tss
    box = pred_scores.new_zeros(())
    dfl = pred_scores.new_zeros(())
    if fg_mask.sum():
        tb = target_bboxes / stride_tensor
        weight = target_scores.sum(-1)[fg_mask].unsqueeze(-1)
        iou = ciou_generic(pred_bboxes[fg_mask], tb[fg_mask])
        box = ((1.0 - iou) * weight).sum() / tss
        x1y1, x2y2 = tb.chunk(2, -1)
        ltrb = torch.cat((anchor_points - x1y1, x2y2 - anchor_points), -1).clamp(0, reg_max - 1 - 0.01)
        t = ltrb[fg_mask]
        tl = t.long()
        wl = (tl + 1) - t
        wr = 1 - wl
        logits = pred_distri[fg_mask].view(-1, reg_max)
        ce = (F.cross_entropy(logits, tl.view(-1), reduction="none").view(tl.shape) * wl
              + F.cross_entropy(logits, (tl + 1).view(-1), reduction="none").view(tl.shape) * wr).mean(-1, keepdim=True)
        dfl = (ce * weight).sum() / tss
    return torch.stack((box, cls, dfl))


# --------------------------------------------------------------------------------------
# post-NMS tail  (utils/ops.py:92-127, 319-338)
# --------------------------------------------------------------------------------------


def scale_boxes(img1_shape, boxes, img0_shape):
    """Restatement of ``ops.scale_boxes(img1_shape, boxes, img0_shape)`` with the default
    ``ratio_pad=None, padding=True, xywh=False`` followed by ``clip_boxes`` (in place, returns boxes)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    boxes[..., 0] -= pad_x
    boxes[..., 1] -= pad_y
    boxes[..., 2] -= pad_x
    boxes[..., 3] -= pad_y
    boxes[..., :4] /= gain
    boxes[..., 0] = boxes[..., 0].clamp(0, img0_shape[1])
    boxes[..., 1] = boxes[..., 1].clamp(0, img0_shape[0])
    boxes[..., 2] = boxes[..., 2].clamp(0, img0_shape[1])
    boxes[..., 3] = boxes[..., 3].clamp(0, img0_shape[0])
    return boxes


# --------------------------------------------------------------------------------------
# validator matching  (utils/metrics.py:52-71, models/yolo/detect/val.py:209-228,
# engine/validator.py:222-260 with use_scipy=False)
# --------------------------------------------------------------------------------------


def box_iou(box1, box2, eps=1e-7):
    """Pairwise IoU (N,4) x (M,4) -> (N,M), every op its own fp32 rounding (utils/metrics.py:67-71)."""
    a1, a2 = box1.float().unsqueeze(1).chunk(2, 2)
    b1, b2 = box2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


def match_predictions(pred_classes, true_classes, iou, iouv):
    """(N, n_iou) bool: engine/validator.py:236-260 restated. ``iou`` is (labels, detections). Per threshold:
    matches = pairs of equal class with IoU >= threshold, ordered by IoU descending; each detection keeps its
    first match, then - in ascending detection order - each label keeps its first detection."""
    import numpy as np

    correct = np.zeros((pred_classes.shape[0], len(iouv)), dtype=bool)
    correct_class = true_classes[:, None] == pred_classes
    iou = (iou * correct_class).cpu().numpy()
    for i, threshold in enumerate(torch.as_tensor(iouv).cpu().tolist()):
        lab, det = np.nonzero(iou >= threshold)
        if lab.shape[0] == 0:
            continue
        if lab.shape[0] > 1:
            order = iou[lab, det].argsort()[::-1]                    # IoU descending
            lab, det = lab[order], det[order]
            keep = np.unique(det, return_index=True)[1]              # first (= best) match of every detection
            lab, det = lab[keep], det[keep]                          # now in ascending detection order
            keep = np.unique(lab, return_index=True)[1]              # first detection of every label
            lab, det = lab[keep], det[keep]
        correct[det.astype(int), i] = True
    return torch.tensor(correct, dtype=torch.bool, device=pred_classes.device)


def process_batch(detections, gt_bboxes, gt_cls, iouv):
    """DetectionValidator._process_batch (models/yolo/detect/val.py:226-228)."""
    iou = box_iou(gt_bboxes, detections[:, :4])
    return match_predictions(detections[:, 5], gt_cls, iou, iouv)
