"""Drop-in for ``ultralytics.utils.ops.non_max_suppression`` (detection task).

Reference: ``yolov12/ultralytics/utils/ops.py:167-316`` (signature :167-182, checks :217-220, filter
:234-286, class offset :289-295, ``torchvision.ops.nms`` :296, ``[:max_det]`` :297, gather :311) and
``xywh2xyxy`` :416-433. All images of the batch run together in ``libkzd``; the host does one
device->host read of ``2*b`` counters instead of the reference's per-image ``nonzero`` syncs.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

DEFAULT_POOL = 32768   # first-try candidate pool per image
DEFAULT_N_CAP = 4096   # first-try suppression capacity per image (mask = (n_cap/64)^2 * 512 B)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _round_up(v, m):
    return (v + m - 1) // m * m


MASK_BUDGET = 2 << 30  # bytes of IoU bitmask ((n_cap/64)^2 * 512 B per image) one suppression call may hold


def _pack_labels(labels, b, dev):
    """list of per-image ``(n_i, 5)`` ``[cls, cx, cy, w, h]`` tensors -> zero-padded ``(b, L, 5)`` + counts, or None."""
    if labels is None or len(labels) == 0:
        return None
    if len(labels) != b:
        raise _lib.KzdError(f"labels has {len(labels)} entries for a batch of {b} images")
    counts = [int(lb.shape[0]) for lb in labels]
    L = max(counts)
    if L == 0:
        return None
    out = torch.zeros((b, L, 5), dtype=torch.float32, device=dev)
    for i, lb in enumerate(labels):
        if counts[i]:
            out[i, :counts[i]] = lb[:, :5].to(device=dev, dtype=torch.float32)
    return out, torch.tensor(counts, dtype=torch.int32).to(dev), L


def nms_batched(prediction, conf_thres, iou_thres, classes=None, agnostic=False, multi_label=False, max_det=300,
                max_nms=30000, max_wh=7680.0, in_place=True, pool=None, n_cap=None, xyxy_input=False, labels=None):
    """Batched NMS on a contiguous fp32 CUDA tensor ``prediction (b, 4+nc, A)``.

    Returns ``(rows (b, max_det, 6), counts list[int])``; ``rows[i, :counts[i]]`` are image i's kept
    detections ``[x1,y1,x2,y2,conf,cls]`` in descending score order. ``prediction[:, :4]`` is
    overwritten with xyxy when ``in_place`` (the reference's side effect, utils/ops.py:243-244).
    ``xyxy_input=True``: the box rows already hold xyxy (e.g. a tensor a previous call converted in place).
    ``labels``: apriori labels per image (autolabelling, utils/ops.py:256-261): list of ``(n_i, 5)`` tensors
    ``[cls, cx, cy, w, h]`` in pixels, appended to the image's candidates with confidence 1.
    """
    b, ch, A = prediction.shape
    nc = ch - 4
    dev = prediction.device
    multi_label = bool(multi_label) and nc > 1
    packed = _pack_labels(labels, b, dev)
    n_lab = packed[2] if packed else 0
    max_cand = (A * nc if multi_label else A) + n_lab
    if pool is None:
        # Typical prediction settings leave a few thousand candidates per page: a first-try pool of n_cap lets
        # one thread-block cluster per image order them in a single launch (more candidates than n_cap would
        # overflow the suppression stage anyway). Validation settings (conf 0.001, multi_label:
        # models/yolo/detect/val.py:92-102) produce tens of thousands, so they start with the big pool.
        small = max_nms >= DEFAULT_N_CAP and conf_thres >= 0.01 and not multi_label
        pool = (n_cap or DEFAULT_N_CAP) if small else DEFAULT_POOL
    pool = min(max_cand, pool)
    cls_t, n_cls = None, 0
    if classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=dev)
        n_cls = cls_t.numel()
        if n_cls == 0:  # nothing can pass the class filter
            return torch.zeros((b, max_det, 6), device=dev), [0] * b
    out_rows = torch.empty((b, max_det, 6), dtype=torch.float32, device=dev)
    counts = torch.empty((2, b), dtype=torch.int32, device=dev)  # [0] candidates, [1] kept
    lib = _lib.lib
    mode = 2 if xyxy_input else (1 if in_place else 0)
    cap_hint = None  # suppression capacity learned from a failed first try
    with torch.cuda.device(dev):
        while True:
            keys = torch.empty((b, pool), dtype=torch.int64, device=dev)
            rows = torch.empty((b, pool, 6), dtype=torch.float32, device=dev)
            _lib.check(lib.kzd_nms_filter(prediction.data_ptr(), b, A, nc, float(conf_thres), int(multi_label),
                                          cls_t.data_ptr() if n_cls else None, n_cls, mode, keys.data_ptr(),
                                          rows.data_ptr(), counts[0].data_ptr(), pool, _stream(dev)),
                       "kzd_nms_filter")
            if packed:
                _lib.check(lib.kzd_nms_append_labels(packed[0].data_ptr(), packed[1].data_ptr(), b, n_lab, A, nc,
                                                     float(conf_thres), int(multi_label),
                                                     cls_t.data_ptr() if n_cls else None, n_cls, keys.data_ptr(),
                                                     rows.data_ptr(), counts[0].data_ptr(), pool, _stream(dev)),
                           "kzd_nms_append_labels")
            if mode == 1:
                mode = 2  # boxes are xyxy from now on
            cap = min(cap_hint or n_cap or DEFAULT_N_CAP, max_nms, pool)
            cap = _round_up(max(cap, 64), 64)
            while True:
                # the IoU bitmask is (cap/64)^2 * 512 B per image (112 MB at the validator's 30,000 candidates): the
                # batch goes through the suppression stage in groups that keep it under MASK_BUDGET
                per_img = (cap // 64) ** 2 * 512
                group = max(1, min(b, MASK_BUDGET // max(per_img, 1)))
                ws_bytes = lib.kzd_nms_workspace_bytes(group, pool, cap)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                for i0 in range(0, b, group):
                    g = min(group, b - i0)
                    _lib.check(lib.kzd_nms_suppress(keys[i0:].data_ptr(), rows[i0:].data_ptr(), counts[0, i0:].data_ptr(), g,
                                                    pool, cap, float(iou_thres), int(bool(agnostic)), int(max_det),
                                                    int(max_nms), float(max_wh), out_rows[i0:].data_ptr(),
                                                    counts[1, i0:].data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev)),
                               "kzd_nms_suppress")
                host = counts.cpu()  # the one synchronisation of the call
                n_cand, n_keep = host[0].tolist(), host[1].tolist()
                if min(n_keep) >= 0:
                    return out_rows, n_keep
                worst = max(n_cand)
                cap = cap_hint = _round_up(min(worst, max_nms), 64)
                if worst > pool:
                    break  # pool overflow: redo the filter with room for everything
            pool = max_cand


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000,
                        max_wh=7680, in_place=True, rotated=False):
    """Same signature and return convention as the reference (utils/ops.py:167-182): a list of
    ``(n_i, 6)`` tensors ``[x1,y1,x2,y2,conf,cls]`` per image, descending score, at most ``max_det``.

    ``labels``: apriori labels per image (autolabelling, utils/ops.py:256-261).

    Out of scope (raise ``NotImplementedError``; ``install()`` routes those calls to the reference's
    own function): rotated boxes, mask coefficients (``nm > 0``), end-to-end ``(b, n, 6)`` inputs.
    ``max_time_img`` is accepted and ignored: the batch takes well under the reference's
    2 s + 0.05 s/image budget.
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if rotated or prediction.shape[-1] == 6:
        raise NotImplementedError("kuzushiji_vision_b200.non_max_suppression: rotated / end2end inputs")
    _lib.require_cuda(prediction, "prediction")
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    if prediction.shape[1] - nc - 4 != 0:
        raise NotImplementedError("kuzushiji_vision_b200.non_max_suppression: mask coefficients (nm > 0)")
    direct = prediction.dtype == torch.float32 and prediction.is_contiguous()
    work = prediction if direct else prediction.to(torch.float32).contiguous()
    rows, n_keep = nms_batched(work, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, max_nms,
                               float(max_wh), in_place, labels=labels if labels is not None and len(labels) else None)
    if in_place and not direct:
        prediction[:, :4] = work[:, :4].to(prediction.dtype)
    empty = torch.zeros((0, 6), device=prediction.device)
    return [rows[i, :n] if n else empty for i, n in zip(range(bs), n_keep)]


def xywh2xyxy(x):
    """utils/ops.py:416-433 (torch glue; the NMS kernel does this conversion itself)."""
    assert x.shape[-1] == 4, f"input shape last dimension expected 4 but input shape is {x.shape}"
    half = x[..., 2:] / 2
    return torch.cat((x[..., :2] - half, x[..., :2] + half), -1)


def _scale_params(img1_shape, img0_shape, ratio_pad=None):
    """(pad_x, pad_y), 1/gain and the clip limits, from gain and pad exactly as utils/ops.py:110-119 computes them
    (Python floats). ``boxes /= gain`` with a host scalar is, in aten, a multiplication by the reciprocal taken
    in double precision and rounded to fp32 - so that reciprocal is what the kernel gets."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
               round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    return [float(pad[0]), float(pad[1]), 1.0 / float(gain), float(img0_shape[1]), float(img0_shape[0])]


def scale_boxes_batched(rows, n_keep, img1_shape, orig_shapes):
    """``ops.scale_boxes`` + ``clip_boxes`` (utils/ops.py:92-127, 319-338) for the whole NMS output of a
    batch at once, in place: ``rows (b, max_det, 6)`` as returned by ``nms_batched``, ``n_keep`` the kept
    counts (list or None), ``img1_shape`` the network input (h, w), ``orig_shapes`` one (h0, w0[, c]) per
    image - what ``DetectionPredictor.postprocess`` does image by image (detect/predict.py:37-41)."""
    _lib.require_cuda(rows, "rows")
    b, max_det, stride = rows.shape
    dev = rows.device
    params = torch.tensor([_scale_params(img1_shape, s) for s in orig_shapes], dtype=torch.float32).to(dev)
    counts = None if n_keep is None else torch.as_tensor(n_keep, dtype=torch.int32).to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.kzd_scale_boxes(rows.data_ptr(), counts.data_ptr() if counts is not None else None, b,
                                            max_det, stride, params.data_ptr(), _stream(dev)), "kzd_scale_boxes")
    return rows


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """Drop-in for ``ops.scale_boxes`` on one image's ``(n, >=4)`` CUDA boxes (same signature,
    utils/ops.py:92); rescales in place and returns ``boxes``. ``padding=False`` / ``xywh=True`` are not
    on this path."""
    if not padding or xywh:
        raise NotImplementedError("kuzushiji_vision_b200.scale_boxes: padding=False / xywh=True")
    _lib.require_cuda(boxes, "boxes")
    if boxes.numel() == 0:
        return boxes
    if boxes.dtype != torch.float32 or boxes.stride(-1) != 1 or boxes.dim() != 2:
        raise _lib.KzdError("scale_boxes expects a 2-D fp32 tensor with unit stride in the last dimension")
    dev = boxes.device
    params = torch.tensor([_scale_params(img1_shape, img0_shape, ratio_pad)], dtype=torch.float32).to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.kzd_scale_boxes(boxes.data_ptr(), None, 1, boxes.shape[0], boxes.stride(0),
                                            params.data_ptr(), _stream(dev)), "kzd_scale_boxes")
    return boxes


# ------------------------------------------------------------------------------------------------
# validator matching (models/yolo/detect/val.py:209-228, engine/validator.py:222-260)
# ------------------------------------------------------------------------------------------------

def match_predictions_batched(det_rows, n_keep, gt, n_gt, iouv):
    """True-positive flags of a whole batch's kept detections: ``det_rows (b, max_det, >=6)`` as ``nms_batched``
    returns them (descending confidence; already in the labels' coordinate space, e.g. after
    ``scale_boxes_batched``), ``n_keep`` kept counts (list / tensor / None = all rows), ``gt (b, G, 5)``
    ``[cls, x1,y1,x2,y2]`` zero-padded labels with ``n_gt`` labels per image, ``iouv`` the IoU thresholds
    (``torch.linspace(0.5, 0.95, 10)`` in the reference, engine/validator.py:97). Returns bool ``(b, max_det,
    n_iou)`` - row ``d`` of image ``i`` is what ``DetectionValidator._process_batch`` gives for that detection."""
    _lib.require_cuda(det_rows, "det_rows")
    b, max_det, stride = det_rows.shape
    dev = det_rows.device
    det = det_rows if (det_rows.dtype == torch.float32 and det_rows.is_contiguous()) else det_rows.float().contiguous()
    gt = gt.to(device=dev, dtype=torch.float32).contiguous()
    G = gt.shape[1]
    cnt = None if n_keep is None else torch.as_tensor(n_keep, dtype=torch.int32).to(dev)
    gcnt = torch.as_tensor(n_gt, dtype=torch.int32).to(dev)
    thr = torch.as_tensor(iouv, dtype=torch.float32).to(dev).contiguous()
    n_iou = thr.numel()
    correct = torch.empty((b, max_det, n_iou), dtype=torch.uint8, device=dev)
    ws_bytes = _lib.lib.kzd_match_workspace_bytes(b, max_det, G, n_iou)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.kzd_match_predictions(det.data_ptr(), cnt.data_ptr() if cnt is not None else None, b, max_det,
                                                  stride, gt.data_ptr() if G else None, gcnt.data_ptr(), G, thr.data_ptr(),
                                                  n_iou, correct.data_ptr(), ws.data_ptr(), ws_bytes, _stream(dev)),
                   "kzd_match_predictions")
    return correct.view(torch.bool)


def process_batch(detections, gt_bboxes, gt_cls, iouv):
    """Drop-in body for ``DetectionValidator._process_batch(self, detections, gt_bboxes, gt_cls)`` on one image's
    CUDA tensors (models/yolo/detect/val.py:209-228): ``detections (N, 6)``, ``gt_bboxes (M, 4)``, ``gt_cls (M,)``
    -> bool ``(N, len(iouv))``."""
    n, m = detections.shape[0], gt_bboxes.shape[0]
    if n == 0:
        return torch.zeros((0, len(iouv)), dtype=torch.bool, device=detections.device)
    gt = torch.cat((gt_cls.reshape(m, 1).float(), gt_bboxes.reshape(m, 4).float()), 1).unsqueeze(0)
    return match_predictions_batched(detections[:, :6].unsqueeze(0), None, gt, [m], iouv)[0]
