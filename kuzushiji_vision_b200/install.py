"""Monkey-patch an importable ``ultralytics`` (the reference vendors 8.3.63 under ``yolov12/``) so that
``v8DetectionLoss`` / the predictor / the validator use libkzd on CUDA tensors.

The reference has no plugin registry for this path: the boundary is four module
attributes, patched here and restored by ``uninstall()``:

  ultralytics.utils.tal.TaskAlignedAssigner    <- tal.TaskAlignedAssigner      (utils/tal.py:14)
  ultralytics.utils.loss.TaskAlignedAssigner   <- same (bound by ``from .tal import``, utils/loss.py:9)
  ultralytics.utils.loss.v8DetectionLoss.bbox_decode <- head.loss_bbox_decode  (utils/loss.py:197)
  ultralytics.nn.modules.head.Detect._inference      <- head.detect_inference  (nn/modules/head.py:100)
  ultralytics.utils.ops.non_max_suppression    <- ops.non_max_suppression      (utils/ops.py:167)
  ultralytics.utils.loss.v8DetectionLoss.preprocess  <- loss.preprocess         (utils/loss.py:180)
  ultralytics.models.yolo.detect.predict.DetectionPredictor.postprocess <- batched NMS + scale/clip + Results
                                                     (models/yolo/detect/predict.py:23; only with patch_predictor=True)
  ultralytics.models.yolo.detect.val.DetectionValidator._process_batch <- ops.process_batch
                                                     (models/yolo/detect/val.py:209; only with patch_validator=True)

Calls that are outside this path's scope (CPU tensors, rotated boxes, masks, export formats) are
handed to the reference's own, unpatched function - they never reach libkzd and libkzd never
computes on the CPU.
"""
from __future__ import annotations

import functools
import importlib

from . import head as _head
from . import loss as _loss
from . import ops as _ops
from . import tal as _tal

_saved = {}


def _is_cuda(x):
    if isinstance(x, (list, tuple)):
        x = x[0]
    return bool(getattr(x, "is_cuda", False))


def detect_in_scope(module, detect_cls):
    """True when ``module``'s ``_inference`` is the plain ``Detect`` decode that libkzd implements
    (nn/modules/head.py:100-131 with ``decode_bboxes`` :146-148 = ``dist2bbox(xywh=True)``, 16 DFL bins).
    ``Detect._inference`` is a class attribute, so every subclass inherits the patch: heads that decode
    differently - ``OBB`` (``decode_bboxes`` overridden with ``dist2rbox`` and ``self.angle``), end-to-end
    heads (``v10Detect``: ``xywh=False``, head.py:148), other ``reg_max``, export formats - keep the
    reference's own function."""
    if getattr(module, "export", False) or getattr(module, "format", None) == "imx":
        return False
    if getattr(module, "end2end", False) or getattr(module, "reg_max", 16) != _head.REG_MAX:
        return False
    if getattr(type(module), "decode_bboxes", None) is not getattr(detect_cls, "decode_bboxes", None):
        return False
    return getattr(module, "no", None) == getattr(module, "nc", 0) + 4 * _head.REG_MAX


def install(patch_decode_train=True, patch_preprocess=True, patch_validator=False, patch_predictor=False):
    """Patch the importable ``ultralytics``; returns the list of patched attribute names."""
    if _saved:
        return sorted(_saved)
    u_tal = importlib.import_module("ultralytics.utils.tal")
    u_loss = importlib.import_module("ultralytics.utils.loss")
    u_ops = importlib.import_module("ultralytics.utils.ops")
    u_head = importlib.import_module("ultralytics.nn.modules.head")

    ref_assigner = u_tal.TaskAlignedAssigner
    ref_nms = u_ops.non_max_suppression
    ref_inference = u_head.Detect._inference
    ref_bbox_decode = u_loss.v8DetectionLoss.bbox_decode
    ref_preprocess = u_loss.v8DetectionLoss.preprocess

    class TaskAlignedAssigner(_tal.TaskAlignedAssigner):
        """libkzd on CUDA inputs; the reference's own module for anything else."""

        def forward(self, pd_scores, *rest):
            if not _is_cuda(pd_scores):
                ref = ref_assigner(self.topk, self.num_classes, self.alpha, self.beta, self.eps)
                return ref.forward(pd_scores, *rest)
            return super().forward(pd_scores, *rest)

    @functools.wraps(ref_nms)
    def non_max_suppression(prediction, *args, **kwargs):
        pred = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
        if _is_cuda(pred):
            try:
                return _ops.non_max_suppression(prediction, *args, **kwargs)
            except NotImplementedError:
                pass  # rotated / masks / labels / end2end: not this path
        return ref_nms(prediction, *args, **kwargs)

    def _inference(self, x):
        if not (_is_cuda(x) and detect_in_scope(self, u_head.Detect)):
            return ref_inference(self, x)
        return _head.detect_inference(self, x)

    def bbox_decode(self, anchor_points, pred_dist):
        if not _is_cuda(pred_dist):
            return ref_bbox_decode(self, anchor_points, pred_dist)
        return _head.loss_bbox_decode(self, anchor_points, pred_dist)

    def preprocess(self, targets, batch_size, scale_tensor):
        if not _is_cuda(targets):
            return ref_preprocess(self, targets, batch_size, scale_tensor)
        return _loss.preprocess(targets, batch_size, scale_tensor)

    _saved.update({
        "tal.TaskAlignedAssigner": (u_tal, "TaskAlignedAssigner", ref_assigner),
        "loss.TaskAlignedAssigner": (u_loss, "TaskAlignedAssigner", u_loss.TaskAlignedAssigner),
        "ops.non_max_suppression": (u_ops, "non_max_suppression", ref_nms),
        "head.Detect._inference": (u_head.Detect, "_inference", ref_inference),
    })
    u_tal.TaskAlignedAssigner = TaskAlignedAssigner
    u_loss.TaskAlignedAssigner = TaskAlignedAssigner
    u_ops.non_max_suppression = non_max_suppression
    u_head.Detect._inference = _inference
    if patch_decode_train:
        _saved["loss.v8DetectionLoss.bbox_decode"] = (u_loss.v8DetectionLoss, "bbox_decode", ref_bbox_decode)
        u_loss.v8DetectionLoss.bbox_decode = bbox_decode
    if patch_preprocess:
        _saved["loss.v8DetectionLoss.preprocess"] = (u_loss.v8DetectionLoss, "preprocess", ref_preprocess)
        u_loss.v8DetectionLoss.preprocess = preprocess
    if patch_predictor:
        try:
            u_pred = importlib.import_module("ultralytics.models.yolo.detect.predict")
        except Exception:
            u_pred = None
        if u_pred is not None:
            ref_postprocess = u_pred.DetectionPredictor.postprocess

            def postprocess(self, preds, img, orig_imgs):
                """DetectionPredictor.postprocess (models/yolo/detect/predict.py:23-42): NMS of the whole batch in
                libkzd, scale_boxes + clip_boxes of all kept rows in one kernel, then the reference's own Results
                objects (engine/results.py) around the per-image row slices."""
                import torch

                pred = preds[0] if isinstance(preds, (list, tuple)) else preds
                a = self.args
                plain = (_is_cuda(pred) and pred.dim() == 3 and pred.shape[-1] != 6 and _is_cuda(img)
                         and pred.shape[1] - 4 == len(self.model.names))
                if not plain:
                    return ref_postprocess(self, preds, img, orig_imgs)
                direct = pred.dtype == torch.float32 and pred.is_contiguous()
                work = pred if direct else pred.to(torch.float32).contiguous()
                rows, n_keep = _ops.nms_batched(work, a.conf, a.iou, classes=a.classes, agnostic=a.agnostic_nms,
                                                max_det=a.max_det)
                if not direct:
                    pred[:, :4] = work[:, :4].to(pred.dtype)  # the reference's in-place xyxy side effect
                if not isinstance(orig_imgs, list):
                    orig_imgs = u_ops.convert_torch2numpy_batch(orig_imgs)
                _ops.scale_boxes_batched(rows, n_keep, tuple(img.shape[2:]), [o.shape for o in orig_imgs])
                empty = torch.zeros((0, 6), device=pred.device)
                return [u_pred.Results(o, path=pth, names=self.model.names, boxes=rows[i, :n] if n else empty)
                        for i, (n, o, pth) in enumerate(zip(n_keep, orig_imgs, self.batch[0]))]

            _saved["predict.DetectionPredictor.postprocess"] = (u_pred.DetectionPredictor, "postprocess", ref_postprocess)
            u_pred.DetectionPredictor.postprocess = postprocess
    if patch_validator:
        # the validator module pulls in the whole engine (datasets, cv2, ...): optional, skipped when it does not import
        try:
            u_val = importlib.import_module("ultralytics.models.yolo.detect.val")
        except Exception:
            u_val = None
        if u_val is not None:
            ref_process_batch = u_val.DetectionValidator._process_batch

            def _process_batch(self, detections, gt_bboxes, gt_cls):
                if not (_is_cuda(detections) and _is_cuda(gt_bboxes)):
                    return ref_process_batch(self, detections, gt_bboxes, gt_cls)
                return _ops.process_batch(detections, gt_bboxes, gt_cls, self.iouv)

            _saved["val.DetectionValidator._process_batch"] = (u_val.DetectionValidator, "_process_batch", ref_process_batch)
            u_val.DetectionValidator._process_batch = _process_batch
    return sorted(_saved)


def uninstall():
    for owner, name, orig in _saved.values():
        setattr(owner, name, orig)
    _saved.clear()


def installed():
    return bool(_saved)
