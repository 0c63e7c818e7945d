"""kuzushiji_vision_b200 - the dense-detection hot path of Kotomiya07/kuzushiji-vision (its vendored
Ultralytics YOLOv12) as hand-written sm_100a CUDA behind the reference's own Python signatures.

    TaskAlignedAssigner(...).forward(...)   utils/tal.py:39      -> tal.TaskAlignedAssigner
    Detect._inference / DFL decode          nn/modules/head.py:100 -> head.detect_inference, head.decode_inference
    v8DetectionLoss.bbox_decode             utils/loss.py:197    -> head.bbox_decode
    ops.non_max_suppression                 utils/ops.py:167     -> ops.non_max_suppression
    v8DetectionLoss.preprocess              utils/loss.py:180    -> loss.preprocess
    v8DetectionLoss.__call__ (loss tail)    utils/loss.py:206    -> loss.DetectionLoss, loss.loss_tail
    install()                               patches an importable ``ultralytics`` in place
    PagePipeline                            the four stages of one page batch captured as one CUDA graph

The compute lives in ``libkzd.so`` (C ABI: ``include/kzd.h``); importing this package loads it and
raises if it is missing - there is no CPU or torch fallback.
"""
from . import _lib  # noqa: F401  (loads libkzd.so or raises)
from ._lib import KzdError
from .head import bbox_decode, decode_inference, detect_inference
from .install import install, installed, uninstall
from .loss import DetectionLoss, loss_tail, preprocess
from .ops import (match_predictions_batched, nms_batched, non_max_suppression, process_batch, scale_boxes,
                  scale_boxes_batched, xywh2xyxy)
from .pipeline import PagePipeline
from .tal import TaskAlignedAssigner, dist2bbox, infer_levels, make_anchors

__version__ = "0.1.0"
__all__ = [
    "TaskAlignedAssigner", "make_anchors", "dist2bbox", "infer_levels",
    "detect_inference", "decode_inference", "bbox_decode",
    "non_max_suppression", "nms_batched", "xywh2xyxy", "scale_boxes", "scale_boxes_batched",
    "match_predictions_batched", "process_batch",
    "DetectionLoss", "loss_tail", "preprocess", "PagePipeline", "install", "uninstall", "installed", "KzdError",
]
