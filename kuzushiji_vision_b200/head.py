"""Drop-in for the Detect-head inference decode and the training-side ``bbox_decode``.

Reference: ``Detect._inference`` nn/modules/head.py:100-131, ``Detect.decode_bboxes`` :146-148,
``DFL.forward`` nn/modules/block.py:71-74, ``dist2bbox`` utils/tal.py:348-357 and
``v8DetectionLoss.bbox_decode`` utils/loss.py:197-204 (all under ``yolov12/ultralytics/``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .tal import make_anchors

REG_MAX = 16


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def decode_inference(levels_x, strides, nc, with_train_boxes=False, nms_pool=None):
    """Fused DFL softmax-expectation + dist2bbox(xywh) + x stride + class sigmoid.

    ``levels_x``: list of per-level maps ``(b, 64+nc, H_l, W_l)`` (CUDA), ``strides``: floats.
    Returns ``y (b, 4+nc, A)`` fp32 - exactly what ``Detect._inference`` returns (head.py:131).

    ``with_train_boxes=True`` returns ``(y, xyxy, xyxy_px)``: the same kernel also emits the training-side
    decode of these maps - ``v8DetectionLoss.bbox_decode`` (utils/loss.py:197-204) in grid units and its
    ``* stride_tensor`` copy (loss.py:236) - so validation-during-training (engine/validator.py:180-185)
    reads the head outputs once and needs no ``(b,A,64)`` transposed copy for the decode.

    ``nms_pool=(conf_thres, keys, rows, counts)`` (the candidate pool ``ops.nms_batched`` allocates: int64
    ``(b,cap)``, fp32 ``(b,cap,6)``, int32 ``(b)``) additionally runs the first stage of
    ``non_max_suppression`` (utils/ops.py:234-286) inside the decode kernel: ``y``'s box rows come back as
    xyxy (what ``in_place=True`` leaves behind, ops.py:243-244) and the pool is filled, so the NMS that follows
    skips its filter pass over ``y``.
    """
    x0 = levels_x[0]
    _lib.require_cuda(x0, "Detect input")
    b = x0.shape[0]
    no = 4 * REG_MAX + nc
    # fp16 / bf16 maps (half inference, AMP) are read natively; anything else is taken as fp32
    code = _lib.dtype_code(x0.dtype)
    in_dtype = x0.dtype if code else torch.float32
    xs = []
    for xi in levels_x:
        if xi.shape[1] != no:
            raise _lib.KzdError(f"level has {xi.shape[1]} channels, expected 64+nc={no}")
        xs.append(xi.to(in_dtype).contiguous())
    lv, A = _lib.levels_array([(xi.shape[2], xi.shape[3], float(s)) for xi, s in zip(xs, strides)])
    ptrs = (C.c_void_p * len(xs))(*[xi.data_ptr() for xi in xs])
    y = torch.empty((b, 4 + nc, A), dtype=torch.float32, device=x0.device)
    if nms_pool is not None:
        conf, keys, rows, counts = nms_pool
        xyxy = torch.empty((b, A, 4), dtype=torch.float32, device=x0.device) if with_train_boxes else None
        px = torch.empty_like(xyxy) if with_train_boxes else None
        with torch.cuda.device(x0.device):
            rc = _lib.lib.kzd_decode_filter(ptrs, code, lv, len(xs), b, nc, y.data_ptr(),
                                            xyxy.data_ptr() if with_train_boxes else None,
                                            px.data_ptr() if with_train_boxes else None, float(conf), keys.data_ptr(),
                                            rows.data_ptr(), counts.data_ptr(), keys.shape[1], _stream(x0.device))
        _lib.check(rc, "kzd_decode_filter")
        return (y, xyxy, px) if with_train_boxes else y
    if with_train_boxes:
        xyxy = torch.empty((b, A, 4), dtype=torch.float32, device=x0.device)
        px = torch.empty_like(xyxy)
        with torch.cuda.device(x0.device):
            rc = _lib.lib.kzd_decode_dual(ptrs, code, lv, len(xs), b, nc, y.data_ptr(), xyxy.data_ptr(), px.data_ptr(),
                                          _stream(x0.device))
        _lib.check(rc, "kzd_decode_dual")
        return y, xyxy, px
    with torch.cuda.device(x0.device):
        rc = _lib.lib.kzd_decode_infer(ptrs, code, lv, len(xs), b, nc, y.data_ptr(), _stream(x0.device))
    _lib.check(rc, "kzd_decode_infer")
    return y


def detect_inference(self, x):
    """Replacement for ``Detect._inference(self, x)`` (nn/modules/head.py:100-131).

    Keeps the module's cached ``anchors`` / ``strides`` / ``shape`` attributes in the state the
    reference leaves them in (head.py:105-107); export-format branches (:109-127) are not handled.
    """
    if self.export:
        raise NotImplementedError("kuzushiji_vision_b200: export-format decode branches are out of scope")
    shape = x[0].shape
    if self.dynamic or self.shape != shape or getattr(self, "_kzd_strides", None) is None:
        self.anchors, self.strides = (t.transpose(0, 1) for t in make_anchors(x, self.stride, 0.5))
        self.shape = shape
        self._kzd_strides = [float(s) for s in self.stride.tolist()]  # one D2H per shape change
    y = decode_inference(x, self._kzd_strides, self.nc)
    return y if x[0].dtype == torch.float32 else y.to(x[0].dtype)


class _BboxDecodeFn(torch.autograd.Function):
    """``softmax(16 bins) . arange(16)`` per side, then ``anchor -/+ distance`` (xyxy, grid units).
    With ``stride_tensor`` the kernel also emits the pixel-space copy ``xyxy * stride`` (non-differentiable,
    the assigner's ``pd_bboxes``, utils/loss.py:236)."""

    @staticmethod
    def forward(ctx, pred_dist, anchor_points, stride_tensor):
        b, A, c = pred_dist.shape
        code = _lib.dtype_code(pred_dist.dtype)  # fp16 / bf16 (AMP) are read natively
        pd = pred_dist.contiguous() if code else pred_dist.to(torch.float32).contiguous()
        anc = anchor_points.to(torch.float32).contiguous()
        out = torch.empty((b, A, 4), dtype=torch.float32, device=pd.device)
        st = px = None
        if stride_tensor is not None:
            st = stride_tensor.to(torch.float32).contiguous()
            px = torch.empty_like(out)
        with torch.cuda.device(pd.device):
            rc = _lib.lib.kzd_decode_train(pd.data_ptr(), code, anc.data_ptr(), st.data_ptr() if st is not None else None,
                                           b, A, out.data_ptr(), px.data_ptr() if px is not None else None,
                                           _stream(pd.device))
        _lib.check(rc, "kzd_decode_train")
        ctx.save_for_backward(pd)
        ctx.in_dtype = pred_dist.dtype
        if px is None:
            return out
        ctx.mark_non_differentiable(px)
        return out, px

    @staticmethod
    def backward(ctx, grad_out, *unused):
        (pd,) = ctx.saved_tensors  # fp32, or the half-precision tensor itself: its gradient is written in the same dtype
        b, A, _ = pd.shape
        g = grad_out.to(torch.float32).contiguous()
        gin = torch.empty_like(pd)
        with torch.cuda.device(pd.device):
            rc = _lib.lib.kzd_decode_train_bwd(pd.data_ptr(), _lib.dtype_code(pd.dtype), g.data_ptr(), b, A, gin.data_ptr(),
                                               _stream(pd.device))
        _lib.check(rc, "kzd_decode_train_bwd")
        return gin.to(ctx.in_dtype), None, None


def bbox_decode(anchor_points, pred_dist, stride_tensor=None):
    """Training-side decode: ``pred_dist (b,A,64)``, ``anchor_points (A,2)`` grid units -> xyxy
    ``(b,A,4)``. Same argument order as ``v8DetectionLoss.bbox_decode`` (utils/loss.py:197);
    differentiable wrt ``pred_dist`` (the box / DFL losses back-propagate through it, loss.py:252).

    Extension: with ``stride_tensor (A,1)`` it returns ``(xyxy, xyxy_px)`` where ``xyxy_px = xyxy *
    stride_tensor`` is produced by the same kernel (the caller glue of utils/loss.py:236)."""
    _lib.require_cuda(pred_dist, "pred_dist")
    if pred_dist.shape[-1] != 4 * REG_MAX:
        raise _lib.KzdError(f"pred_dist has {pred_dist.shape[-1]} channels, expected {4 * REG_MAX}")
    return _BboxDecodeFn.apply(pred_dist, anchor_points, stride_tensor)


def loss_bbox_decode(self, anchor_points, pred_dist):
    """Method form used to patch ``v8DetectionLoss.bbox_decode`` (utils/loss.py:197-204)."""
    if not self.use_dfl:
        from .tal import dist2bbox

        return dist2bbox(pred_dist, anchor_points, xywh=False)
    return bbox_decode(anchor_points, pred_dist)
