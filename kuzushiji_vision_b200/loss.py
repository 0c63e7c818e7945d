"""Drop-ins for the parts of ``v8DetectionLoss`` next to the hot path.

Reference: ``yolov12/ultralytics/utils/loss.py`` - ``preprocess`` :180-195, the loss tail :243-258
(``BboxLoss`` :91-113, ``DFLoss`` :65-88).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def preprocess(targets, batch_size, scale_tensor):
    """``v8DetectionLoss.preprocess(targets, batch_size, scale_tensor)`` (utils/loss.py:180-195):
    ``targets (n, 6) = [image index, class, cx, cy, w, h]`` -> ``(batch_size, G, 5) = [class, x1, y1, x2,
    y2]`` in pixels, ``G`` = largest per-image count, zero rows as padding. One host read-back (to size
    the output) instead of the reference's one per image."""
    nl, ne = targets.shape
    dev = targets.device
    if nl == 0:
        return torch.zeros(batch_size, 0, ne - 1, device=dev)
    _lib.require_cuda(targets, "targets")
    if ne != 6:
        raise _lib.KzdError(f"targets has {ne} columns, expected 6 (image, class, xywh)")
    t = targets.to(torch.float32).contiguous()
    scale = scale_tensor.to(device=dev, dtype=torch.float32).contiguous()
    meta = torch.empty(2 * batch_size + 1, dtype=torch.int32, device=dev)
    counts, first, flags = meta[:batch_size], meta[batch_size:2 * batch_size], meta[2 * batch_size:]
    lib = _lib.lib
    with torch.cuda.device(dev):
        for attempt in range(2):
            _lib.check(lib.kzd_gt_count(t.data_ptr(), nl, batch_size, counts.data_ptr(), first.data_ptr(),
                                        flags.data_ptr(), _stream(dev)), "kzd_gt_count")
            host = meta.cpu()  # the one synchronisation
            flag = int(host[-1])
            if flag & 2:
                raise _lib.KzdError("preprocess: image index outside [0, batch_size)")
            if not (flag & 1):
                break
            # image indices not grouped: stable sort keeps the order of appearance inside an image
            t = t[torch.sort(t[:, 0], stable=True).indices].contiguous()
        G = int(host[:batch_size].max())
        out = torch.empty((batch_size, G, 5), dtype=torch.float32, device=dev)
        _lib.check(lib.kzd_gt_pack(t.data_ptr(), nl, batch_size, G, scale.data_ptr(), counts.data_ptr(),
                                   first.data_ptr(), out.data_ptr(), _stream(dev)), "kzd_gt_pack")
    return out


def loss_preprocess(self, targets, batch_size, scale_tensor):
    """Method form used to patch ``v8DetectionLoss.preprocess``."""
    return preprocess(targets, batch_size, scale_tensor)


# ------------------------------------------------------------------------------------------------
# loss tail (utils/loss.py:243-258) and the whole criterion on libkzd
# ------------------------------------------------------------------------------------------------

class _LossTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_scores, pred_distri, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask,
                target_scores_sum, target_labels, score_dtype):
        b, A, nc = pred_scores.shape
        dev = pred_scores.device
        f32 = torch.float32
        tl = None if target_labels is None else target_labels.to(torch.int64).contiguous()
        # sparse form: half-precision head outputs (AMP) go to libkzd as they are - no up-cast copies of (b,A,nc) / (b,A,64)
        native = tl is not None
        ps_code = _lib.dtype_code(pred_scores.dtype) if native else 0
        pd_code = _lib.dtype_code(pred_distri.dtype) if native else 0
        ps = pred_scores.contiguous() if ps_code else pred_scores.to(f32).contiguous()
        pd = pred_distri.contiguous() if pd_code else pred_distri.to(f32).contiguous()
        anc = anchor_points.to(f32).contiguous()
        st = stride_tensor.to(f32).reshape(-1).contiguous()
        tb = target_bboxes.to(f32).contiguous()
        ts = target_scores.to(f32).contiguous()
        fg = fg_mask.contiguous().view(torch.uint8)
        tss = target_scores_sum.to(f32).reshape(1).contiguous()
        sums = torch.empty(3, dtype=torch.float64, device=dev)
        loss3 = torch.empty(3, dtype=f32, device=dev)
        with torch.cuda.device(dev):
            if tl is None:
                _lib.check(_lib.lib.kzd_loss_tail_fwd(ps.data_ptr(), pd.data_ptr(), anc.data_ptr(), st.data_ptr(),
                                                      tb.data_ptr(), ts.data_ptr(), fg.data_ptr(), tss.data_ptr(), b, A, nc,
                                                      sums.data_ptr(), loss3.data_ptr(), _stream(dev)), "kzd_loss_tail_fwd")
            else:
                _lib.check(_lib.lib.kzd_loss_tail_sparse_fwd(ps.data_ptr(), pd.data_ptr(), anc.data_ptr(), st.data_ptr(),
                                                             tb.data_ptr(), ts.data_ptr(), tl.data_ptr(), fg.data_ptr(),
                                                             tss.data_ptr(), b, A, nc, int(score_dtype), ps_code, pd_code,
                                                             sums.data_ptr(), loss3.data_ptr(), _stream(dev)),
                           "kzd_loss_tail_sparse_fwd")
        ctx.save_for_backward(ps, pd, anc, st, tb, ts, fg, tss, *(() if tl is None else (tl,)))
        ctx.dtypes = (pred_scores.dtype, pred_distri.dtype)
        ctx.score_dtype = int(score_dtype)
        ctx.codes = (ps_code, pd_code)
        return loss3

    @staticmethod
    def backward(ctx, grad_loss3):
        ps, pd, anc, st, tb, ts, fg, tss, *rest = ctx.saved_tensors
        b, A, nc = ps.shape
        dev = ps.device
        g = grad_loss3.to(torch.float32).contiguous()
        g_ps, g_pd = torch.empty_like(ps), torch.empty_like(pd)
        with torch.cuda.device(dev):
            if not rest:
                _lib.check(_lib.lib.kzd_loss_tail_bwd(ps.data_ptr(), pd.data_ptr(), anc.data_ptr(), st.data_ptr(),
                                                      tb.data_ptr(), ts.data_ptr(), fg.data_ptr(), tss.data_ptr(), b, A, nc,
                                                      g.data_ptr(), g_ps.data_ptr(), g_pd.data_ptr(), _stream(dev)),
                           "kzd_loss_tail_bwd")
            else:
                _lib.check(_lib.lib.kzd_loss_tail_sparse_bwd(ps.data_ptr(), pd.data_ptr(), anc.data_ptr(), st.data_ptr(),
                                                             tb.data_ptr(), ts.data_ptr(), rest[0].data_ptr(), fg.data_ptr(),
                                                             tss.data_ptr(), b, A, nc, ctx.score_dtype, ctx.codes[0],
                                                             ctx.codes[1], g.data_ptr(), g_ps.data_ptr(), g_pd.data_ptr(),
                                                             _stream(dev)),
                           "kzd_loss_tail_sparse_bwd")
        return g_ps.to(ctx.dtypes[0]), g_pd.to(ctx.dtypes[1]), None, None, None, None, None, None, None, None


def loss_tail(pred_scores, pred_distri, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask,
              target_scores_sum=None, target_labels=None, score_dtype=None):
    """Fused BCE + CIoU + DFL losses on assigned targets (utils/loss.py:243-254): returns a ``(3,)`` tensor
    ``[box, cls, dfl]`` before the ``hyp`` gains, differentiable wrt ``pred_scores`` (logits, (b,A,nc)) and
    ``pred_distri`` ((b,A,64)). ``target_bboxes`` are the assigner's pixel-space boxes (divided by
    ``stride_tensor`` inside); ``target_scores_sum`` defaults to ``target_scores.sum()`` (a device scalar -
    e.g. ``assigner.counters[0]`` - avoids that extra pass; ``max(., 1)`` is applied in the kernel).

    Sparse-target form: with ``target_labels`` (the assigner's first output) the BCE never reads the dense
    ``(b,A,nc)`` ``target_scores``: ``BCE(x, t) = softplus(x) - t x``, the first term is one streaming pass over
    ``pred_scores`` and the second only exists at the foreground ``(anchor, target_labels[anchor])`` elements.
    It relies on what ``TaskAlignedAssigner`` guarantees (utils/tal.py:228-236): a foreground row of
    ``target_scores`` is one-hot at the anchor's target label, background rows are zero. ``score_dtype``: dtype
    the reference rounds the BCE target to under AMP (``target_scores.to(dtype)``, utils/loss.py:247; the box /
    DFL weight keeps the fp32 scores, :93) - default ``pred_scores.dtype``; pass fp32 ``target_scores``."""
    _lib.require_cuda(pred_scores, "pred_scores")
    if target_scores_sum is None:
        target_scores_sum = target_scores.sum()
    code = _lib.dtype_code(pred_scores.dtype if score_dtype is None else score_dtype)
    return _LossTailFn.apply(pred_scores, pred_distri, anchor_points, stride_tensor, target_bboxes, target_scores,
                             fg_mask, target_scores_sum, target_labels, code)


class DetectionLoss:
    """The reference's detection criterion with every device stage on libkzd: same call signature and
    return value as ``v8DetectionLoss.__call__(preds, batch)`` (utils/loss.py:206-260) -
    ``(loss.sum() * batch_size, loss.detach())`` with ``loss = [box, cls, dfl]`` after the ``hyp`` gains.

    ``stride``: per-level strides, ``nc``: classes, ``hyp``: object with ``box``, ``cls``, ``dfl`` gains
    (``cfg/default.yaml``: 7.5 / 0.5 / 1.5). Only torch glue left: the level ``cat``/``permute`` and sigmoid
    (utils/loss.py:210-215, 235), which belong to the caller side of the three reference signatures.
    """

    def __init__(self, stride, nc, hyp=None, tal_topk=10, reg_max=16):
        from .tal import TaskAlignedAssigner

        if reg_max != 16:
            raise _lib.KzdError("libkzd decodes reg_max = 16 only")
        self.stride = [float(s) for s in stride]
        self.nc, self.reg_max, self.no = int(nc), reg_max, int(nc) + 4 * reg_max
        self.box = getattr(hyp, "box", 7.5)
        self.cls = getattr(hyp, "cls", 0.5)
        self.dfl = getattr(hyp, "dfl", 1.5)
        self.assigner = TaskAlignedAssigner(topk=tal_topk, num_classes=self.nc, alpha=0.5, beta=6.0)
        self._anchors = {}

    def _anchor_cache(self, feats):
        from .tal import make_anchors

        key = (tuple(tuple(f.shape[2:]) for f in feats), str(feats[0].device))
        hit = self._anchors.get(key)
        if hit is None:
            pts, st = make_anchors(feats, self.stride, 0.5)
            hit = (pts.float(), st.float(), (pts * st).float())
            self._anchors = {key: hit}
        return hit

    def __call__(self, preds, batch):
        from .head import bbox_decode

        feats = preds[1] if isinstance(preds, tuple) else preds
        dev = feats[0].device
        bs = feats[0].shape[0]
        # utils/loss.py:210-215 in one pass per tensor: the per-level channel slices are permuted views and the cat
        # writes them straight into the anchor-major layout (the reference's cat -> split -> permute -> contiguous
        # moves the head outputs twice; same values, same autograd graph shape)
        c4 = self.reg_max * 4
        views = [xi.view(bs, self.no, -1) for xi in feats]
        pred_distri = torch.cat([v[:, :c4].permute(0, 2, 1) for v in views], 1)
        pred_scores = torch.cat([v[:, c4:].permute(0, 2, 1) for v in views], 1)
        dtype = pred_scores.dtype
        imgsz = torch.tensor(feats[0].shape[2:], device=dev, dtype=dtype) * self.stride[0]
        anchor_points, stride_tensor, anc_px = self._anchor_cache(feats)

        targets = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1), batch["bboxes"]), 1)
        targets = preprocess(targets.to(dev), bs, imgsz[[1, 0, 1, 0]])
        gt_labels, gt_bboxes = targets.split((1, 4), 2)
        mask_gt = gt_bboxes.sum(2, keepdim=True).gt_(0.0)

        pred_bboxes, pd_bboxes_px = bbox_decode(anchor_points, pred_distri, stride_tensor)
        target_labels, target_bboxes, target_scores, fg_mask, _ = self.assigner(
            pred_scores.detach().sigmoid(), pd_bboxes_px, anc_px, gt_labels, gt_bboxes, mask_gt)
        tss = self.assigner.counters[0]
        # sparse-target BCE: the dense target_scores is never read (only its foreground (anchor, label) elements);
        # under AMP the BCE target is rounded to the logits' dtype inside the kernel, the box / DFL weight is not
        loss3 = loss_tail(pred_scores, pred_distri, anchor_points, stride_tensor, target_bboxes, target_scores,
                          fg_mask.bool() if fg_mask.dtype != torch.bool else fg_mask, tss,
                          target_labels=target_labels.long(), score_dtype=dtype)
        loss = loss3 * torch.tensor([self.box, self.cls, self.dfl], device=dev, dtype=loss3.dtype)
        return loss.sum() * bs, loss.detach()
