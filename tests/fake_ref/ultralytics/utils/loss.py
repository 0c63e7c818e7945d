import torch

from oracle import torch_oracle as O

from .tal import TaskAlignedAssigner, make_anchors  # bound at import time, like utils/loss.py:9


class v8DetectionLoss:
    def __init__(self, stride, nc, device):
        self.stride, self.nc, self.no, self.reg_max, self.device = stride, nc, nc + 64, 16, device
        self.use_dfl = True
        self.assigner = TaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0)
        self.proj = torch.arange(16, dtype=torch.float, device=device)

    def preprocess(self, targets, batch_size, scale_tensor):
        return O.preprocess(targets, batch_size, scale_tensor)

    def bbox_decode(self, anchor_points, pred_dist):
        return O.decode_train(pred_dist, anchor_points)

    def __call__(self, preds, batch):
        feats = preds[1] if isinstance(preds, tuple) else preds
        bs = feats[0].shape[0]
        cat = torch.cat([xi.view(bs, self.no, -1) for xi in feats], 2)
        pred_distri, pred_scores = cat.split((64, self.nc), 1)
        pred_scores = pred_scores.permute(0, 2, 1).contiguous()
        pred_distri = pred_distri.permute(0, 2, 1).contiguous()
        imgsz = torch.tensor(feats[0].shape[2:], device=self.device, dtype=pred_scores.dtype) * self.stride[0]
        anchor_points, stride_tensor = make_anchors(feats, self.stride, 0.5)
        targets = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1), batch["bboxes"]), 1)
        targets = self.preprocess(targets.to(self.device), bs, scale_tensor=imgsz[[1, 0, 1, 0]])
        gt_labels, gt_bboxes = targets.split((1, 4), 2)
        mask_gt = gt_bboxes.sum(2, keepdim=True).gt_(0.0)
        pred_bboxes = self.bbox_decode(anchor_points, pred_distri)
        _, target_bboxes, target_scores, fg_mask, _ = self.assigner(
            pred_scores.detach().sigmoid(), (pred_bboxes.detach() * stride_tensor).type(gt_bboxes.dtype),
            anchor_points * stride_tensor, gt_labels, gt_bboxes, mask_gt)
        l3 = O.loss_tail(pred_scores, pred_distri, anchor_points, stride_tensor, target_bboxes, target_scores, fg_mask)
        loss = l3 * torch.tensor([7.5, 0.5, 1.5], device=self.device)
        return loss.sum() * bs, loss.detach(), (target_bboxes, target_scores, fg_mask)
