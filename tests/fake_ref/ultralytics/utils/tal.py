import torch
import torch.nn as nn

from oracle import torch_oracle as O


class TaskAlignedAssigner(nn.Module):
    def __init__(self, topk=13, num_classes=80, alpha=1.0, beta=6.0, eps=1e-9):
        super().__init__()
        self.topk, self.num_classes, self.bg_idx = topk, num_classes, num_classes
        self.alpha, self.beta, self.eps = alpha, beta, eps

    @torch.no_grad()
    def forward(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt):
        return O.tal_assign(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt, self.topk,
                            self.num_classes, self.alpha, self.beta, self.eps)


def make_anchors(feats, strides, grid_cell_offset=0.5):
    return O.make_anchors([f.shape[2:] for f in feats], strides, grid_cell_offset, feats[0].dtype, feats[0].device)
