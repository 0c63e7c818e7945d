"""Small, short program for ncu: a few launches of every hot-path kernel on the bench workload."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd  # noqa: E402
from kuzushiji_vision_b200 import synth  # noqa: E402

b, G, imgsz, nc = 16, 600, 1280, 1
dev = "cuda:0"
d = synth.assigner_inputs(b, G, imgsz, nc, synth.P2_P5, seed=0, device=dev)
lv = synth.head_outputs(b, nc, imgsz, synth.P2_P5, seed=100, device=dev, sparse_cls=3000)
asg = kzd.TaskAlignedAssigner(10, nc, 0.5, 6.0)   # lone call: a warp per GT
asg2 = kzd.TaskAlignedAssigner(10, nc, 0.5, 6.0)  # as inside PagePipeline: two GTs per warp
asg2.overlapped = True
A = d["anc_points"].shape[0]
keys = torch.zeros((b, 32768), dtype=torch.int64, device=dev)
rows = torch.zeros((b, 32768, 6), device=dev)
counts = torch.zeros(b, dtype=torch.int32, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    boxes, px = kzd.bbox_decode(d["anchor_points"], d["pred_distri"], d["stride_tensor"])
    out = asg(d["pd_scores"], px, d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    out2 = asg2(d["pd_scores"], px, d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    y = kzd.decode_inference(lv, synth.P2_P5, nc)
    # the bench's dominant kernel: one read of the maps -> y (xyxy) + training-side boxes + NMS candidate pool
    y2, xyxy, px2 = kzd.decode_inference(lv, synth.P2_P5, nc, with_train_boxes=True, nms_pool=(0.25, keys, rows, counts))
    det = kzd.nms_batched(y, 0.25, 0.7, max_det=1000)                        # first-try pool 4096: cluster ordering
    det2 = kzd.nms_batched(y2, 0.25, 0.7, max_det=1000, pool=32768, xyxy_input=True)  # big pool: chunk sort + rank
torch.cuda.synchronize()
assert det[1] == det2[1] and torch.equal(det[0], det2[0])
print("ok", int(out[3].sum()), sum(det[1]))
