"""Chained-path report ('stage-wise vs pipeline parity'): run the whole chain with
libkzd (own decode -> own assign / NMS) and with the oracle (oracle decode -> oracle assign / NMS) and count
how many decisions differ because a decoded box or score differs in the last bits. Stage-wise parity
(every stage fed the SAME input as the oracle) is exact and is what tests/ enforce; this report only
quantifies the legitimate downstream effect of ulp-level decode differences.
Writes gpurun_out/chain_flips.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd  # noqa: E402
from kuzushiji_vision_b200 import synth  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

dev = "cuda:0"
b, G, imgsz, nc, strides = 8, 600, 1280, 1, synth.P2_P5
res = {"config": dict(b=b, G=G, imgsz=imgsz, nc=nc)}
d = synth.assigner_inputs(b, G, imgsz, nc, strides, seed=0, device=dev)
# training side
boxes_k = kzd.bbox_decode(d["anchor_points"], d["pred_distri"])
boxes_o = O.decode_train(d["pred_distri"], d["anchor_points"])
res["decode_train_bit_equal_frac"] = float((boxes_k == boxes_o).float().mean())
res["decode_train_max_rel"] = float(((boxes_k - boxes_o).abs() / boxes_o.abs().clamp(min=1)).max())
a_k = kzd.TaskAlignedAssigner(10, nc, 0.5, 6.0)(d["pd_scores"], boxes_k * d["stride_tensor"], d["anc_points"],
                                                 d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
a_o = O.tal_assign(d["pd_scores"], boxes_o * d["stride_tensor"], d["anc_points"], d["gt_labels"], d["gt_bboxes"],
                   d["mask_gt"], 10, nc, 0.5, 6.0)
res["assign_anchors"] = int(a_k[3].numel())
res["assign_fg"] = int(a_o[3].sum())
res["assign_fg_mask_flips"] = int((a_k[3] != a_o[3]).sum())
res["assign_gt_idx_flips"] = int((a_k[4] != a_o[4]).sum())
# inference side
lv = synth.head_outputs(b, nc, imgsz, strides, seed=100, device=dev, sparse_cls=3000)
y_k = kzd.decode_inference(lv, strides, nc)
y_o = O.decode_infer(lv, strides, nc)
res["decode_infer_bit_equal_frac"] = float((y_k == y_o).float().mean())
n_k = kzd.non_max_suppression(y_k.clone(), 0.25, 0.7, max_det=1000)
n_o = O.non_max_suppression(y_o.clone(), 0.25, 0.7, max_det=1000)
res["nms_rows_oracle"] = int(sum(t.shape[0] for t in n_o))
res["nms_images_with_different_count"] = int(sum(a.shape[0] != c.shape[0] for a, c in zip(n_k, n_o)))
same_order = 0
for a, c in zip(n_k, n_o):
    m = min(a.shape[0], c.shape[0])
    same_order += int(((a[:m, :4] - c[:m, :4]).abs().amax(1) < 1e-2).sum())
res["nms_rows_matching_position_and_box"] = same_order
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/chain_flips.json", "w"), indent=1)
print(json.dumps(res, indent=1))
