"""Small pass over every libkzd kernel for compute-sanitizer (memcheck / racecheck / synccheck): decodes (all forms, fp32 and
fp16 maps, wide class head), assigner (sparse and dense GT layouts, contested-anchor stress, non-grid anchors, topk 20),
NMS (bucket, cluster and chunk-sort ordering, multi-label, apriori labels, two-phase walk), pipeline graph, loss tail
(dense + sparse, forward + backward), GT packing, scale/clip, validator matching. Sizes are kept small: the tool slows the
kernels 10-100x. usage: compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd  # noqa: E402
from kuzushiji_vision_b200 import loss as kloss  # noqa: E402
from kuzushiji_vision_b200 import ops as kops  # noqa: E402
from kuzushiji_vision_b200 import synth  # noqa: E402

dev = "cuda:0"
done = []


def assign_case(b, G, imgsz, nc, layout, topk=10, jitter=False, dt=None):
    d = synth.assigner_inputs(b, G, imgsz, nc, synth.P2_P5, seed=3, device=dev, layout=layout)
    if jitter:
        d["anc_points"] = d["anc_points"] + torch.rand_like(d["anc_points"])
    ps = d["pd_scores"] if dt is None else d["pd_scores"].to(dt)
    asg = kzd.TaskAlignedAssigner(topk, nc, 0.5, 6.0)
    out = asg(ps, d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    out = asg(ps, d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])  # persistent state reused
    torch.cuda.synchronize()
    done.append(f"assign b={b} G={G} nc={nc} {layout} topk={topk} fg={int(out[3].sum())}")
    return d, out, asg


# ---- decodes ----
for nc, dt in ((1, torch.float32), (3, torch.float16), (40, torch.float32)):
    lv = [x.to(dt) for x in synth.head_outputs(2, nc, 320, synth.P2_P5, seed=1, device=dev, sparse_cls=200)]
    y = kzd.decode_inference(lv, synth.P2_P5, nc)
    y2, xyxy, px = kzd.decode_inference(lv, synth.P2_P5, nc, with_train_boxes=True)
    A = y.shape[2]
    keys = torch.zeros((2, 512), dtype=torch.int64, device=dev)
    rows = torch.zeros((2, 512, 6), device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    y3 = kzd.decode_inference(lv, synth.P2_P5, nc, with_train_boxes=True, nms_pool=(0.25, keys, rows, counts))
    torch.cuda.synchronize()
    done.append(f"decode nc={nc} {dt}")
lv = synth.head_outputs(2, 2, 200, synth.P3_P5, seed=2, device=dev)  # level sizes not divisible by 4: scalar kernel
kzd.decode_inference(lv, synth.P3_P5, 2)
d, _, _ = assign_case(2, 60, 320, 2, "columns")
pd_ = d["pred_distri"].clone().requires_grad_(True)
boxes, px = kzd.bbox_decode(d["anchor_points"], pd_, d["stride_tensor"])
(boxes.sum() + px.sum() * 0.1).backward()
torch.cuda.synchronize()
done.append("decode_train fwd+bwd")

# ---- assigner variants ----
assign_case(2, 400, 320, 2, "uniform")             # many contested anchors: strip index in the patch kernel
assign_case(3, 45, 320, 2, "uniform", topk=20)     # a warp per GT
assign_case(1, 7, 256, 1, "columns", topk=1)       # odd GT count, one-to-one
assign_case(2, 40, 256, 2, "uniform", jitter=True)  # anchors off the grid: all-anchors scan
assign_case(2, 60, 320, 3, "uniform", dt=torch.float16)  # native half scores
assign_case(1, 30, 320, 300, "columns")            # wide class head: dense target_scores zero fill

# ---- NMS ----
lv = synth.head_outputs(3, 4, 320, synth.P2_P5, seed=5, device=dev, sparse_cls=600)
y = kzd.decode_inference(lv, synth.P2_P5, 4)
for kw in (dict(), dict(multi_label=True), dict(agnostic=True), dict(classes=[1, 2]), dict(max_det=50)):
    det = kzd.non_max_suppression(y.clone(), 0.25, 0.5, **kw)
rows_, n_keep = kops.nms_batched(y.clone(), 0.25, 0.5, max_det=300, pool=8192, n_cap=4160)   # chunk sort + rank ordering
rows_, n_keep = kops.nms_batched(y.clone(), 0.01, 0.5, max_det=20, pool=4096, n_cap=1024)    # two-phase walk
labels = [torch.tensor([[1.0, 50, 60, 20, 30], [2.0, 150, 160, 25, 35]], device=dev), torch.zeros((0, 5), device=dev),
          torch.tensor([[0.0, 80, 90, 30, 30]], device=dev)]
det = kzd.non_max_suppression(y.clone(), 0.25, 0.5, labels=labels)
torch.cuda.synchronize()
done.append("nms variants")

# ---- pipeline graph ----
b, G, imgsz, nc = 2, 60, 320, 1
d = synth.assigner_inputs(b, G, imgsz, nc, synth.P2_P5, seed=0, device=dev)
pipe = kzd.PagePipeline(b, synth.level_shapes(imgsz, synth.P2_P5), synth.P2_P5, nc, G, device=dev, conf_thres=0.5, iou_thres=0.7,
                        max_det=100, fuse_decodes=True, fuse_filter=True)
pipe.load(levels=d["levels"], pd_scores=d["pd_scores"], gt_labels=d["gt_labels"], gt_bboxes=d["gt_bboxes"], mask_gt=d["mask_gt"])
for _ in range(2):
    pipe.step()
pipe.results()
done.append("pipeline graph")

# ---- loss tail, GT packing, scale/clip, validator matching ----
for nc in (1, 5):
    d, out, asg = assign_case(2, 80, 320, nc, "uniform")
    tl, tb, ts, fg, gi = out
    for sparse in (False, True):
        x = torch.randn_like(d["pd_scores"]).requires_grad_(True)
        p = d["pred_distri"].clone().requires_grad_(True)
        l3 = kloss.loss_tail(x, p, d["anchor_points"], d["stride_tensor"], tb, ts, fg, asg.counters[0].float(),
                             target_labels=tl if sparse else None)
        l3.sum().backward()
    torch.cuda.synchronize()
    done.append(f"loss tail nc={nc}")
t = torch.cat([torch.cat((torch.full((n, 1), float(j)), torch.zeros(n, 1), torch.rand(n, 4) * 0.5 + 0.2), 1)
               for j, n in enumerate((5, 0, 37, 64))]).to(dev)
kloss.preprocess(t, 4, torch.tensor([320.0, 320.0, 320.0, 320.0], device=dev))
detr = torch.rand((3, 50, 6), device=dev) * 300
kzd.scale_boxes_batched(detr, [50, 0, 17], (320, 320), [(500, 400)] * 3)
dets = torch.rand((40, 6), device=dev) * 300
dets[:, 2:4] += dets[:, :2]
dets[:, 5] = torch.randint(0, 3, (40,), device=dev).float()
gtb = torch.rand((15, 4), device=dev) * 300
gtb[:, 2:] += gtb[:, :2]
kops.process_batch(dets, gtb, torch.randint(0, 3, (15,), device=dev).float(), torch.linspace(0.5, 0.95, 10, device=dev))
torch.cuda.synchronize()
done.append("gt pack, scale boxes, matching")
print("sanitize target ok:", "; ".join(done))
