"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
Inputs come from ``kuzushiji_vision_b200.synth`` (seeded, CPU generator); outputs are whatever the
reference's own functions return on CPU, fp32, torch 2.11.0 / torchvision 0.26.0:
  TaskAlignedAssigner.forward  (yolov12/ultralytics/utils/tal.py:39)
  v8DetectionLoss.bbox_decode  (utils/loss.py:197)  - called as an unbound function on a stub
  Detect._inference            (nn/modules/head.py:100)
  ops.non_max_suppression      (utils/ops.py:167)
  DetectionValidator._process_batch (models/yolo/detect/val.py:209) + BaseValidator.match_predictions
                               (engine/validator.py:222)
Inputs are stored next to the outputs so the fixtures do not depend on the generator staying fixed.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kuzushiji_vision_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def np_(t):
    return t.detach().cpu().numpy()


def assign_case(R, name, b, G, imgsz, nc, strides, layout, seed, topk=10, alpha=0.5, beta=6.0):
    d = synth.assigner_inputs(b, G, imgsz, nc, strides, seed=seed, layout=layout)
    asg = R.TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=alpha, beta=beta)
    tl, tb, ts, fg, gi = asg(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"],
                             d["mask_gt"])
    stub = types.SimpleNamespace(use_dfl=True, proj=torch.arange(16, dtype=torch.float))
    dec_train = R.v8DetectionLoss.bbox_decode(stub, d["anchor_points"], d["pred_distri"])
    det = R.Detect(nc=nc, ch=tuple([64] * len(strides)))
    det.stride = torch.tensor([float(s) for s in strides])
    det.eval()
    y = det._inference([l.clone() for l in d["levels"]])
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        meta=np.array([b, G, imgsz, nc, topk], dtype=np.int64), hyp=np.array([alpha, beta], dtype=np.float64),
        strides=np.array(strides, dtype=np.int64),
        pd_scores=np_(d["pd_scores"]), pd_bboxes=np_(d["pd_bboxes"]), anc_points=np_(d["anc_points"]),
        gt_labels=np_(d["gt_labels"]), gt_bboxes=np_(d["gt_bboxes"]), mask_gt=np_(d["mask_gt"]),
        pred_distri=np_(d["pred_distri"]), anchor_points=np_(d["anchor_points"]),
        **{f"level{i}": np_(l) for i, l in enumerate(d["levels"])},
        target_labels=np_(tl), target_bboxes=np_(tb), target_scores=np_(ts), fg_mask=np_(fg), target_gt_idx=np_(gi),
        decode_train=np_(dec_train), decode_infer=np_(y),
    )
    print(name, "fg", int(fg.sum()), "A", d["anc_points"].shape[0])
    return y


def nms_case(R, name, y, **kw):
    pred = y.clone()
    out = R.non_max_suppression(pred, **kw)
    kw_arr = {k: np.array(v) for k, v in kw.items()}
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), prediction=np_(y), mutated=np_(pred), n_img=np.array(len(out)),
        **{f"kw_{k}": v for k, v in kw_arr.items()}, **{f"out{i}": np_(o) for i, o in enumerate(out)},
    )
    print(name, [tuple(o.shape) for o in out])


def micro_case(R):
    """8x8 grid, stride 8, two overlapping GTs + one padded GT, 12-px predictions."""
    s, n = 8.0, 8
    xs = (torch.arange(n, dtype=torch.float32) + 0.5) * s
    anc = torch.stack((xs.repeat(n), xs.repeat_interleave(n)), 1)
    pd_b = torch.cat((anc - 6.0, anc + 6.0), 1).unsqueeze(0)
    pd_s = torch.full((1, n * n, 2), 0.5)
    gt_b = torch.tensor([[[4.0, 4.0, 36.0, 36.0], [20.0, 20.0, 60.0, 60.0], [0.0, 0.0, 0.0, 0.0]]])
    gt_l = torch.tensor([[[1.0], [0.0], [0.0]]])
    mg = (gt_b.sum(2, keepdim=True) > 0).float()
    out = R.TaskAlignedAssigner(topk=10, num_classes=2, alpha=0.5, beta=6.0)(pd_s, pd_b, anc, gt_l, gt_b, mg)
    np.savez_compressed(
        os.path.join(OUT, "assign_micro.npz"), meta=np.array([1, 3, 64, 2, 10], dtype=np.int64),
        hyp=np.array([0.5, 6.0]), pd_scores=np_(pd_s), pd_bboxes=np_(pd_b), anc_points=np_(anc), gt_labels=np_(gt_l),
        gt_bboxes=np_(gt_b), mask_gt=np_(mg), target_labels=np_(out[0]), target_bboxes=np_(out[1]),
        target_scores=np_(out[2]), fg_mask=np_(out[3]), target_gt_idx=np_(out[4]),
    )
    print("assign_micro fg", int(out[3].sum()), "max score", float(out[2].max()))


EDGE_BOXES = [
    [0.0, 0.0, 320.0, 320.0],        # whole page (tens of thousands of in-GT anchors)
    [-50.0, 40.0, 30.0, 90.0],       # over the left border
    [300.0, 100.0, 400.0, 180.0],    # over the right border
    [100.0, -80.0, 180.0, 20.0],     # over the top
    [60.0, 290.0, 140.0, 500.0],     # over the bottom
    [-200.0, -200.0, -10.0, -10.0],  # off the page
    [400.0, 400.0, 500.0, 500.0],    # off the page
    [50.0, 50.0, 50.0, 120.0],       # zero width
    [64.5, 64.5, 65.5, 65.5],        # thinner than a cell
    [62.0, 62.0, 66.0, 66.0],        # edges exactly on P2 centres (strict test excludes them)
    [10.0, 10.0, 310.0, 30.0],       # wide and flat
    [150.0, 5.0, 170.0, 315.0],      # tall and thin
    [2.0, 2.0, 6.0, 6.0],            # exactly one P2 centre, corner of the page
    [0.0, 0.0, 2.0, 2.0],            # contains no centre
    [314.0, 314.0, 320.0, 320.0],    # last P2 cell
    [-1e4, -1e4, 1e4, 1e4],          # far larger than the page
]


def edge_case(R):
    """GT boxes at the corners of the in-GT geometry (same set as tests/test_gpu_parity.py::
    test_assign_gt_edge_geometry): what the reference's own assigner returns for them."""
    b, G, imgsz, nc, strides = 2, 16, 320, 2, synth.P2_P5
    d = synth.assigner_inputs(b, G, imgsz, nc, strides, seed=93, layout="uniform")
    d["gt_bboxes"][:] = torch.tensor(EDGE_BOXES)
    d["mask_gt"][:] = 1
    d["mask_gt"][1, 3] = 0
    out = R.TaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0)(
        d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    np.savez_compressed(
        os.path.join(OUT, "assign_edge.npz"), meta=np.array([b, G, imgsz, nc, 10], dtype=np.int64),
        hyp=np.array([0.5, 6.0]), pd_scores=np_(d["pd_scores"]), pd_bboxes=np_(d["pd_bboxes"]),
        anc_points=np_(d["anc_points"]), gt_labels=np_(d["gt_labels"]), gt_bboxes=np_(d["gt_bboxes"]),
        mask_gt=np_(d["mask_gt"]), target_labels=np_(out[0]), target_bboxes=np_(out[1]), target_scores=np_(out[2]),
        fg_mask=np_(out[3]), target_gt_idx=np_(out[4]),
    )
    print("assign_edge fg", int(out[3].sum()))


def match_case(R, name="match_val", seed=5):
    """DetectionValidator._process_batch (models/yolo/detect/val.py:209-228) = box_iou + BaseValidator.
    match_predictions (engine/validator.py:222-260) of the unmodified reference, called as unbound functions on a
    stub that carries ``iouv``. Images with jittered copies of the labels as detections (descending confidence),
    duplicates competing for one label, wrong-class detections, images without labels / without detections."""
    import importlib

    val = importlib.import_module("ultralytics.models.yolo.detect.val")
    iouv = torch.linspace(0.5, 0.95, 10)
    stub = types.SimpleNamespace(iouv=iouv)
    stub.match_predictions = types.MethodType(val.DetectionValidator.match_predictions, stub)
    g = torch.Generator().manual_seed(seed)
    out = {"iouv": np_(iouv)}
    cases = [(40, 90, 3), (120, 300, 1), (0, 25, 2), (30, 0, 2), (7, 7, 5), (200, 260, 8)]
    for k, (m, n, nc) in enumerate(cases):
        xy = torch.rand(m, 2, generator=g) * 900 + 20
        wh = torch.rand(m, 2, generator=g) * 80 + 12
        gt = torch.cat((xy, xy + wh), 1)
        gcls = torch.randint(0, nc, (m,), generator=g).float()
        if m and n:
            src = torch.randint(0, m, (n,), generator=g)
            jit = (torch.rand(n, 4, generator=g) - 0.5) * wh[src].repeat(1, 2) * 0.35
            box = gt[src] + jit
            cls = torch.where(torch.rand(n, generator=g) < 0.85, gcls[src], torch.randint(0, nc, (n,), generator=g).float())
        else:
            bx = torch.rand(n, 2, generator=g) * 900
            box = torch.cat((bx, bx + 40), 1)
            cls = torch.zeros(n)
        conf = torch.rand(n, generator=g).sort(descending=True).values
        det = torch.cat((box, conf[:, None], cls[:, None]), 1)
        if n and m:
            correct = val.DetectionValidator._process_batch(stub, det, gt, gcls)
        else:
            correct = torch.zeros((n, 10), dtype=torch.bool)  # update_metrics skips the call (detect/val.py:135-158)
        out[f"det{k}"], out[f"gt{k}"], out[f"gcls{k}"], out[f"correct{k}"] = np_(det), np_(gt), np_(gcls), np_(correct)
        print(name, k, "labels", m, "detections", n, "tp@0.5", int(correct[:, 0].sum()) if n else 0)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)


def main():
    R = ref_loader.load()
    if "--match-only" in sys.argv:
        match_case(R)
        return
    y1 = assign_case(R, "path_p3p5_nc1", 2, 40, 320, 1, synth.P3_P5, "columns", seed=11)
    y2 = assign_case(R, "path_p2p5_nc5", 1, 30, 256, 5, synth.P2_P5, "uniform", seed=12)
    assign_case(R, "path_p2p5_topk13", 1, 24, 256, 3, synth.P2_P5, "columns", seed=13, topk=13, alpha=1.0)
    micro_case(R)
    edge_case(R)
    match_case(R)
    nms_case(R, "nms_nc1", y1, conf_thres=0.5, iou_thres=0.7, max_det=300)
    nms_case(R, "nms_nc5", y2, conf_thres=0.6, iou_thres=0.45, max_det=100)
    nms_case(R, "nms_nc5_agnostic", y2, conf_thres=0.7, iou_thres=0.5, max_det=1000, agnostic=True)
    nms_case(R, "nms_nc5_multilabel", y2, conf_thres=0.8, iou_thres=0.6, max_det=200, multi_label=True)
    nms_case(R, "nms_nc5_classes", y2, conf_thres=0.6, iou_thres=0.6, max_det=200, classes=[1, 3])


if __name__ == "__main__":
    main()
