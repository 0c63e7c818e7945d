"""Randomised parity sweep against the same-device oracle (not part of the test suite: minutes, not seconds).
Assigner: random batch / GT count / page size / level set / class count / topk / alpha / beta / layout / score dtype;
NMS: random conf / iou / max_det / agnostic / multi_label / class filter. Every mismatch is printed with its configuration.
With KZD_FUZZ_REF=1 the checker is the UNMODIFIED reference package on CUDA (oracle/_ref, staged by build()) instead of the
oracle restatement.
usage: python tools/fuzz_parity.py [n_assign] [n_nms] [seed] [n_decode_loss]"""
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd  # noqa: E402
from kuzushiji_vision_b200 import synth  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

USE_REF = os.environ.get("KZD_FUZZ_REF", "0") == "1"
if USE_REF:
    from oracle import ref_loader  # noqa: E402

    R = ref_loader.load()
dev = "cuda:0"
n_assign = int(sys.argv[1]) if len(sys.argv) > 1 else 150
n_nms = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = random.Random(int(sys.argv[3]) if len(sys.argv) > 3 else 1234)
bad = 0
t0 = time.time()
for it in range(n_assign):
    b = rng.choice([1, 1, 2, 3, 4])
    G = rng.choice([1, 2, 7, 16, 33, 64, 100, 150, 257, 400])
    imgsz = rng.choice([160, 200, 256, 320, 416, 640])
    strides = rng.choice([synth.P2_P5, synth.P3_P5])
    nc = rng.choice([1, 1, 2, 5, 37])
    topk = rng.choice([1, 3, 10, 10, 13, 16, 20, 32])
    alpha, beta = rng.choice([(0.5, 6.0), (0.5, 6.0), (1.0, 6.0), (0.5, 2.0), (2.0, 1.0), (0.5, 3.0)])
    layout = rng.choice(["columns", "uniform"])
    dt = rng.choice([None, None, torch.float16, torch.bfloat16])
    cfg = dict(b=b, G=G, imgsz=imgsz, strides=strides, nc=nc, topk=topk, alpha=alpha, beta=beta, layout=layout, dt=dt, seed=it)
    d = synth.assigner_inputs(b, G, imgsz, nc, strides, seed=1000 + it, device=dev, layout=layout)
    if rng.random() < 0.15:  # a few invalid rows in the middle of the GT list
        k = rng.randrange(G)
        d["gt_bboxes"][:, k] = 0
        d["mask_gt"][:, k] = 0
    ps = d["pd_scores"] if dt is None else d["pd_scores"].to(dt)
    args = (ps, d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    asg = kzd.TaskAlignedAssigner(topk, nc, alpha, beta)
    asg.overlapped = rng.random() < 0.5  # two GTs per warp / a warp per GT
    mine = asg(*args)
    if USE_REF:
        with torch.no_grad():
            ref = R.TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=alpha, beta=beta).to(dev)(*args)
    else:
        ref = O.tal_assign(*args, topk, nc, alpha, beta)
    ok = all(torch.equal(mine[i], ref[i]) for i in (0, 1, 3, 4))
    try:
        torch.testing.assert_close(mine[2], ref[2].float(), rtol=1e-5, atol=1e-9)
    except AssertionError:
        ok = False
    if not ok:
        bad += 1
        print("ASSIGN MISMATCH", cfg, "fg diff", int((mine[3] != ref[3]).sum()), "idx diff", int((mine[4] != ref[4]).sum()), flush=True)
print("checker:", "unmodified reference package (oracle/_ref) on CUDA" if USE_REF else "oracle/torch_oracle on CUDA", flush=True)
print(f"assign: {n_assign} cases, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
bad_nms = 0
for it in range(n_nms):
    b = rng.choice([1, 2, 3])
    nc = rng.choice([1, 1, 3, 8])
    imgsz = rng.choice([256, 320, 640])
    strides = rng.choice([synth.P2_P5, synth.P3_P5])
    n_obj = rng.choice([50, 300, 1500, 5000])
    lv = synth.head_outputs(b, nc, imgsz, strides, seed=5000 + it, device=dev, sparse_cls=min(n_obj, synth.num_anchors(imgsz, strides)))
    y = kzd.decode_inference(lv, strides, nc)
    kw = dict(conf_thres=rng.choice([0.1, 0.25, 0.5]), iou_thres=rng.choice([0.3, 0.45, 0.7]), max_det=rng.choice([10, 100, 300, 1000]),
              agnostic=rng.random() < 0.3, multi_label=rng.random() < 0.3)
    if nc > 1 and rng.random() < 0.3:
        kw["classes"] = [0, nc - 1]
    mine = kzd.non_max_suppression(y.clone(), **kw)
    ref = R.non_max_suppression(y.clone(), **kw) if USE_REF else O.non_max_suppression(y.clone(), **kw)
    ok = len(mine) == len(ref) and all(m.shape == r.shape and torch.equal(m, r) for m, r in zip(mine, ref))
    if not ok:
        bad_nms += 1
        print("NMS MISMATCH", dict(b=b, nc=nc, imgsz=imgsz, n_obj=n_obj, **kw), [tuple(m.shape) for m in mine], [tuple(r.shape) for r in ref], flush=True)
print(f"nms: {n_nms} cases, {bad_nms} mismatches, {time.time() - t0:.0f} s total", flush=True)
# ---- decodes and loss tail (oracle only: floating-point tolerances of the tests) ----
from kuzushiji_vision_b200 import loss as kloss  # noqa: E402

n_dec = int(sys.argv[4]) if len(sys.argv) > 4 else 60
bad_dec = 0
for it in range(n_dec):
    b = rng.choice([1, 2, 3])
    nc = rng.choice([1, 1, 3, 40])
    imgsz = rng.choice([160, 200, 256, 320, 416])
    strides = rng.choice([synth.P2_P5, synth.P3_P5])
    dt = rng.choice([torch.float32, torch.float32, torch.float16, torch.bfloat16])
    lv = [x.to(dt) for x in synth.head_outputs(b, nc, imgsz, strides, seed=9000 + it, device=dev)]
    lv32 = [x.float() for x in lv]
    y = kzd.decode_inference(lv, strides, nc)
    ref = O.decode_infer(lv32, strides, nc)
    _, st = O.make_anchors([x.shape[2:] for x in lv], strides, device=dev)
    e_box = ((y[:, :4] - ref[:, :4]).abs() / torch.maximum(ref[:, :4].abs(), st.view(1, 1, -1))).max().item()
    e_cls = (y[:, 4:] - ref[:, 4:]).abs().max().item()
    d = synth.assigner_inputs(b, rng.choice([5, 40, 120]), imgsz, nc, strides, seed=9500 + it, device=dev, layout="uniform")
    pdt = d["pred_distri"].to(dt)
    out = kzd.bbox_decode(d["anchor_points"], pdt)
    rt = O.decode_train(pdt.float(), d["anchor_points"])
    e_tr = ((out - rt).abs() / rt.abs().clamp(min=1.0)).max().item()
    # loss tail (sparse form) on the assigner's targets, forward and both gradients
    logits = torch.cat([x.view(b, 64 + nc, -1) for x in d["levels"]], 2)[:, 64:].permute(0, 2, 1).contiguous()
    tl, tb, ts, fg, gi = kzd.TaskAlignedAssigner(10, nc, 0.5, 6.0)(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"],
                                                                    d["gt_bboxes"], d["mask_gt"])
    x1, p1 = logits.clone().requires_grad_(True), d["pred_distri"].clone().requires_grad_(True)
    x2, p2 = logits.clone().requires_grad_(True), d["pred_distri"].clone().requires_grad_(True)
    mine = kloss.loss_tail(x1, p1, d["anchor_points"], d["stride_tensor"], tb, ts, fg, target_labels=tl)
    refl = O.loss_tail(x2, p2, d["anchor_points"], d["stride_tensor"], tb, ts, fg)
    mine.sum().backward()
    refl.sum().backward()
    e_loss = ((mine - refl).abs() / refl.abs().clamp(min=1e-6)).max().item()
    if p2.grad is None:  # no foreground anchor on this batch: the oracle's loss does not touch pred_distri
        p2.grad = torch.zeros_like(p2)
    gs = max(p2.grad.abs().max().item(), 1e-12)
    e_gx = ((x1.grad - x2.grad).abs() / x2.grad.abs().clamp(min=1e-7)).max().item()
    e_gp = ((p1.grad - p2.grad).abs().max().item()) / gs
    if e_box > 1e-5 or e_cls > 1e-6 or e_tr > 1e-5 or e_loss > 5e-5 or e_gx > 2e-4 or e_gp > 5e-5:
        bad_dec += 1
        print("DECODE/LOSS MISMATCH", dict(b=b, nc=nc, imgsz=imgsz, strides=strides, dt=dt), e_box, e_cls, e_tr, e_loss, e_gx, e_gp, flush=True)
print(f"decode + loss: {n_dec} cases, {bad_dec} out of tolerance, {time.time() - t0:.0f} s total", flush=True)
sys.exit(1 if bad or bad_nms or bad_dec else 0)
