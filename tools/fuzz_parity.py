"""Randomised parity sweep against the same-device oracle (not part of the test suite: minutes, not seconds).
Assigner: random batch / GT count / page size / level set / class count / topk / alpha / beta / layout / score dtype;
NMS: random conf / iou / max_det / agnostic / multi_label / class filter. Every mismatch is printed with its configuration.
With KZD_FUZZ_REF=1 the checker is the UNMODIFIED reference package on CUDA (oracle/_ref, staged by build()) instead of the
oracle restatement.
usage: python tools/fuzz_parity.py [n_assign] [n_nms] [seed]"""
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
sys.exit(1 if bad or bad_nms else 0)
