"""GPU parity against the UNMODIFIED reference executed on the same CUDA device.

``oracle/make_ref.py`` (run by ``__graft_entry__.build()`` in the build container) stages a byte-identical
copy of the reference's vendored Ultralytics package under the git-ignored ``oracle/_ref/yolov12``; it
travels to the GPU box with the snapshot. These tests import it there (``oracle/ref_loader.py``) and compare
libkzd, through the drop-in signatures / the C ABI, with the reference's own
``TaskAlignedAssigner.forward`` (utils/tal.py:39), ``Detect._inference`` (nn/modules/head.py:100),
``v8DetectionLoss`` (utils/loss.py:155-260) and ``ops.non_max_suppression`` (utils/ops.py:167) on the same
inputs: indices and NMS rows bit-exact, scores / decoded boxes within 1e-5 (BASELINE.json north_star).
Skipped (not failed) when no copy of the reference is importable.
"""
import types

import pytest
import torch

import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference not staged (oracle/_ref)")]
DEV = "cuda:0"
RTOL = 1e-5


@pytest.fixture(scope="module")
def R():
    torch.backends.cudnn.allow_tf32 = False  # the DFL 1x1 conv must not round the softmax to TF32
    torch.backends.cuda.matmul.allow_tf32 = False
    ns = ref_loader.load()
    # the reference's NMS gives up after 2 s + 0.05 s/image (utils/ops.py:238,312-314): pay the one-off
    # torchvision import / kernel-load cost here, not inside a compared call
    warm = torch.rand(1, 6, 64, device=DEV)
    ns.non_max_suppression(warm, 0.1, 0.5)
    torch.cuda.synchronize()
    return ns


def _check_assign_exact(mine, ref):
    names = ("target_labels", "target_bboxes", "target_scores", "fg_mask", "target_gt_idx")
    for n, m, r in zip(names, mine, ref):
        assert m.dtype == r.dtype and m.shape == r.shape, n
    assert torch.equal(mine[3], ref[3]), f"fg_mask differs at {int((mine[3] != ref[3]).sum())} anchors"
    assert torch.equal(mine[4], ref[4]), f"target_gt_idx differs at {int((mine[4] != ref[4]).sum())} anchors"
    assert torch.equal(mine[0], ref[0]) and torch.equal(mine[1], ref[1])
    torch.testing.assert_close(mine[2], ref[2], rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("cfg", [
    # b, G, imgsz, nc, strides, layout, topk, alpha
    (1, 200, 640, 1, synth.P3_P5, "columns", 10, 0.5),    # BASELINE configs[0]
    (2, 150, 640, 5, synth.P2_P5, "uniform", 10, 0.5),    # heavy overlap: contested anchors
    (2, 600, 1280, 1, synth.P2_P5, "columns", 10, 0.5),   # configs[1] pages (two of the sixteen)
    (1, 1000, 1280, 1, synth.P2_P5, "columns", 10, 0.5),  # configs[4] page (~1000 GT)
    (1, 300, 1280, 64, synth.P2_P5, "uniform", 10, 0.5),  # multi-class, class-aware metric
    (3, 64, 320, 2, synth.P2_P5, "uniform", 13, 1.0),     # the class defaults topk=13, alpha=1
    (1, 7, 256, 1, synth.P2_P5, "columns", 1, 0.5),       # one-to-one (topk=1, E2EDetectLoss utils/loss.py:728-743)
])
def test_assign_matches_unmodified_reference_on_cuda(R, cfg):
    b, G, imgsz, nc, strides, layout, topk, alpha = cfg
    d = synth.assigner_inputs(b, G, imgsz, nc, strides, seed=11, device=DEV, layout=layout)
    args = (d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    ref = R.TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=alpha, beta=6.0).to(DEV)(*args)
    mine = kzd.TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=alpha, beta=6.0)(*args)
    torch.cuda.synchronize()
    _check_assign_exact(mine, ref)


def _ref_detect(R, nc, strides):
    det = R.Detect(nc=nc, ch=tuple(64 for _ in strides)).to(DEV).eval()
    det.stride = torch.tensor([float(s) for s in strides], device=DEV)
    return det


@pytest.mark.parametrize("cfg", [(2, 1, 640, synth.P3_P5), (2, 7, 320, synth.P2_P5), (2, 1, 1280, synth.P2_P5)])
def test_detect_inference_matches_unmodified_reference_on_cuda(R, cfg):
    b, nc, imgsz, strides = cfg
    lv = synth.head_outputs(b, nc, imgsz, strides, seed=21, device=DEV)
    det = _ref_detect(R, nc, strides)
    with torch.no_grad():
        ref = det._inference([x.clone() for x in lv])
    y = kzd.decode_inference(lv, strides, nc)
    st = det.strides.reshape(1, 1, -1)
    rel = ((y[:, :4] - ref[:, :4]).abs() / torch.maximum(ref[:, :4].abs(), st.expand_as(ref[:, :4]))).max()
    assert rel < RTOL, float(rel)                       # relative to max(|ref|, stride)
    assert (y[:, 4:] - ref[:, 4:]).abs().max() <= 1e-6  # sigmoid
    # the drop-in method on the reference's own module object keeps its cached state
    det2 = _ref_detect(R, nc, strides)
    y2 = kzd.detect_inference(det2, [x.clone() for x in lv])
    assert torch.equal(y2, y) and det2.shape == lv[0].shape and torch.equal(det2.anchors, det.anchors)


@pytest.mark.parametrize("kw", [
    dict(conf_thres=0.25, iou_thres=0.7, max_det=300),
    dict(conf_thres=0.25, iou_thres=0.45, max_det=1000, agnostic=True),
    dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True),   # validator settings, detect/val.py:92-102
    dict(conf_thres=0.3, iou_thres=0.6, max_det=100, classes=[0, 2]),
])
def test_nms_matches_unmodified_reference_on_cuda(R, kw):
    b, nc, imgsz, strides = 3, 4, 640, synth.P2_P5
    lv = synth.head_outputs(b, nc, imgsz, strides, seed=31, device=DEV, sparse_cls=1500)
    y = kzd.decode_inference(lv, strides, nc)
    y_ref, y_mine = y.clone(), y.clone()
    ref = R.non_max_suppression(y_ref, **kw)
    mine = kzd.non_max_suppression(y_mine, **kw)
    assert len(ref) == len(mine)
    for r, m in zip(ref, mine):
        assert r.shape == m.shape and torch.equal(r, m)
    assert torch.equal(y_ref, y_mine)  # the in-place xyxy side effect (utils/ops.py:243-244)


def test_nms_full_size_matches_unmodified_reference_on_cuda(R):
    """BASELINE configs[3] pages: 1536x1536, ~3000 candidates/page, conf .25, iou .7, max_det 2000."""
    b, nc, imgsz, strides = 4, 1, 1536, synth.P2_P5
    lv = synth.head_outputs(b, nc, imgsz, strides, seed=33, device=DEV, sparse_cls=3000)
    y = kzd.decode_inference(lv, strides, nc)
    ref = R.non_max_suppression(y.clone(), 0.25, 0.7, max_det=2000)
    mine = kzd.non_max_suppression(y.clone(), 0.25, 0.7, max_det=2000)
    assert [tuple(t.shape) for t in ref] == [tuple(t.shape) for t in mine]
    assert all(torch.equal(r, m) for r, m in zip(ref, mine))


def _stub_model(R, nc, strides):
    """The three attributes v8DetectionLoss.__init__ reads (utils/loss.py:158-177)."""
    det = _ref_detect(R, nc, strides)
    m = torch.nn.Module()
    m.model = torch.nn.ModuleList([det])
    m.args = types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
    return m.to(DEV)


def _batch(b, per_image, nc, seed):
    g = torch.Generator().manual_seed(seed)
    rows = []
    for j, n in enumerate(per_image):
        cxcy = torch.rand(n, 2, generator=g) * 0.8 + 0.1
        wh = torch.rand(n, 2, generator=g) * 0.08 + 0.03
        cls = torch.randint(0, nc, (n, 1), generator=g).float()
        rows.append(torch.cat((torch.full((n, 1), float(j)), cls, cxcy, wh), 1))
    t = torch.cat(rows)
    return {"batch_idx": t[:, 0].to(DEV), "cls": t[:, 1:2].to(DEV), "bboxes": t[:, 2:].to(DEV)}


def test_install_on_the_real_package_cuda(R):
    """install() patches the REAL ultralytics modules: the reference's own v8DetectionLoss.__call__ and
    Detect._inference -> ops.non_max_suppression flows give the same answers before and after patching,
    and after patching they really run on libkzd (loss, assigner targets, decoded boxes, NMS rows)."""
    b, nc, imgsz, strides = 2, 3, 640, synth.P2_P5
    feats = synth.head_outputs(b, nc, imgsz, strides, seed=51, device=DEV, sparse_cls=600)
    batch = _batch(b, [40, 17], nc, seed=6)

    def flow():
        model = _stub_model(R, nc, strides)
        crit = R.loss.v8DetectionLoss(model)           # built after patching, like nn/tasks.py:288-289
        seen = {}
        fwd = crit.assigner.forward

        def spy(*a, **k):
            seen["out"] = fwd(*a, **k)
            return seen["out"]

        crit.assigner.forward = spy
        xs = [f.clone().requires_grad_(True) for f in feats]
        total, items = crit(xs, batch)
        total.backward()
        grads = [x.grad.clone() for x in xs]
        det = model.model[0]
        with torch.no_grad():
            y = det._inference([f.clone() for f in feats])
        y_pre = y.clone()
        out = R.ops.non_max_suppression(y, 0.25, 0.7, max_det=300)
        return crit, total.detach(), items, seen["out"], y_pre, out, grads

    assert not kzd.installed()
    ref = flow()
    assert not isinstance(ref[0].assigner, kzd.TaskAlignedAssigner)
    names = kzd.install()
    try:
        assert len(names) == 6
        got = flow()
        assert isinstance(got[0].assigner, kzd.TaskAlignedAssigner) and got[0].assigner.counters is not None
    finally:
        kzd.uninstall()
    _check_assign_exact(got[3], ref[3])
    torch.testing.assert_close(got[2], ref[2], rtol=2e-5, atol=1e-6)   # loss items [box, cls, dfl]
    torch.testing.assert_close(got[1], ref[1], rtol=2e-5, atol=1e-6)
    for g_, r_ in zip(got[6], ref[6]):                                  # gradients through the patched bbox_decode
        torch.testing.assert_close(g_, r_, rtol=1e-4, atol=1e-7)
    st = got[0].stride.repeat_interleave(torch.tensor([f.shape[2] * f.shape[3] for f in feats], device=DEV))
    rel = ((got[4][:, :4] - ref[4][:, :4]).abs() / torch.maximum(ref[4][:, :4].abs(), st.reshape(1, 1, -1))).max()
    assert rel < RTOL and (got[4][:, 4:] - ref[4][:, 4:]).abs().max() <= 1e-6
    # NMS rows: each flow ran on its own decode output; like with like on the patched flow's y
    want = R.ops.non_max_suppression(got[4].clone(), 0.25, 0.7, max_det=300)
    assert all(torch.equal(a, w) for a, w in zip(got[5], want))


def test_install_leaves_out_of_scope_heads_to_the_reference(R):
    """Detect._inference is a class attribute, so subclasses inherit the patch: OBB (decode_bboxes overridden),
    end-to-end heads and other reg_max values must keep the reference's decode."""
    nc, strides = 2, synth.P3_P5
    lv = synth.head_outputs(1, nc, 256, strides, seed=3, device=DEV)
    kzd.install()
    try:
        import importlib

        inst = importlib.import_module("kuzushiji_vision_b200.install")  # the package attribute is the function

        det = _ref_detect(R, nc, strides)
        assert inst.detect_in_scope(det, R.Detect)
        e2e = _ref_detect(R, nc, strides)
        e2e.end2end = True
        assert not inst.detect_in_scope(e2e, R.Detect)
        with torch.no_grad():
            y_e2e = e2e._inference([x.clone() for x in lv])  # reference path: xyxy boxes (head.py:148)
            y_det = det._inference([x.clone() for x in lv])
        assert not torch.equal(y_e2e[:, :4], y_det[:, :4]) and (y_e2e[:, 4:] - y_det[:, 4:]).abs().max() <= 1e-6
        obb = R.head.OBB(nc=nc, ne=1, ch=(64, 64, 64)).to(DEV).eval()
        assert not inst.detect_in_scope(obb, R.Detect)
        seg = R.head.Segment(nc=nc, nm=4, npr=8, ch=(64, 64, 64)).to(DEV).eval()
        assert inst.detect_in_scope(seg, R.Detect)  # Segment / Pose use the plain Detect decode
    finally:
        kzd.uninstall()


def test_validator_process_batch_patched_on_the_real_package(R):
    """install(patch_validator=True) replaces DetectionValidator._process_batch (models/yolo/detect/val.py:209):
    same (N, 10) bool matrix as the reference's own numpy matching, on CUDA detections."""
    import importlib

    val = importlib.import_module("ultralytics.models.yolo.detect.val")
    iouv = torch.linspace(0.5, 0.95, 10, device=DEV)
    stub = types.SimpleNamespace(iouv=iouv)
    stub.match_predictions = types.MethodType(val.DetectionValidator.match_predictions, stub)
    b, nc, imgsz, strides = 2, 2, 640, synth.P2_P5
    lv = synth.head_outputs(b, nc, imgsz, strides, seed=92, device=DEV, sparse_cls=400)
    dets = kzd.non_max_suppression(kzd.decode_inference(lv, strides, nc), 0.25, 0.6, max_det=300)
    g = torch.Generator().manual_seed(4)
    cases = []
    for det in dets:
        pick = torch.randint(0, det.shape[0], (60,), generator=g).to(DEV)
        gt = det[pick, :4] + ((torch.rand(60, 4, generator=g) - 0.5) * 16).to(DEV)
        cases.append((det, gt, det[pick, 5].clone()))
    ref = [val.DetectionValidator._process_batch(stub, *c) for c in cases]
    names = kzd.install(patch_validator=True)
    try:
        assert len(names) == 7
        got = [val.DetectionValidator._process_batch(stub, *c) for c in cases]
    finally:
        kzd.uninstall()
    for a, r in zip(got, ref):
        assert a.dtype == torch.bool and a.device == r.device and torch.equal(a, r)
    assert any(r.any() for r in ref)


def test_nms_apriori_labels_match_unmodified_reference_on_cuda(R):
    b, nc, imgsz, strides = 2, 3, 320, synth.P2_P5
    lv = synth.head_outputs(b, nc, imgsz, strides, seed=46, device=DEV, sparse_cls=200)
    y = kzd.decode_inference(lv, strides, nc)
    g = torch.Generator().manual_seed(3)
    labels = [torch.cat((torch.randint(0, nc, (n, 1), generator=g).float(), torch.rand(n, 2, generator=g) * 250 + 30,
                         torch.rand(n, 2, generator=g) * 50 + 10), 1).to(DEV) for n in (12, 5)]
    for kw in (dict(conf_thres=0.4, iou_thres=0.6), dict(conf_thres=0.3, iou_thres=0.6, multi_label=True)):
        ref = R.non_max_suppression(y.clone(), labels=labels, **kw)
        mine = kzd.non_max_suppression(y.clone(), labels=labels, **kw)
        assert all(r.shape == m.shape and torch.equal(r, m) for r, m in zip(ref, mine))


def test_predictor_postprocess_patched_on_the_real_package(R):
    """install(patch_predictor=True): DetectionPredictor.postprocess (models/yolo/detect/predict.py:23-42) - NMS,
    scale_boxes to the original image, Results/Boxes construction (engine/results.py) - gives the same Results as
    the reference's per-image loop."""
    import importlib

    import numpy as np

    pred_mod = importlib.import_module("ultralytics.models.yolo.detect.predict")
    b, nc, imgsz, strides = 3, 2, 640, synth.P2_P5
    lv = synth.head_outputs(b, nc, imgsz, strides, seed=93, device=DEV, sparse_cls=300)
    for x in lv:
        x[2, 64:] = -20.0  # an image without detections
    y = kzd.decode_inference(lv, strides, nc)
    img = torch.zeros((b, 3, imgsz, imgsz), device=DEV)
    orig = [np.zeros((900, 1200, 3), np.uint8), np.zeros((640, 640, 3), np.uint8), np.zeros((333, 500, 3), np.uint8)]
    stub = types.SimpleNamespace(
        args=types.SimpleNamespace(conf=0.3, iou=0.6, agnostic_nms=False, max_det=200, classes=None),
        batch=[[f"page{i}.jpg" for i in range(b)]], model=types.SimpleNamespace(names={0: "char", 1: "mark"}))
    ref = pred_mod.DetectionPredictor.postprocess(stub, y.clone(), img, orig)
    names = kzd.install(patch_predictor=True)
    try:
        assert "predict.DetectionPredictor.postprocess" in names
        got = pred_mod.DetectionPredictor.postprocess(stub, y.clone(), img, orig)
    finally:
        kzd.uninstall()
    assert len(got) == len(ref) == b
    for g_, r_ in zip(got, ref):
        assert type(g_) is type(r_) and g_.path == r_.path and g_.orig_shape == r_.orig_shape
        assert g_.boxes.data.shape == r_.boxes.data.shape and torch.equal(g_.boxes.data, r_.boxes.data)
        assert torch.equal(g_.boxes.xywhn, r_.boxes.xywhn)
    assert ref[0].boxes.data.shape[0] > 0 and ref[2].boxes.data.shape[0] == 0
