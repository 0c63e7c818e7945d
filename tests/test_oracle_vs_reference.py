"""CPU, build container only: live comparison of the oracle with the unmodified reference imported
from /root/reference (skipped where that tree does not exist, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ref_loader
from oracle import torch_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


@pytest.mark.parametrize("cfg", [
    (2, 60, 320, 1, (8, 16, 32), "columns", 10, 0.5),
    (1, 50, 256, 4, (4, 8, 16, 32), "uniform", 10, 0.5),
    (1, 40, 256, 2, (4, 8, 16, 32), "columns", 13, 1.0),
])
def test_assign_live(R, cfg):
    from kuzushiji_vision_b200 import synth

    b, G, imgsz, nc, strides, layout, topk, alpha = cfg
    d = synth.assigner_inputs(b, G, imgsz, nc, strides, seed=21, layout=layout)
    args = (d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
    ref = R.TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=alpha, beta=6.0)(*args)
    mine = O.tal_assign(*args, topk, nc, alpha, 6.0)
    for r, m in zip(ref, mine):
        assert r.dtype == m.dtype
        if r.dtype == torch.float32:
            torch.testing.assert_close(m, r, rtol=1e-5, atol=1e-8)
        else:
            assert torch.equal(r, m)


def test_ciou_live(R):
    g = torch.Generator().manual_seed(5)
    a = torch.rand(2000, 2, generator=g) * 100
    gt = torch.cat((a, a + torch.rand(2000, 2, generator=g) * 50 + 1), 1)
    c = torch.rand(2000, 2, generator=g) * 100
    pd = torch.cat((c, c + torch.rand(2000, 2, generator=g) * 50 + 1), 1)
    ref = R.bbox_iou(gt, pd, xywh=False, CIoU=True).squeeze(-1)
    torch.testing.assert_close(O.ciou_gt_pd(gt, pd), ref, rtol=0, atol=0)


def test_anchors_live(R):
    feats = [torch.zeros(1, 1, 5, 7), torch.zeros(1, 1, 3, 4)]
    ra, rs = R.make_anchors(feats, [8, 16], 0.5)
    oa, os_ = O.make_anchors([(5, 7), (3, 4)], [8, 16])
    assert torch.equal(ra, oa) and torch.equal(rs, os_)
    from kuzushiji_vision_b200 import make_anchors

    ka, ks = make_anchors(feats, [8, 16], 0.5)
    assert torch.equal(ra, ka) and torch.equal(rs, ks)


def test_nms_live(R):
    from kuzushiji_vision_b200 import synth

    lv = synth.head_outputs(2, 3, 256, synth.P2_P5, seed=4, sparse_cls=300)
    y = O.decode_infer(lv, synth.P2_P5, 3)
    for kw in (dict(conf_thres=0.25, iou_thres=0.7, max_det=300), dict(conf_thres=0.25, iou_thres=0.5, agnostic=True),
               dict(conf_thres=0.3, iou_thres=0.6, multi_label=True, max_det=50)):
        a = R.non_max_suppression(y.clone(), **kw)
        b_ = O.non_max_suppression(y.clone(), **kw)
        assert all(torch.equal(x, z) for x, z in zip(a, b_))


def test_loss_tail_and_preprocess_live(R):
    import types

    from kuzushiji_vision_b200 import synth

    b, G, imgsz, nc = 2, 40, 320, 3
    d = synth.assigner_inputs(b, G, imgsz, nc, synth.P2_P5, seed=4)
    logits = torch.cat([l.view(b, 64 + nc, -1) for l in d["levels"]], 2)[:, 64:].permute(0, 2, 1).contiguous()
    tl, tb, ts, fg, gi = O.tal_assign(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"],
                                      d["mask_gt"], 10, nc, 0.5, 6.0)
    tss = max(ts.sum(), 1)
    ref_cls = torch.nn.BCEWithLogitsLoss(reduction="none")(logits, ts).sum() / tss
    pred_bboxes = R.dist2bbox(d["pred_distri"].view(b, -1, 4, 16).softmax(3).matmul(torch.arange(16.0)),
                              d["anchor_points"], xywh=False)
    ref_box, ref_dfl = R.loss.BboxLoss(16)(d["pred_distri"], pred_bboxes, d["anchor_points"], tb / d["stride_tensor"],
                                           ts, tss, fg)
    mine = O.loss_tail(logits, d["pred_distri"], d["anchor_points"], d["stride_tensor"], tb, ts, fg)
    torch.testing.assert_close(mine, torch.stack((ref_box, ref_cls, ref_dfl)), rtol=1e-6, atol=0)
    # preprocess
    g = torch.Generator().manual_seed(0)
    rows = [torch.cat((torch.full((n, 1), float(j)), torch.randint(0, 5, (n, 1), generator=g).float(),
                       torch.rand(n, 4, generator=g)), 1) for j, n in enumerate([5, 0, 17, 3])]
    t = torch.cat(rows)
    sc = torch.tensor([1280.0, 960.0, 1280.0, 960.0])
    ref = R.v8DetectionLoss.preprocess(types.SimpleNamespace(device="cpu"), t.clone(), 4, sc)
    assert torch.equal(ref, O.preprocess(t.clone(), 4, sc))


def test_scale_boxes_live(R):
    g = torch.Generator().manual_seed(2)
    boxes = torch.rand(200, 4, generator=g) * 700 - 30
    for img1, img0 in [((640, 640), (1080, 1920, 3)), ((320, 512), (700, 500)), ((1280, 1280), (333, 333))]:
        ref = R.ops.scale_boxes(img1, boxes.clone(), img0)
        assert torch.equal(ref, O.scale_boxes(img1, boxes.clone(), img0))


def test_validator_matching_live(R):
    """box_iou / match_predictions restatement against the reference's own functions on random cases."""
    import importlib
    import types

    val = importlib.import_module("ultralytics.models.yolo.detect.val")
    iouv = torch.linspace(0.5, 0.95, 10)
    stub = types.SimpleNamespace(iouv=iouv)
    stub.match_predictions = types.MethodType(val.DetectionValidator.match_predictions, stub)
    g = torch.Generator().manual_seed(21)
    for m, n, nc in ((25, 60, 2), (80, 200, 1), (3, 1, 1)):
        xy = torch.rand(m, 2, generator=g) * 500
        wh = torch.rand(m, 2, generator=g) * 60 + 10
        gt = torch.cat((xy, xy + wh), 1)
        gcls = torch.randint(0, nc, (m,), generator=g).float()
        src = torch.randint(0, m, (n,), generator=g)
        box = gt[src] + (torch.rand(n, 4, generator=g) - 0.5) * 12
        det = torch.cat((box, torch.rand(n, 1, generator=g).sort(0, descending=True).values,
                         torch.randint(0, nc, (n, 1), generator=g).float()), 1)
        assert torch.equal(O.box_iou(gt, det[:, :4]), R.metrics.box_iou(gt, det[:, :4]))
        assert torch.equal(O.process_batch(det, gt, gcls, iouv), val.DetectionValidator._process_batch(stub, det, gt, gcls))


def test_nms_apriori_labels_live(R):
    """labels= (autolabelling, utils/ops.py:256-261): oracle restatement against the reference."""
    from kuzushiji_vision_b200 import synth

    g = torch.Generator().manual_seed(8)
    d = synth.assigner_inputs(2, 10, 256, 3, synth.P3_P5, seed=4)
    y = O.decode_infer(d["levels"], synth.P3_P5, 3)
    labels = []
    for n in (6, 0):
        lb = torch.cat((torch.randint(0, 3, (n, 1), generator=g).float(), torch.rand(n, 2, generator=g) * 200 + 20,
                        torch.rand(n, 2, generator=g) * 40 + 8), 1)
        labels.append(lb)
    for kw in (dict(conf_thres=0.5, iou_thres=0.6), dict(conf_thres=0.3, iou_thres=0.6, multi_label=True, classes=[0, 2])):
        ref = R.non_max_suppression(y.clone(), labels=labels, **kw)
        got = O.non_max_suppression(y.clone(), labels=labels, **kw)
        assert all(torch.equal(a, b) for a, b in zip(got, ref))
        assert (ref[0][:, 4] == 1.0).sum() >= 1
