"""CPU: the oracle (oracle/torch_oracle.py) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py). Assign / NMS / train-decode must be bit-identical on the CPU;
the inference decode differs from the reference's conv-based DFL only by summation order."""
import pytest
import torch

from conftest import load_golden
from oracle import torch_oracle as O

PATH_CASES = ["path_p3p5_nc1", "path_p2p5_nc5", "path_p2p5_topk13"]


@pytest.mark.parametrize("name", PATH_CASES + ["assign_micro", "assign_edge"])
def test_assign_matches_reference_golden(name):
    g = load_golden(name)
    b, G, imgsz, nc, topk = (int(v) for v in g["meta"])
    alpha, beta = (float(v) for v in g["hyp"])
    out = O.tal_assign(g["pd_scores"], g["pd_bboxes"], g["anc_points"], g["gt_labels"], g["gt_bboxes"], g["mask_gt"],
                       topk, nc, alpha, beta)
    assert torch.equal(out[0], g["target_labels"])
    assert torch.equal(out[1], g["target_bboxes"])
    assert torch.equal(out[3], g["fg_mask"])
    assert torch.equal(out[4], g["target_gt_idx"])
    torch.testing.assert_close(out[2], g["target_scores"], rtol=1e-5, atol=1e-8)
    assert out[0].dtype == torch.int64 and out[3].dtype == torch.bool and out[4].dtype == torch.int64


def test_micro_case_known_answer():
    """SURVEY.md Appendix E style known-answer: every fg row has exactly one non-zero score, background
    rows carry GT 0's label and box, the padded GT owns nothing."""
    g = load_golden("assign_micro")
    fg = g["fg_mask"][0]
    ts = g["target_scores"][0]
    assert int(fg.sum()) == 19
    assert ((ts > 0).sum(1)[fg] == 1).all()
    assert (ts[~fg] == 0).all()
    assert (g["target_gt_idx"][0] != 2).all()
    assert torch.equal(g["target_bboxes"][0][~fg], g["gt_bboxes"][0, 0].expand(int((~fg).sum()), 4))
    assert (g["target_labels"][0][~fg] == int(g["gt_labels"][0, 0, 0])).all()


@pytest.mark.parametrize("name", PATH_CASES)
def test_decode_matches_reference_golden(name):
    g = load_golden(name)
    nc = int(g["meta"][3])
    strides = [int(s) for s in g["strides"]]
    levels = [g[f"level{i}"] for i in range(len(strides))]
    assert torch.equal(O.decode_train(g["pred_distri"], g["anchor_points"]), g["decode_train"])
    y = O.decode_infer(levels, strides, nc)
    ref = g["decode_infer"]
    A = ref.shape[2]
    _, st = O.make_anchors([l.shape[2:] for l in levels], strides)
    scale = st.view(1, 1, A)
    rel = ((y[:, :4] - ref[:, :4]).abs() / torch.maximum(ref[:, :4].abs(), scale)).max()
    assert rel < 1e-5, rel
    assert (y[:, 4:] - ref[:, 4:]).abs().max() < 1e-6


@pytest.mark.parametrize("name", ["nms_nc1", "nms_nc5", "nms_nc5_agnostic", "nms_nc5_multilabel", "nms_nc5_classes"])
def test_nms_matches_reference_golden(name):
    g = load_golden(name)
    kw = {k[3:]: (v.tolist() if hasattr(v, "tolist") else v) for k, v in g.items() if k.startswith("kw_")}
    for k in ("max_det",):
        kw[k] = int(kw[k])
    for k in ("agnostic", "multi_label"):
        if k in kw:
            kw[k] = bool(kw[k])
    pred = g["prediction"].clone()
    out = O.non_max_suppression(pred, **kw)
    assert torch.equal(pred, g["mutated"])  # in-place xywh -> xyxy side effect
    for i in range(int(g["n_img"])):
        assert torch.equal(out[i], g[f"out{i}"])


def test_greedy_nms_restatement_matches_torchvision():
    import torchvision

    gen = torch.Generator().manual_seed(3)
    xy = torch.rand(400, 2, generator=gen) * 200
    wh = torch.rand(400, 2, generator=gen) * 60 + 4
    boxes = torch.cat((xy, xy + wh), 1)
    scores = torch.rand(400, generator=gen)
    scores[10:20] = scores[5]  # ties exercise the stable sort
    for thr in (0.3, 0.5, 0.7):
        assert torch.equal(O.greedy_nms(boxes, scores, thr), torchvision.ops.nms(boxes, scores, thr))


def test_assign_g0_branch():
    out = O.tal_assign(torch.rand(2, 50, 3), torch.rand(2, 50, 4), torch.rand(50, 2), torch.zeros(2, 0, 1),
                       torch.zeros(2, 0, 4), torch.zeros(2, 0, 1), 10, 3)
    assert all(o.dtype == torch.float32 for o in out)
    assert (out[0] == 3).all() and out[2].shape == (2, 50, 3)


def test_validator_matching_matches_reference_golden():
    """oracle.process_batch (box_iou + match_predictions restated) against what the unmodified reference's
    DetectionValidator._process_batch returned (tests/golden/make_golden.py::match_case)."""
    g = load_golden("match_val")
    iouv = g["iouv"]
    for k in range(int(g["n_cases"])):
        det, gt, gcls, want = g[f"det{k}"], g[f"gt{k}"], g[f"gcls{k}"], g[f"correct{k}"]
        if det.shape[0] == 0 or gt.shape[0] == 0:
            assert not want.any()
            continue
        got = O.process_batch(det, gt, gcls, iouv)
        assert got.dtype == torch.bool and torch.equal(got, want), k
