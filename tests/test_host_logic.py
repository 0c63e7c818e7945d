"""CPU: host-side logic of the package - the C ABI loads and exports every symbol of include/kzd.h,
argument validation, grid inference, synthetic pages, drop-in signatures, sharding."""
import inspect
import os
import re

import pytest
import torch

import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import _lib, dist as kdist, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "kzd.h")).read()
    declared = set(re.findall(r"\b(kzd_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(_lib.lib, name), f"libkzd.so does not export {name}"
    assert declared == set(_lib.EXPORTS)
    assert _lib.lib.kzd_version() == 100


def test_argument_validation_without_gpu():
    # bad arguments are rejected before any CUDA call is made
    rc = _lib.lib.kzd_decode_train(None, 0, None, None, 1, 1, None, None, None)
    assert rc == -1 and b"null" in _lib.lib.kzd_last_error()
    lv, _ = _lib.levels_array([(4, 4, 8.0)])
    rc = _lib.lib.kzd_tal_assign(1, 1, 1, 1, 1, 1, 1, 16, 1, 1, 99, 0.5, 6.0, 1e-9, lv, 1, 0, 0, 1, 1, 1, 1, 1, None,
                                 1, 0, 0, None)
    assert rc == -1 and b"topk" in _lib.lib.kzd_last_error()
    assert _lib.lib.kzd_assign_workspace_bytes(16, 136000, 600) > 16 * 136000 * 16
    assert _lib.lib.kzd_nms_workspace_bytes(2, 1000, 256) > 0


def test_cpu_tensors_are_refused():
    d = synth.assigner_inputs(1, 10, 128, 1, synth.P3_P5)
    with pytest.raises(kzd.KzdError):
        kzd.TaskAlignedAssigner(10, 1, 0.5, 6.0)(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"],
                                                  d["gt_bboxes"], d["mask_gt"])
    with pytest.raises(kzd.KzdError):
        kzd.non_max_suppression(torch.zeros(1, 5, 10))
    with pytest.raises(kzd.KzdError):
        kzd.bbox_decode(d["anchor_points"], d["pred_distri"])
    with pytest.raises(kzd.KzdError):
        kzd.decode_inference(d["levels"], synth.P3_P5, 1)


def test_g0_branch_is_host_logic():
    out = kzd.TaskAlignedAssigner(10, 3)(torch.rand(2, 50, 3), torch.rand(2, 50, 4), torch.rand(50, 2),
                                         torch.zeros(2, 0, 1), torch.zeros(2, 0, 4), torch.zeros(2, 0, 1))
    assert all(o.dtype == torch.float32 for o in out) and (out[0] == 3).all()


def test_nms_asserts_like_reference():
    with pytest.raises(AssertionError):
        kzd.non_max_suppression(torch.zeros(1, 5, 10), conf_thres=1.5)
    with pytest.raises(AssertionError):
        kzd.non_max_suppression(torch.zeros(1, 5, 10), iou_thres=-0.1)


def test_infer_levels():
    d = synth.assigner_inputs(1, 5, 256, 1, synth.P2_P5)
    assert kzd.infer_levels(d["anc_points"]) == [(64, 64, 4.0), (32, 32, 8.0), (16, 16, 16.0), (8, 8, 32.0)]
    pts = d["anc_points"].clone()
    pts[100, 0] += 0.25
    assert kzd.infer_levels(pts) is None
    assert kzd.infer_levels(torch.rand(50, 2)) is None
    # non-square level
    xs = (torch.arange(7.0) + 0.5) * 8
    ys = (torch.arange(3.0) + 0.5) * 8
    pts = torch.stack((xs.repeat(3), ys.repeat_interleave(7)), 1)
    assert kzd.infer_levels(pts) == [(3, 7, 8.0)]


def test_signatures_match_reference_interface():
    sig = inspect.signature(kzd.non_max_suppression)
    assert list(sig.parameters) == ["prediction", "conf_thres", "iou_thres", "classes", "agnostic", "multi_label",
                                    "labels", "max_det", "nc", "max_time_img", "max_nms", "max_wh", "in_place",
                                    "rotated"]
    assert sig.parameters["iou_thres"].default == 0.45 and sig.parameters["max_det"].default == 300
    sig = inspect.signature(kzd.TaskAlignedAssigner.__init__)
    assert [p for p in sig.parameters][1:] == ["topk", "num_classes", "alpha", "beta", "eps"]
    assert sig.parameters["topk"].default == 13 and sig.parameters["eps"].default == 1e-9
    sig = inspect.signature(kzd.TaskAlignedAssigner.forward)
    assert [p for p in sig.parameters][1:] == ["pd_scores", "pd_bboxes", "anc_points", "gt_labels", "gt_bboxes",
                                               "mask_gt"]
    a = kzd.TaskAlignedAssigner(topk=10, num_classes=4, alpha=0.5, beta=6.0)
    assert (a.topk, a.num_classes, a.bg_idx, a.alpha, a.beta, a.eps) == (10, 4, 4, 0.5, 6.0, 1e-9)


def test_synth_pages_are_seeded_and_well_formed():
    l1, b1, m1 = synth.gt_pages(2, 100, 1280, nc=4, seed=3)
    l2, b2, m2 = synth.gt_pages(2, 100, 1280, nc=4, seed=3)
    assert torch.equal(b1, b2) and torch.equal(l1, l2)
    n = m1.sum(1).flatten()
    assert ((n >= 70) & (n <= 100)).all()
    valid = m1.squeeze(-1).bool()
    wh = (b1[..., 2:] - b1[..., :2])[valid]
    assert (wh > 15.9).all() and (wh < 56.1).all()
    assert (b1[valid][:, :2] > 20).all() and (b1[valid][:, 2:] < 1260).all()
    assert (b1[~valid] == 0).all()
    assert synth.num_anchors(1280, synth.P2_P5) == 136000 and synth.num_anchors(1536, synth.P2_P5) == 195840
    lv = synth.head_outputs(1, 2, 128, synth.P2_P5, seed=1, sparse_cls=50)
    assert [tuple(x.shape) for x in lv] == [(1, 66, 32, 32), (1, 66, 16, 16), (1, 66, 8, 8), (1, 66, 4, 4)]


def test_shard_range_partitions_exactly():
    for n, w in [(128, 8), (16, 3), (5, 8), (0, 2)]:
        spans = [kdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_numa_binding_is_a_noop_without_a_gpu():
    """bind_to_gpu_numa never raises and leaves the affinity alone when there is no device / no topology."""
    import os

    from kuzushiji_vision_b200 import dist as kdist

    before = os.sched_getaffinity(0)
    assert kdist.bind_to_gpu_numa(0) is None or isinstance(kdist.bind_to_gpu_numa(0), int)
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_decode_filter_argument_validation_without_gpu():
    """kzd_decode_filter rejects a missing candidate pool / nc == 0 before touching the device."""
    import ctypes as C

    from kuzushiji_vision_b200 import _lib

    lv, _ = _lib.levels_array([(4, 4, 8.0)])
    ptrs = (C.c_void_p * 1)(None)
    rc = _lib.lib.kzd_decode_filter(ptrs, 0, lv, 1, 1, 1, None, None, None, 0.25, None, None, None, 16, None)
    assert rc != 0 and b"candidate pool" in _lib.lib.kzd_last_error()


def test_install_patches_and_restores_reference_attributes():
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("/root/reference not present")
    ref_loader.load()
    import ultralytics.nn.modules.head as u_head
    import ultralytics.utils.loss as u_loss
    import ultralytics.utils.ops as u_ops
    import ultralytics.utils.tal as u_tal

    orig = (u_tal.TaskAlignedAssigner, u_loss.TaskAlignedAssigner, u_ops.non_max_suppression,
            u_head.Detect._inference, u_loss.v8DetectionLoss.bbox_decode, u_loss.v8DetectionLoss.preprocess)
    names = kzd.install()
    try:
        assert len(names) == 6 and kzd.installed()
        assert issubclass(u_loss.TaskAlignedAssigner, kzd.TaskAlignedAssigner)
        assert u_tal.TaskAlignedAssigner is u_loss.TaskAlignedAssigner
        # CPU tensors go to the reference's own code, never to libkzd
        d = synth.assigner_inputs(1, 10, 128, 1, synth.P3_P5)
        asg = u_loss.TaskAlignedAssigner(topk=10, num_classes=1, alpha=0.5, beta=6.0)
        out = asg(d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
        ref = orig[0](topk=10, num_classes=1, alpha=0.5, beta=6.0)(d["pd_scores"], d["pd_bboxes"], d["anc_points"],
                                                                   d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
        assert all(torch.equal(a, b) for a, b in zip(out, ref))
        assert len(u_ops.non_max_suppression(torch.zeros(1, 5, 10))) == 1
    finally:
        kzd.uninstall()
    assert (u_tal.TaskAlignedAssigner, u_loss.TaskAlignedAssigner, u_ops.non_max_suppression, u_head.Detect._inference,
            u_loss.v8DetectionLoss.bbox_decode, u_loss.v8DetectionLoss.preprocess) == orig


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the kzd arm): one JSON line with the contract's
    keys, the same `config` object the kzd arm prints for this world size, the reference package (or the oracle port)
    named in `cpu_baseline`, no device copies and no kernel launches claimed."""
    import importlib.util
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "2"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pages/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    spec = importlib.util.spec_from_file_location("kzd_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    assert d["config"] == bench.workload_config(bench.CFG5, 1)  # what the kzd arm prints as `config`
    assert "configs[4]" in d["config"]["workload"] and d["scaling"] == bench.CFG5["scaling"]


def test_bench_algorithmic_bytes_match_survey_8d():
    """bench.py's roofline numerators are SURVEY.md 8(d)'s per-page formulas times the pages of one launch: pinned against the
    survey's own config values (config 2: assign 125 MB / batch, train decode 592 MB; config 4: inference decode 1.754 GB,
    NMS 230 MB), and the fused-filter NMS stage drops a meaningless fraction instead of printing it."""
    import importlib.util
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("kzd_bench2", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    ab = bench.algorithmic_bytes(16, 136000, 600, 1, 16 * 2940, 16 * 1000)
    assert abs(ab["assign"] / 1e6 - 125.3) < 0.5
    assert ab["decode_train"] == (272 + 16) * 136000 * 16 and abs(272 * 136000 * 16 / 1e6 - 591.9) < 0.1
    assert ab["decode_infer"] == 280 * 136000 * 16
    assert ab["decode_dual"] == (4 * 65 + 4 * 5 + 32) * 136000 * 16
    ab4 = bench.algorithmic_bytes(32, 195840, 600, 1, 32 * 2940, 32 * 2000)
    assert abs(ab4["decode_infer"] / 1e9 - 1.754) < 0.002 and abs(ab4["nms"] / 1e6 - 229.4) < 1.0
    ab3 = bench.algorithmic_bytes(16, 136000, 600, 4300, 0, 0)
    assert abs(ab3["assign"] / 1e9 - 37.5) < 0.1  # survey: 2.35 GB / page
    e = bench.stage_entry("nms", 0.036, 9_374_586_360, 6524.9, True)
    assert e["frac"] is None and "note" in e
    e = bench.stage_entry("assign", 0.0675, 125_350_400, 6524.9, True)
    assert abs(e["frac"] - 0.2846) < 0.002
    w1, w8 = bench.workload_config(bench.CFG5, 1), bench.workload_config(bench.CFG5, 8)
    assert w1["pages_per_gpu"] == 128 and w8["pages_per_gpu"] == 16 and w1["global_pages"] == w8["global_pages"] == 128
