"""Per-call device times of the C-ABI entry points on the bench workload (CUDA events, warm, many
iterations back to back so the GPU never waits for the host). Writes gpurun_out/stage_times.json.

    python tools/stage_times.py [--iters 50] [--config 2|4]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd  # noqa: E402
from kuzushiji_vision_b200 import _lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--config", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
CFG = {  # BASELINE.json configs (3: class count of config 3 at a batch that fits the oracle-free timing run)
    1: (1, 200, 640, 1, 1000, 300, (8, 16, 32)),
    2: (16, 600, 1280, 1, 3000, 1000, (4, 8, 16, 32)),
    3: (4, 600, 1280, 4300, 3000, 1000, (4, 8, 16, 32)),
    4: (32, 600, 1536, 1, 3000, 2000, (4, 8, 16, 32)),
    5: (16, 1000, 1280, 1, 3000, 1000, (4, 8, 16, 32)),
}
b, G, imgsz, nc, cands, max_det, strides = CFG[args.config]
d = synth.assigner_inputs(b, G, imgsz, nc, strides, seed=0, device=dev)
lv = synth.head_outputs(b, nc, imgsz, strides, seed=100, device=dev, sparse_cls=cands)
A = d["anc_points"].shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=args.iters, flush_l2=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3  # us


res = {"config": dict(b=b, G=G, imgsz=imgsz, nc=nc, A=A, cands=cands, max_det=max_det)}
asg = kzd.TaskAlignedAssigner(10, nc, 0.5, 6.0)
aargs = (d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
res["decode_train_us"] = timeit(lambda: kzd.bbox_decode(d["anchor_points"], d["pred_distri"]))
res["decode_train_px_us"] = timeit(lambda: kzd.bbox_decode(d["anchor_points"], d["pred_distri"], d["stride_tensor"]))
res["assign_us"] = timeit(lambda: asg(*aargs))
res["decode_infer_us"] = timeit(lambda: kzd.decode_inference(lv, strides, nc))
y = kzd.decode_inference(lv, strides, nc)

# NMS split into its two C-ABI calls, no host sync in between
lib = _lib.lib
pool, cap = min(A, 32768), 4096
keys = torch.empty((b, pool), dtype=torch.int64, device=dev)
rows = torch.empty((b, pool, 6), dtype=torch.float32, device=dev)
counts = torch.empty((2, b), dtype=torch.int32, device=dev)
out_rows = torch.empty((b, max_det, 6), dtype=torch.float32, device=dev)
ws_bytes = lib.kzd_nms_workspace_bytes(b, pool, cap)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def nms_filter():
    _lib.check(lib.kzd_nms_filter(y.data_ptr(), b, A, nc, 0.25, 0, None, 0, 0, keys.data_ptr(), rows.data_ptr(),
                                  counts[0].data_ptr(), pool, st), "filter")


def nms_suppress():
    _lib.check(lib.kzd_nms_suppress(keys.data_ptr(), rows.data_ptr(), counts[0].data_ptr(), b, pool, cap, 0.7, 0,
                                    max_det, 30000, 7680.0, out_rows.data_ptr(), counts[1].data_ptr(), ws.data_ptr(),
                                    ws_bytes, st), "suppress")


res["nms_filter_us"] = timeit(nms_filter)
nms_filter()
res["nms_suppress_us"] = timeit(nms_suppress)
res["nms_counts"] = counts.cpu().tolist()
res["nms_api_us"] = timeit(lambda: kzd.nms_batched(y, 0.25, 0.7, max_det=max_det, in_place=False))
res["glue_mul_us"] = timeit(lambda: d["pd_bboxes"] * d["stride_tensor"])
peak = 6524.9
res["gbs"] = {
    "decode_train": 272 * A * b / res["decode_train_us"] / 1e3,
    "decode_infer": (272 + 8 * nc) * A * b / res["decode_infer_us"] / 1e3,
    "assign": ((57 + 8 / b) * A + 24 * G) * b / res["assign_us"] / 1e3,
    "nms": ((32 + 4 * nc) * A * b + 24 * 2 * cands * b) / (res["nms_filter_us"] + res["nms_suppress_us"]) / 1e3,
}
if nc > 1:
    res["gbs"]["assign"] = ((49 + 4 * nc + 8 / b) * A + 24 * G) * b / res["assign_us"] / 1e3
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/stage_times_cfg{args.config}.json", "w"), indent=1)
print(json.dumps(res, indent=1))
