"""Device time of the fused decode (+filter) call on the bench workload (own CUDA graph, L2 flushed).
usage: python tools/time_decode.py [label]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth
dev = torch.device("cuda:0")
b, nc, strides = int(os.environ.get("KZD_TD_B", "16")), int(os.environ.get("KZD_TD_NC", "1")), synth.P2_P5
lv = synth.head_outputs(b, nc, 1280, strides, seed=100, device=dev, sparse_cls=3000, on_device=True)
keys = torch.zeros((b, 32768), dtype=torch.int64, device=dev); rows = torch.zeros((b, 32768, 6), device=dev)
counts = torch.zeros(b, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # read after the write: L2 ends up clean
res = {}
for name, fn in (("infer", lambda: kzd.decode_inference(lv, strides, nc)),
                 ("dual+filter", lambda: kzd.decode_inference(lv, strides, nc, with_train_boxes=True, nms_pool=(0.25, keys, rows, counts)))):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): keep = fn()
    tot = 0
    for i in range(22):
        flush.zero_(); flush_rd.sum(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        if i >= 2: tot += e0.elapsed_time(e1)
    res[name] = round(tot / 20 * 1e3, 1)
    # back to back (no flush): sustained
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    res[name + "_b2b"] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
print(sys.argv[1] if len(sys.argv) > 1 else "", "decode us:", res, "cands", int(counts.sum()))
