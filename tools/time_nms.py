"""Device time of kzd_nms_suppress on the bench workload (own CUDA graph, L2 flushed before each replay).
usage: python tools/time_nms.py [label]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import _lib, synth
dev = torch.device("cuda:0")
b, nc, cands, max_det, strides = 16, 1, 3000, int(os.environ.get('KZD_TN_MAXDET', '1000')), synth.P2_P5
lv = synth.head_outputs(b, nc, 1280, strides, seed=100, device=dev, sparse_cls=cands)
y = kzd.decode_inference(lv, strides, nc)
A = y.shape[2]
lib = _lib.lib
pool, cap = min(A, 32768), 4096
keys = torch.empty((b, pool), dtype=torch.int64, device=dev)
rows = torch.empty((b, pool, 6), dtype=torch.float32, device=dev)
counts = torch.zeros((2, b), dtype=torch.int32, device=dev)
out_rows = torch.zeros((b, max_det, 6), dtype=torch.float32, device=dev)
ws_bytes = lib.kzd_nms_workspace_bytes(b, pool, cap)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # read after the write: L2 ends up clean
def st(): return C.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(lib.kzd_nms_filter(y.data_ptr(), b, A, nc, 0.25, 0, None, 0, 0, keys.data_ptr(), rows.data_ptr(), counts[0].data_ptr(), pool, st()), "filter")
def suppress():
    _lib.check(lib.kzd_nms_suppress(keys.data_ptr(), rows.data_ptr(), counts[0].data_ptr(), b, pool, cap, 0.7, 0, max_det, 30000, 7680.0,
                                    out_rows.data_ptr(), counts[1].data_ptr(), ws.data_ptr(), ws_bytes, st()), "suppress")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): suppress()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): suppress()
tot = 0
for i in range(22):
    flush.zero_(); flush_rd.sum(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    if i >= 2: tot += e0.elapsed_time(e1)
print(sys.argv[1] if len(sys.argv) > 1 else "", "nms_suppress graph us:", round(tot / 20 * 1e3, 1), "kept", counts[1].tolist()[:4],
      "checksum", float(out_rows.double().sum()))
import collections
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(10):
        flush.zero_(); flush_rd.sum(); g.replay()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and "nms_" in e.name], key=lambda e: e.time_range.start)
for e in evs:
    agg.setdefault(e.name.replace("kzd::", "").split("(")[0], []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
print(sys.argv[1] if len(sys.argv) > 1 else "", "kernels:", {k[-24:]: round(sum(v) / len(v), 1) for k, v in agg.items()})
# gaps between consecutive nms kernels of one replay (start of next - end of previous)
per = len(evs) // 10
if per >= 2:
    gaps = []
    for r in range(10):
        seq = evs[r * per:(r + 1) * per]
        gaps.append([round(seq[j + 1].time_range.start - seq[j].time_range.end, 1) for j in range(per - 1)])
    print("gaps us (replay 5):", gaps[5], "span us:", round(evs[6 * per - 1].time_range.end - evs[5 * per].time_range.start, 1))
