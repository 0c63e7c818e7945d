"""us/step of the bench's timed region (depth PagePipelines in flight, fused decodes + filter) for A/B runs.
usage: python tools/pipe_value.py [label] [depth]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth
dev = torch.device("cuda:0")
b, G, imgsz, nc, strides = 16, int(os.environ.get('KZD_PV_G', '600')), 1280, 1, synth.P2_P5
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fuse_filter = os.environ.get("KZD_BENCH_FILTER", "1") != "0"
lv = synth.head_outputs(b, nc, imgsz, strides, seed=0, device=dev, sparse_cls=3000)
gl, gb, mg = synth.gt_pages(b, G, imgsz, nc, seed=0, device=dev)
cat = torch.cat([x.view(b, 64 + nc, -1) for x in lv], 2)
pd_scores = cat[:, 64:].permute(0, 2, 1).contiguous().sigmoid()
del cat
skip = set(filter(None, os.environ.get("KZD_SKIP", "").split(",")))
if "assign" in skip:
    kzd.PagePipeline._assign = lambda self, px: (None,) * 5
if "nms" in skip:
    kzd.PagePipeline._nms = lambda self, y: None
pipes = []
for _ in range(depth):
    p = kzd.PagePipeline(b, synth.level_shapes(imgsz, strides), strides, nc, G, device=dev, topk=10, alpha=0.5, beta=6.0,
                         conf_thres=0.25, iou_thres=0.7, max_det=1000, fuse_decodes=True, fuse_filter=fuse_filter)
    p.load(levels=lv, pd_scores=pd_scores, gt_labels=gl, gt_bboxes=gb, mask_gt=mg)
    pipes.append(p)
streams = [torch.cuda.Stream() for _ in range(depth)]
cur = torch.cuda.current_stream()
def go(k):
    for s in streams: s.wait_stream(cur)
    for i in range(k): pipes[i % depth].step(streams[i % depth])
    for s in streams: cur.wait_stream(s)
go(9); torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(300); e1.record(); torch.cuda.synchronize()
    res.append(round(e0.elapsed_time(e1) / 300 * 1e3, 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): pipes[0].step()
e1.record(); torch.cuda.synchronize()
print(sys.argv[1] if len(sys.argv) > 1 else "", f"depth={depth} us/step:", res, "single:", round(e0.elapsed_time(e1) / 100 * 1e3, 1),
      "kept", sum(pipes[0].results()[1]), "skip", sorted(skip))
