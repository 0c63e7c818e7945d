"""Per-kernel device times INSIDE the PagePipeline graph replay (torch.profiler / CUPTI activity records; warm,
back-to-back replays of one pipeline on one stream, so kernels run as they do in the bench's single-stream
figure, not cold and serialised as under ncu). usage: python tools/graph_kernel_times.py [n_replays]"""
import collections, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda:0")
b, G, imgsz, nc, strides = 16, 600, 1280, 1, synth.P2_P5
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lv = synth.head_outputs(b, nc, imgsz, strides, seed=0, device=dev, sparse_cls=3000)
gl, gb, mg = synth.gt_pages(b, G, imgsz, nc, seed=0, device=dev)
cat = torch.cat([x.view(b, 64 + nc, -1) for x in lv], 2)
pd_scores = cat[:, 64:].permute(0, 2, 1).contiguous().sigmoid()
pipe = kzd.PagePipeline(b, synth.level_shapes(imgsz, strides), strides, nc, G, device=dev, conf_thres=0.25, iou_thres=0.7,
                        max_det=1000, fuse_decodes=True, fuse_filter=True)
pipe.load(levels=lv, pd_scores=pd_scores, gt_labels=gl, gt_bboxes=gb, mask_gt=mg)
for _ in range(5):
    pipe.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n_rep):
        pipe.step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
t_first, t_last = None, None
for e in prof.events():
    if e.device_type.name != "CUDA" or "Memset" in e.name or "Memcpy" in e.name:
        continue
    name = e.name.replace("kzd::", "").split("(")[0]
    agg.setdefault(name, []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
print("Per-kernel device times inside ONE PagePipeline graph replay (CUPTI). The two branches of the graph run concurrently: a\n"
      "kernel's duration includes the time its CTAs wait for SM slots behind the other branch (the persistent candidates kernel\n"
      "holds every SM's register file while it runs); stages alone: tools/time_assign.py, tools/time_nms.py.\n")
print("| kernel (inside the graph) | launches per step | mean us | us per step |\n|---|---|---|---|")
tot = 0.0
for k, v in agg.items():
    per = len(v) / n_rep
    tot += sum(v) / n_rep
    print(f"| `{k[-60:]}` | {per:g} | {sum(v) / len(v):.1f} | {sum(v) / n_rep:.1f} |")
print(f"\nsum of kernel times per step: {tot:.1f} us (kernels of the two graph branches overlap; wall time per step is lower)")
