"""Ablation: marginal cost of each stage inside the double-buffered graph pipeline."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth, head
from kuzushiji_vision_b200.pipeline import PagePipeline
dev = torch.device("cuda:0")
b, G, imgsz, nc = 16, 600, 1280, 1

class P(PagePipeline):
    skip = ()
    def _train_half(self):
        with torch.no_grad():
            if "decode_train" in self.skip:
                boxes = boxes_px = self._px
            else:
                boxes, boxes_px = head.bbox_decode(self.anchor_points, self.pred_distri, self.stride_tensor)
            if "assign" in self.skip:
                return boxes, boxes_px, (None,) * 5
            tgt = self.assigner(self.pd_scores, boxes_px, self.anc_points, self.gt_labels, self.gt_bboxes, self.mask_gt)
        return boxes, boxes_px, tgt
    def _infer_half(self, decoded_event=None):
        if "nms" in self.skip or "decode_infer" in self.skip:
            with torch.no_grad():
                y = self._y if "decode_infer" in self.skip else head.decode_inference(self.levels, self.strides, self.nc)
            if decoded_event is not None:
                decoded_event.record(torch.cuda.current_stream(self.device))
            if "nms" not in self.skip:
                kzd.nms_batched  # not captured: needs host sync
            return y
        return super()._infer_half(decoded_event)

def run(skip):
    P.skip = skip
    pipes = []
    for _ in range(2):
        p = P.__new__(P)
        p._px = torch.zeros((b, 136000, 4), device=dev); p._y = torch.zeros((b, 5, 136000), device=dev)
        P.__init__(p, b, synth.level_shapes(imgsz, synth.P2_P5), synth.P2_P5, nc, G, device=dev, max_det=1000)
        pipes.append(p)
    d = synth.assigner_inputs(b, G, imgsz, nc, synth.P2_P5, seed=0, device=dev)
    lv = synth.head_outputs(b, nc, imgsz, synth.P2_P5, seed=0, device=dev, sparse_cls=3000)
    for p in pipes:
        p.load(levels=lv, pred_distri=d["pred_distri"], pd_scores=d["pd_scores"], gt_labels=d["gt_labels"], gt_bboxes=d["gt_bboxes"], mask_gt=d["mask_gt"])
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    def go(k):
        for i in range(k): pipes[i % 2].step(streams[i % 2])
    go(6); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    go(200)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    print(f"skip={skip!s:40s} us/step={e0.elapsed_time(e1)/200*1e3:7.1f}")
    del pipes

for skip in [(), ("nms",), ("assign",), ("nms", "assign"), ("decode_train", "assign"), ("decode_infer", "nms"), ("nms", "assign", "decode_infer"), ("nms","assign","decode_train")]:
    run(skip)
