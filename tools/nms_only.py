import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth
dev="cuda:0"
lv = synth.head_outputs(16, 1, 1280, synth.P2_P5, seed=100, device=dev, sparse_cls=3000)
y = kzd.decode_inference(lv, synth.P2_P5, 1)
for md in (64, 64, 1000, 1000, 3000, 3000):
    rows, n = kzd.nms_batched(y, 0.25, 0.7, max_det=md, in_place=False)
torch.cuda.synchronize(); print("ok", n[:3])
