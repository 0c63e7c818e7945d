import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth
dev="cuda:0"
d = synth.assigner_inputs(16, 600, 1280, 1, synth.P2_P5, seed=0, device=dev)
asg = kzd.TaskAlignedAssigner(10, 1, 0.5, 6.0)
args = (d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
for _ in range(4): out = asg(*args)
torch.cuda.synchronize(); print("ok", int(out[3].sum()))
