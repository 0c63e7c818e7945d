"""Time kzd_decode_train under the current KZD_TRAIN_VARIANT (L2 flushed, CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import synth
from oracle import torch_oracle as O
dev = "cuda:0"
b, A = 16, 136000
g = torch.Generator(device=dev).manual_seed(0)
pd = torch.randn((b, A, 64), device=dev, generator=g)
anc, st = O.make_anchors(synth.level_shapes(1280, synth.P2_P5), synth.P2_P5, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = O.decode_train(pd[:1], anc)
out, px = kzd.bbox_decode(anc, pd, st)
err = ((out[:1] - ref).abs() / ref.abs().clamp(min=1)).max().item()
tot = 0.0
for i in range(22):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); kzd.bbox_decode(anc, pd, st); e1.record(); torch.cuda.synchronize()
    if i >= 2: tot += e0.elapsed_time(e1)
us = tot / 20 * 1e3
print(f"variant={os.environ.get('KZD_TRAIN_VARIANT','default')} us={us:.1f} GB/s={(272+16)*A*b/us/1e3:.0f} (algorithmic 272B: {272*A*b/us/1e3:.0f}) relerr={err:.2e}")
