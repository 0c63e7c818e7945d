"""Phase timeline of assign_patch_kernel from a -DKZD_TRACE build (KZD_LIB_PATH=.../libkzd_trace.so): per-CTA clock64
stamps between the kernel's barriers, globaltimer at entry and exit. usage: python tools/trace_patch.py [G]"""
import ctypes as C, os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
from kuzushiji_vision_b200 import _lib, synth
dev = "cuda:0"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 600
d = synth.assigner_inputs(16, G, 1280, 1, synth.P2_P5, seed=0, device=dev)
asg = kzd.TaskAlignedAssigner(10, 1, 0.5, 6.0)
args = (d["pd_scores"], d["pd_bboxes"], d["anc_points"], d["gt_labels"], d["gt_bboxes"], d["mask_gt"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
for _ in range(3): asg(*args)
for cold in (True, False):
    if cold:
        flush.zero_(); flush_rd.sum()
    asg(*args); torch.cuda.synchronize()
    lib = C.CDLL(_lib.LIB_PATH)
    n_cta = 16 * ((G * 10 + 511) // 512)
    buf = np.zeros(8192 * 8, dtype=np.int64)
    lib.kzd_debug_trace(buf.ctypes.data_as(C.c_void_p), 8192 * 8)
    t = buf.reshape(8192, 8)[:n_cta]
    t0 = t[:, 0].min()
    names = ["entry->bar0", "loads (picks, exch, rows)", "contested scan", "targets/metric", "ticket", "tail"]
    print(f"cold={cold} G={G} CTAs={n_cta}: entry spread {(t[:,0].max()-t0)/1e3:.1f} us, last exit {(t[:,7].max()-t0)/1e3:.1f} us, median non-tail exit {np.median(t[:,7]-t0)/1e3:.1f} us")
    ph = [t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]]
    for n, v in zip(names[1:5], ph):
        print(f"   {n:28s} median {np.median(v)/1.965e3:6.2f} us  max {v.max()/1.965e3:6.2f} us")
    tail = t[t[:, 6] != 0]
    tl = (tail[:, 6] - tail[:, 5])
    print(f"   tail ({len(tail)} CTAs) median {np.median(tl)/1.965e3:.2f} us max {tl.max()/1.965e3:.2f} us")
    buf[:] = 0
