"""One-off characterisation of the aten CUDA behaviours the assigner parity depends on
(SURVEY.md section 7 'hard parts'): pow special-casing and torch.topk's tie / zero-fill order.
Writes gpurun_out/probe.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = {}
dev = "cuda:0"
x = torch.rand(1 << 20, device=dev)
out["pow_half_is_sqrt"] = bool(torch.equal(x.pow(0.5), x.sqrt()))
out["pow_two_is_mul"] = bool(torch.equal(x.pow(2), x * x))
out["gpu"] = torch.cuda.get_device_name(0)

# zero-fill order of topk: rows that are almost all zeros, k larger than the number of positives
res = {}
for A in (8400, 34000, 136000):
    for rows in (1, 600, 9600):
        m = torch.zeros(rows, A, device=dev)
        pos = torch.randint(1000, A, (rows, 4), device=dev)
        m.scatter_(1, pos, torch.rand(rows, 4, device=dev) + 0.1)
        idx = torch.topk(m, 10, dim=-1).indices
        vals = m.gather(1, idx)
        zero_picks = [sorted(idx[r][vals[r] == 0].tolist()) for r in range(min(rows, 3))]
        n_pos = int((vals[0] > 0).sum())
        expect = list(range(10 - n_pos))
        all_lowest = True
        for r in range(rows if rows <= 600 else 200):
            zp = sorted(idx[r][vals[r] == 0].tolist())
            npos = int((vals[r] > 0).sum())
            if zp != list(range(10 - npos)):
                all_lowest = False
                break
        res[f"A{A}_rows{rows}"] = {"zero_picks_first_rows": zero_picks, "lowest_index_first": all_lowest}
out["topk_zero_fill"] = res

# positive ties at the k-th boundary
m = torch.zeros(4, 5000, device=dev)
m[:, [100, 300, 500, 900, 1500]] = 0.5
m[:, 2000] = 0.9
out["topk_pos_ties_k3"] = sorted(torch.topk(m, 3, dim=-1).indices[0].tolist())
out["topk_pos_ties_k4"] = sorted(torch.topk(m, 4, dim=-1).indices[0].tolist())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out))
