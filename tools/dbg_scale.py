import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kuzushiji_vision_b200 as kzd
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
rows = (torch.rand(1, 200, 6, generator=g) * 640).to(dev)
for shp in ((900, 1200), (333, 500), (640, 640)):
    a = rows.clone(); kzd.scale_boxes_batched(a, [200], (640, 640), [shp])
    gain = min(640 / shp[0], 640 / shp[1]); pad = (round((640 - shp[1] * gain) / 2 - 0.1), round((640 - shp[0] * gain) / 2 - 0.1))
    b = rows.clone()[0]
    v = b[:, :4]
    v[..., 0] -= pad[0]; v[..., 1] -= pad[1]; v[..., 2] -= pad[0]; v[..., 3] -= pad[1]
    c = v.clone(); d = v.clone()
    v[..., :4] /= gain
    c = c * torch.tensor(1.0 / gain, device=dev)          # reciprocal in double, rounded to float
    d = d * (torch.tensor(1.0, device=dev) / torch.tensor(gain, device=dev))   # float reciprocal of float gain
    e = (rows.clone()[0][:, :4] - torch.tensor([pad[0], pad[1], pad[0], pad[1]], device=dev)) / torch.tensor(gain, device=dev)  # true division
    for name, t in (("inplace /=", v), ("mul double-recip", c), ("mul float-recip", d), ("true div", e)):
        t = t.clone()
        t[..., 0].clamp_(0, shp[1]); t[..., 1].clamp_(0, shp[0]); t[..., 2].clamp_(0, shp[1]); t[..., 3].clamp_(0, shp[0])
        print(shp, name, "equal to kzd:", bool(torch.equal(t, a[0, :, :4])), "n diff", int((t != a[0, :, :4]).sum()))
