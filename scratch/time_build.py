import sys, time; sys.path.insert(0, '.')
import torch
from bench import build_data
from crispr_bean_b200.svi import SviEngine
from crispr_bean_b200.device_pack import DeviceScreen
t=time.time(); data = build_data("c5_genome_scale", 101); print("build_data", time.time()-t)
dev = torch.device("cuda")
torch.cuda.synchronize()
for i in range(3):
    t=time.time(); eng = SviEngine(data, "MixtureNormal", dev, num_steps=100); torch.cuda.synchronize(); print("engine total", time.time()-t)
def T(label, f):
    torch.cuda.synchronize(); t=time.time(); r=f(); torch.cuda.synchronize(); print(f"  {label:40s} {1e3*(time.time()-t):8.2f} ms"); return r
x = T("stack permute (host)", lambda: torch.stack([data.X_masked.permute(2,0,1), data.X_bcmatch_masked.permute(2,0,1)]))
xd = T("to device + contiguous", lambda: x.to(device=dev, dtype=torch.float32).contiguous())
x2 = T("stack (R,B,G) host", lambda: torch.stack([data.X_masked, data.X_bcmatch_masked]))
x2p = T("pin", lambda: x2.pin_memory())
xd2 = T("H2D pinned", lambda: x2p.to(dev, non_blocking=True))
xd3 = T("H2D pageable", lambda: x2.to(dev))
xd4 = T("permute on device", lambda: xd3.permute(0,3,1,2).contiguous())
x64 = T("double()", lambda: xd.double())
n64 = x64.sum(-1)
rc = T("row_const lgamma f64", lambda: (torch.lgamma(n64 + 1) - torch.lgamma(x64 + 1).sum(-1) + torch.xlogy(x64, x64 / n64.clamp(min=1.0).unsqueeze(-1)).sum(-1)))
T("ll_const .sum() float()", lambda: float(rc.sum()))
ac = data.allele_counts_control
T("allele counts permute+to", lambda: ac[:, 0].permute(1, 0, 2).to(device=dev, dtype=torch.float32).contiguous())
T("row_mask", lambda: data.repguide_mask.T.to(torch.uint8).to(dev).contiguous())
T("guide_variant", lambda: data.guide_variant.to(dev).contiguous())
