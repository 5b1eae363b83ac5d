"""Kineto (CUPTI) kernel-time breakdown of one eager fused training step — low-overhead complement to the ncu launch list."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from b200seg.module import B200SegModule

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 2048)
use_graph = len(sys.argv) > 3 and sys.argv[3] == "graph"
torch.manual_seed(0)
net = B200SegModule("ocrnet.HRNet_Mscale", 19, use_cuda_graph=use_graph).cuda().train()
with torch.no_grad():
    for n_, p_ in net.named_parameters():
        if p_.dim() == 4 and n_.startswith("backbone"):
            p_.normal_(0, (2.0 / (p_.shape[1] * p_.shape[2] * p_.shape[3])) ** 0.5)
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
images = torch.randn(1, 3, H, W, device="cuda")
gts = torch.randint(0, 19, (1, H, W), device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    loss = net({"images": images, "gts": gts})
    loss.backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev) if hasattr(ev[0], "device_time") else sum(e.cuda_time for e in ev)
from collections import defaultdict
agg = defaultdict(lambda: [0, 0.0])
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
for e in ev:
    d = e.time_range.end - e.time_range.start
    agg[e.name.split("(")[0][:60]][0] += 1
    agg[e.name.split("(")[0][:60]][1] += d
print("kernels %d  sum of kernel time %.3f ms  span %.3f ms" % (len(ev), sum(v[1] for v in agg.values()) / 1000.0, (t1 - t0) / 1000.0))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-60s n=%5d total %8.3f ms avg %8.2f us" % (k, n, us / 1000.0, us / n))
