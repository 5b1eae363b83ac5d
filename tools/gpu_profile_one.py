"""Launch one convolution shape a few times (target for `ncu --set full -k regex:...`)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
from b200seg import raw
h, w, cin, cout, k = [int(a) for a in sys.argv[1:6]]
kc = int(sys.argv[6]) if len(sys.argv) > 6 else 0
x = torch.randn((1, h, w, cin), device="cuda").to(torch.bfloat16)
wt = torch.randn((cout, cin, k, k), device="cuda") * 0.05
w_f, w_d = raw.pack_weight(wt)
bias = torch.zeros(cout, device="cuda")
for _ in range(4):
    y, st = raw.conv2d_fwd(x, w_f, bias if cout == 512 else None, emit_stats=True, force_kc=kc)
dy = torch.randn_like(y)
dw = torch.zeros_like(wt)
for _ in range(2):
    raw.conv2d_wgrad(x, dy, dw, cout, k, 1)
torch.cuda.synchronize()
print("done")
