"""Per-role timeline of CTA 0 of the halo convolution (B200SEG_DBG=8): usage gpu_timeline.py H W C"""
import os, sys, ctypes
os.environ["B200SEG_DBG"] = "8"   # needs a library built with: make EXTRA=-DB200SEG_TIMELINE
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
from b200seg import raw
from b200seg._lib import lib, ptr, stream_ptr
from b200seg.raw import conv_desc
L = lib()
h, w, c = [int(a) for a in sys.argv[1:4]]
stats_on = int(sys.argv[4]) if len(sys.argv) > 4 else 1
x = torch.randn((1, h, w, c), device="cuda").to(torch.bfloat16)
wt = torch.randn((c, c, 3, 3), device="cuda") * 0.05
w_f, _ = raw.pack_weight(wt)
y = torch.empty_like(x)
stats = torch.zeros(296 * 2 * 1024 + 4096, device="cuda")
d = conv_desc(1, h, w, c, c, 3, 1, c, c, False, False, bool(stats_on), 0)
g = ctypes.c_int32(0)
for _ in range(5):
    assert L.b200seg_conv2d_fwd(ctypes.byref(d), ptr(x), ptr(w_f), None, ptr(y), ptr(stats), ctypes.byref(g), stream_ptr()) == 0
torch.cuda.synchronize()
ts = stats[296 * 2 * 1024:296 * 2 * 1024 + 2 * 7 * 16].view(torch.int64).cpu().view(7, 16)
t0 = int(ts[6, 0])
names = ["prod A issued", "mma a_full seen", "mma issued+commit", "epi tfull seen", "epi done", "epi tmem loaded", "misc(start,end)"]
print("shape", h, w, c, "stats", stats_on)
for r in range(7):
    print("%-22s" % names[r], [int(v) - t0 if int(v) else None for v in ts[r][:9]])
