import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
from b200seg import raw
from b200seg._lib import lib, ptr, stream_ptr
from b200seg.raw import conv_desc
L = lib()
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000.0
for (h, w, c) in [(256, 512, 48), (128, 256, 96), (256, 512, 256)]:
    x = torch.randn((1, h, w, c), device="cuda").to(torch.bfloat16)
    wt = torch.randn((c, c, 3, 3), device="cuda") * 0.05
    w_f, w_d = raw.pack_weight(wt)
    y = torch.empty_like(x)
    stats = torch.empty(148 * 2 * 1024, device="cuda")
    def fwd(st=True):
        d = conv_desc(1, h, w, c, c, 3, 1, c, c, False, False, st, 0)
        g = ctypes.c_int32(0)
        assert L.b200seg_conv2d_fwd(ctypes.byref(d), ptr(x), ptr(w_f), None, ptr(y), ptr(stats), ctypes.byref(g), stream_ptr()) == 0
    print("%dx%dx%d dbg=%s: %.1f us" % (h, w, c, os.environ.get("B200SEG_DBG", "0"), timeit(fwd)), flush=True)
