"""Static trace of one training step on the CPU (no GPU, no kernels): runs the real step program (b200seg.model /
b200seg.engine / b200seg.raw) on `meta` tensors with every launching C-ABI entry point replaced by a recorder, and
prints, per kernel class, the number of launches, the algorithmic FLOPs / bytes and the roofline time
sum_launches max(FLOPs / P, bytes / B). Host-only planner queries (stats_elems, ws_bytes, grids ...) still go to the real
library, so buffer sizes are the real ones.

    python -O tools/trace_step.py [--arch ocrnet.HRNet_Mscale] [--height 1024 --width 2048] [--tflops 1386.7 --gbs 6650]

-O is required: the wrappers assert `is_cuda`, and the point of this tool is to run them without a device.
Bytes of a launch = sizes of the tensors handed to it (views count with their own extent), i.e. what the kernel must at
least read or write once; FLOPs are counted for the convolutions (2 x MAC) only."""
import argparse
import collections
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))

import torch  # noqa: E402

from b200seg import _lib, engine as EN, model as M, raw  # noqa: E402

QUERY = re.compile(r"(_elems|_ws_bytes|_grid|_launches|_chunk|_splits|_abi_version|_build_info|_plan_info)$")
SIZES = {}          # fake address -> bytes
EVENTS = []         # (entry point, bytes, flops, detail)
_next = [0x7F0000000000]     # far above any scalar argument (pixel counts, pitches ...) the entry points take


def fake_ptr(t):
    if t is None:
        return None
    a = _next[0]
    _next[0] += 0x1000
    SIZES[a] = t.numel() * t.element_size()
    return a


class Recorder:
    def __init__(self, real):
        self.real = real

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if QUERY.search(name):
            return fn

        def fake(*args):
            nbytes, flops, detail = 0, 0.0, ""
            for a in args:
                if isinstance(a, int) and a in SIZES:
                    nbytes += SIZES[a]
                elif hasattr(a, "_obj"):                      # ctypes.byref(...)
                    obj = a._obj
                    if isinstance(obj, _lib.ConvDesc):
                        k, s = obj.ksize, obj.stride
                        dil = max(1, getattr(obj, "dilation", 1))
                        ho = (obj.h + 2 * obj.pad - dil * (k - 1) - 1) // s + 1
                        wo = (obj.w + 2 * obj.pad - dil * (k - 1) - 1) // s + 1
                        flops = 2.0 * obj.n * ho * wo * obj.cin * obj.cout * k * k
                        detail = "%dx%d c%d->%d k%d s%d" % (obj.h, obj.w, obj.cin, obj.cout, k, s)
                    elif isinstance(obj, ctypes.c_int32):
                        obj.value = 148                       # stats_grid of the convolutions
            EVENTS.append((name[len("b200seg_"):], nbytes, flops, detail))
            return 0
        return fake


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="ocrnet.HRNet_Mscale")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    pk = dict(bf16_tflops_sustained=1386.7, hbm_gbs=6572.9)       # this pool's measured values; refreshed from the file
    try:
        import json
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk.update(json.load(f))
    except (OSError, ValueError):
        pass
    ap.add_argument("--tflops", type=float, default=pk["bf16_tflops_sustained"],
                    help="sustained dense bf16 TFLOP/s (MEASURED_PEAKS.json)")
    ap.add_argument("--gbs", type=float, default=pk["hbm_gbs"], help="HBM copy bandwidth GB/s (MEASURED_PEAKS.json)")
    ap.add_argument("--top", type=int, default=0, help="also list the N largest launches by roofline time")
    ap.add_argument("--fused-bn", action="store_true",
                    help="trace the opt-in program with the BatchNorm finalisation inside the producing launches "
                         "(B200SEG_FUSED_BN=1) instead of the default bn_finalize / bn_bwd_finalize launches")
    ap.add_argument("--bn-cells", action="store_true",
                    help="trace the training default with per-GPU statistics (B200SEG_BN_CELLS=1): deferred BatchNorm "
                         "finalisation inside the consuming apply passes; without the flag: the program with separate "
                         "bn_finalize / bn_bwd_finalize launches (SyncBN, B200SEG_BN_CELLS=0)")
    args = ap.parse_args()
    if __debug__:
        sys.exit("run with python -O (the wrappers assert is_cuda)")

    rec = Recorder(_lib.lib())
    _lib.lib = lambda: rec
    raw.lib = lambda: rec
    raw.ptr = fake_ptr
    raw.stream_ptr = lambda: 0

    from b200seg.module import B200SegModule
    net = B200SegModule(args.arch, 19)
    dev = torch.device("meta")
    tensors = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in net._tensors().items()}
    packed, grads = {}, {}
    for n, p in net.named_parameters():
        if p.dim() == 4:
            o, i, k, _ = p.shape
            i = 16 if i == 3 else i
            packed[n[: -len(".weight")]] = (torch.empty((o, k * k, i), dtype=torch.bfloat16, device=dev),
                                            torch.empty((i, k * k, (o + 7) // 8 * 8), dtype=torch.bfloat16, device=dev))
            grads[n] = torch.empty((o, k * k, i), dtype=torch.float32, device=dev)
        else:
            grads[n] = torch.empty(p.shape, dtype=torch.float32, device=dev)
    images = torch.empty((1, 3, args.height, args.width), dtype=torch.float32, device=dev)
    gts = torch.empty((1, args.height, args.width), dtype=torch.long, device=dev)
    if args.arch.startswith("deepv3."):
        from b200seg import arch as A
        mask = torch.empty((sum(c for _b, c, _p in A.wrn_drop_layout(net.hcfg)),), dtype=torch.float32, device=dev)
    else:
        mask = torch.empty((1, net.ocfg["mid_channels"]), dtype=torch.float32, device=dev)
    bnfold = None
    if args.fused_bn:
        bnfold = {n[: -len(".running_mean")]: (torch.empty(2 * ((v.shape[0] + 15) // 16 * 16), dtype=torch.float64, device=dev),
                                               torch.empty(1, dtype=torch.int32, device=dev))
                  for n, v in tensors.items() if n.endswith(".running_mean")}
    bncells = None
    if args.bn_cells:
        bncells = {n[: -len(".running_mean")]: (torch.empty(2 * ((v.shape[0] + 15) // 16 * 16), dtype=torch.float64, device=dev),
                                                torch.empty(2 * v.shape[0], dtype=torch.float64, device=dev))
                   for n, v in tensors.items() if n.endswith(".running_mean")}
    E = EN.Engine(tensors, grads, packed, True, mask, bnfold=bnfold, bncells=bncells)
    M.train_loss(E, images, gts, args.arch, net.hcfg, net.ocfg)
    n_fwd = len(EVENTS)
    M.run_backward(E)

    P, B = args.tflops * 1e12, args.gbs * 1e9
    agg = collections.OrderedDict()
    tot = [0, 0.0, 0.0, 0.0]
    for i, (name, nbytes, flops, detail) in enumerate(EVENTS):
        t = max(flops / P, nbytes / B)
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        for acc in (a, tot):
            acc[0] += 1; acc[1] += flops; acc[2] += nbytes; acc[3] += t
    print("%s  %dx%d  one crop: %d launches (%d forward + loss, %d backward)" %
          (args.arch, args.height, args.width, len(EVENTS), n_fwd, len(EVENTS) - n_fwd))
    print("%-24s %8s %12s %12s %12s" % ("entry point", "launches", "TFLOP", "GB", "roofline ms"))
    for name, (c, f, b, t) in sorted(agg.items(), key=lambda kv: -kv[1][3]):
        print("%-24s %8d %12.3f %12.2f %12.3f" % (name, c, f / 1e12, b / 1e9, t * 1e3))
    print("%-24s %8d %12.3f %12.2f %12.3f" % ("total", tot[0], tot[1] / 1e12, tot[2] / 1e9, tot[3] * 1e3))
    print("(peaks: %.1f TFLOP/s, %.0f GB/s; roofline ms = sum over launches of max(FLOPs/P, bytes/B))" %
          (args.tflops, args.gbs))
    if args.top:
        big = sorted(EVENTS, key=lambda e: -max(e[2] / P, e[1] / B))[: args.top]
        for name, nbytes, flops, detail in big:
            print("  %-22s %-28s %8.1f MB %8.2f GFLOP %8.1f us" %
                  (name, detail, nbytes / 1e6, flops / 1e9, max(flops / P, nbytes / B) * 1e6))


if __name__ == "__main__":
    main()
