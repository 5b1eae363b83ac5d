"""Runs two eager training steps with every C-ABI call logged (entry point, convolution shape, FLOPs, bytes of the tensors
handed over, kernels launched) and writes gpurun_out/abi_calls.json. Run it UNDER ncu to get the matching launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/abi_launches.csv \
        python tools/gpu_abi_log.py [H W]
    python -O tools/conv_classes.py gpurun_out/abi_launches.csv --abi gpurun_out/abi_calls.json      # here, on the CPU

The k-th b200seg kernel of the launch list belongs to the ABI call whose cumulative launch count covers k, so measured
durations can be attributed to convolution classes exactly (no sequence guessing)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))

import torch  # noqa: E402

from b200seg import _lib, raw  # noqa: E402

CALLS = []       # [entry point, detail, flops, bytes, launches]
SIZES = {}
REAL = _lib.lib()
real_ptr, real_check = raw.ptr, raw.check


def ptr(t):
    if t is not None:
        SIZES[t.data_ptr()] = t.numel() * t.element_size()
    return real_ptr(t)


class Log:
    def __getattr__(self, name):
        fn = getattr(REAL, name)

        def call(*args):
            nbytes, flops, detail = 0, 0.0, ""
            for a in args:
                if isinstance(a, int) and a in SIZES:
                    nbytes += SIZES[a]
                elif hasattr(a, "_obj") and isinstance(a._obj, _lib.ConvDesc):
                    d = a._obj
                    k, s = d.ksize, d.stride
                    ho, wo = (d.h + 2 * d.pad - k) // s + 1, (d.w + 2 * d.pad - k) // s + 1
                    flops = 2.0 * d.n * ho * wo * d.cin * d.cout * k * k
                    detail = "%dx%d c%d->%d k%d s%d" % (d.h, d.w, d.cin, d.cout, k, s)
            rc = fn(*args)
            CALLS.append([name[len("b200seg_"):], detail, flops, nbytes, 0])
            return rc
        return call


def check(rc, what, launches=1):
    if CALLS:
        CALLS[-1][4] += launches
    return real_check(rc, what, launches)


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 2048)
    log = Log()
    _lib.lib = lambda: log
    raw.lib = lambda: log
    raw.ptr = ptr
    raw.check = check
    from b200seg import optim
    optim.lib, optim.check, optim.ptr = (lambda: log), check, ptr
    from b200seg.module import B200SegModule
    torch.manual_seed(0)
    net = B200SegModule("ocrnet.HRNet_Mscale", 19, use_cuda_graph=False).cuda().train()
    with torch.no_grad():
        for n_, p_ in net.named_parameters():
            if p_.dim() == 4 and n_.startswith("backbone"):
                p_.normal_(0, (2.0 / (p_.shape[1] * p_.shape[2] * p_.shape[3])) ** 0.5)
    opt = optim.FusedSGD(net.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    images = torch.randn(1, 3, H, W, device="cuda")
    gts = torch.randint(0, 19, (1, H, W), device="cuda")
    marks = []
    for _ in range(2):
        marks.append(len(CALLS))
        opt.zero_grad(set_to_none=True)
        loss = net({"images": images, "gts": gts})
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "abi_calls.json"), "w") as f:
        json.dump(dict(size=[H, W], step_starts=marks, calls=CALLS), f)
    print("logged %d ABI calls, %d kernels; loss %.4f" % (len(CALLS), sum(c[4] for c in CALLS), float(loss)))


if __name__ == "__main__":
    main()
