"""GPU bring-up check for the tcgen05 implicit-GEMM convolution (run under gpurun; prints one line per case)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
import torch.nn.functional as F

from b200seg import raw

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run_case(n, h, w, cin, cout, k, s, bias, kc, fp32out=False, stats=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((n, h, w, cin), generator=g, device="cuda").to(torch.bfloat16)
    wt = (torch.randn((cout, cin, k, k), generator=g, device="cuda") / (cin * k * k) ** 0.5).contiguous()
    b = torch.randn((cout,), generator=g, device="cuda") if bias else None
    wf, wd = raw.pack_weight(wt)
    wr = wt.to(torch.bfloat16).float()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wr, b, stride=s, padding=1 if k == 3 else 0).permute(0, 2, 3, 1)
    res = raw.conv2d_fwd(x, wf, b, stride=s, out_fp32=fp32out, emit_stats=stats, force_kc=kc)
    torch.cuda.synchronize()
    st = None
    if stats:
        res, st = res
    y = res.float()
    err = (y - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-5 if fp32out else 1.0 / 128) * scale + 1e-6
    ok = err <= tol
    msg = "n%d %dx%d cin%d cout%d k%d s%d bias%d kc%d fp32%d: maxerr %.3e scale %.3e %s" % (
        n, h, w, cin, cout, k, s, int(bias), kc, int(fp32out), err, scale, "OK" if ok else "FAIL")
    if stats:
        tot = st.sum(0)[:, :cout]
        yr = res.float().reshape(-1, cout)
        e1 = (tot[0] - yr.sum(0)).abs().max().item() / max(1.0, yr.sum(0).abs().max().item())
        e2 = (tot[1] - (yr * yr).sum(0)).abs().max().item() / (yr * yr).sum(0).abs().max().item()
        ok2 = e1 < 1e-3 and e2 < 1e-3
        msg += " stats rel %.2e %.2e %s" % (e1, e2, "OK" if ok2 else "FAIL")
        ok = ok and ok2
    if not ok:
        d = (y - ref).abs()
        idx = torch.nonzero(d > tol)
        msg += " nbad %d first %s" % (idx.shape[0], idx[:4].tolist())
    print(msg, flush=True)
    return ok


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    cases = [
        # n, h, w, cin, cout, k, s, bias, kc
        (1, 8, 16, 64, 64, 1, 1, False, 0),
        (1, 8, 16, 64, 64, 3, 1, False, 0),
        (1, 16, 32, 64, 64, 3, 1, False, 0),
        (2, 24, 40, 128, 96, 3, 1, True, 0),
        (1, 16, 32, 32, 32, 3, 1, False, 32),
        (1, 16, 32, 48, 48, 3, 1, False, 16),
        (1, 16, 32, 48, 48, 3, 1, False, 64),
        (1, 32, 64, 96, 96, 3, 1, False, 0),
        (1, 32, 64, 192, 192, 3, 1, False, 0),
        (1, 16, 32, 384, 384, 3, 1, False, 0),
        (1, 32, 64, 720, 512, 3, 1, True, 0),
        (1, 32, 64, 720, 720, 1, 1, True, 0),
        (1, 32, 64, 512, 256, 3, 1, False, 0),
        (1, 32, 64, 64, 64, 3, 2, False, 0),
        (1, 32, 64, 48, 96, 3, 2, False, 0),
        (1, 30, 52, 96, 192, 3, 2, False, 0),
        (1, 19, 1, 512, 256, 1, 1, False, 0),
    ]
    allok = True
    for c in cases:
        try:
            allok &= run_case(*c)
        except Exception as e:  # noqa
            print("case %s raised %r" % (c, e), flush=True)
            allok = False
            torch.cuda.synchronize()
    # fp32-output logit head and stats
    allok &= run_case(1, 32, 64, 512, 19, 1, 1, True, 0, fp32out=True)
    allok &= run_case(2, 32, 64, 48, 48, 3, 1, False, 0, stats=True)
    allok &= run_case(1, 40, 72, 256, 512, 1, 1, False, 0, stats=True)
    allok &= run_case(1, 32, 64, 720, 720, 1, 1, True, 0, stats=True)
    # timing on a full-size layer: 48ch 256x512 (HBM-bound) and 720->512 (tensor-bound)
    for (h, w, cin, cout, k, kc) in [(256, 512, 48, 48, 3, 16), (256, 512, 48, 48, 3, 64), (128, 256, 96, 96, 3, 0),
                                     (64, 128, 192, 192, 3, 0), (32, 64, 384, 384, 3, 0),
                                     (256, 512, 720, 512, 3, 0), (256, 512, 512, 256, 3, 0),
                                     (256, 512, 512, 256, 1, 0)]:
        x = torch.randn((1, h, w, cin), device="cuda").to(torch.bfloat16)
        wt = torch.randn((cout, cin, k, k), device="cuda") * 0.05
        wf, _ = raw.pack_weight(wt)
        for _ in range(3):
            raw.conv2d_fwd(x, wf, force_kc=kc, emit_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 10
        for _ in range(iters):
            raw.conv2d_fwd(x, wf, force_kc=kc, emit_stats=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * h * w * cin * cout * k * k
        byts = 2.0 * h * w * (cin + cout)
        print("time %dx%d cin%d cout%d k%d kc%d: %.3f ms  %.1f TFLOP/s  %.1f GB/s(alg)" % (
            h, w, cin, cout, k, kc, ms, flops / ms / 1e9, byts / ms / 1e6), flush=True)
    print("ALL OK" if allok else "SOME FAILED", flush=True)


if __name__ == "__main__":
    main()
