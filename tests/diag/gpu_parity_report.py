#!/usr/bin/env python
"""Whole-step parity figures of the B200 path against the GPU-run oracle (fp32 cuDNN, TF32 off, bf16-storage
emulation) at the test width and at the BASELINE configuration (HRNet-W48, 1024x2048, one crop).

    python tools/gpu_parity_report.py [out.json]        # on a B200 box

Diagnostic: prints the numbers the thresholds of tests/test_gpu_model.py / test_gpu_zz_fullsize.py are set from."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import _parity as P  # noqa: E402
from oracle import seg_oracle as O  # noqa: E402
from b200seg.module import B200SegModule  # noqa: E402


def summarize(rep):
    cs = sorted(v[0] for v in rep.values())
    rs = sorted(v[1] for v in rep.values())
    n = len(cs)
    worst = sorted(rep.items(), key=lambda kv: kv[1][0])[:8]
    return dict(n=n, cos_min=cs[0], cos_p05=cs[n // 20], cos_med=cs[n // 2], rel_max=rs[-1], rel_p95=rs[-1 - n // 20],
                rel_med=rs[n // 2], worst=[(k, round(v[0], 5), round(v[1], 5)) for k, v in worst])


def train_case(tag, arch, hcfg, n, h, w, sup=0.0, crit=None, emulate=True):
    sd0 = O.synth_state_dict(arch, hcfg, seed=3)
    images, gts = O.synth_batch(n, h, w, seed=5)
    t0 = time.time()
    sd_ref, l_ref = P.oracle_train_step(O, arch, hcfg, sd0, images, gts, sup, crit=O.criterion_rmi if crit == "rmi" else None,
                                        emulate_bf16=emulate)
    torch.cuda.synchronize()
    t1 = time.time()
    net, l = P.product_train_step(B200SegModule, O, arch, hcfg, sd0, images, gts, sup, criterion=crit)
    rep = P.grad_report(net, sd_ref)
    run = P.running_report(net, sd_ref)
    out = dict(tag=tag, loss=l, loss_ref=l_ref, loss_rel=abs(l - l_ref) / abs(l_ref), grads=summarize(rep),
               running_max=max(run.values()), oracle_s=t1 - t0,
               mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    # a spread of named tensors over the depth of the network
    names = [k for k in rep]
    pick = names[:: max(1, len(names) // 24)]
    out["sampled"] = {k: (round(rep[k][0], 5), round(rep[k][1], 5)) for k in pick}
    print(json.dumps(out), flush=True)
    del net, sd_ref
    torch.cuda.empty_cache()
    return out


def eval_case(tag, arch, hcfg, n, h, w, n_scales=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_model import _condition_eval_weights
    P._tf32_off()
    sd0 = O.synth_state_dict(arch, hcfg, seed=3)
    g = torch.Generator().manual_seed(9)
    for k in sd0:
        if k.endswith("running_mean"):
            sd0[k] = 0.1 * torch.randn(sd0[k].shape, generator=g)
        elif k.endswith("running_var"):
            sd0[k] = 0.5 + torch.rand(sd0[k].shape, generator=g)
    _condition_eval_weights(sd0)
    images, _ = O.synth_batch(n, h, w, seed=5)
    sd = {k: v.clone().cuda() for k, v in sd0.items()}
    out = {}
    net = B200SegModule(arch, 19, hcfg=hcfg, n_scales=n_scales)
    net.load_state_dict(sd0)
    net = net.cuda().eval()
    got = net({"images": images.cuda()})
    for emu in (True, False):
        ctx = O.Ctx(sd, training=False, emulate_bf16=emu)
        with torch.no_grad():
            ref = O.mscale_nscale(ctx, images.cuda(), n_scales, hcfg=hcfg) if n_scales else \
                O.mscale_two_scale(ctx, images.cuda(), hcfg=hcfg)
        r = {}
        for k in ref:
            c, rel = P.cos_rel(got[k], ref[k])
            mx = float((got[k] - ref[k]).abs().max() / ref[k].abs().max())
            r[k] = (round(rel, 6), round(mx, 6))
        a, b = got["pred"], ref["pred"]
        agree = float((a.argmax(1) == b.argmax(1)).float().mean())
        top2 = b.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1])
        thr = 0.02 * float(b.abs().max())
        conf = margin > thr
        agree_conf = float(((a.argmax(1) == b.argmax(1)) | ~conf).float().mean())
        out["emulated" if emu else "fp32"] = dict(maps=r, argmax_agree=agree, argmax_agree_margin=agree_conf,
                                                  frac_confident=float(conf.float().mean()))
        del ref
    out["tag"] = tag
    print(json.dumps(out), flush=True)
    del net, got
    torch.cuda.empty_cache()
    return out


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.json")
    res = []
    A = "ocrnet.HRNet_Mscale"
    res.append(train_case("w16_64x128_emul", A, O.HRNET_W16_TEST, 2, 64, 128))
    res.append(train_case("w16_64x128_fp32", A, O.HRNET_W16_TEST, 2, 64, 128, emulate=False))
    res.append(train_case("w16_64x128_sup_emul", A, O.HRNET_W16_TEST, 2, 64, 128, sup=0.05))
    res.append(train_case("w16_ocrnet_emul", "ocrnet.HRNet", O.HRNET_W16_TEST, 2, 64, 128))
    res.append(train_case("w16_basic_emul", "basic.HRNet", O.HRNET_W16_TEST, 2, 64, 128))
    res.append(train_case("w48_256x512_emul", A, O.HRNET_W48, 1, 256, 512))
    res.append(train_case("w48_1024x2048_emul", A, O.HRNET_W48, 1, 1024, 2048))
    res.append(eval_case("eval_w48_1024x2048_two_scale", A, O.HRNET_W48, 1, 1024, 2048))
    res.append(eval_case("eval_w16_two_scale", A, O.HRNET_W16_TEST, 2, 64, 128))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
