#!/usr/bin/env python
"""Diagnostic (GPU): per-tensor gradient agreement of the deepv3.DeepV3PlusW38 train step with the GPU-run oracle
(bf16-storage emulation) next to the oracle's own one-ulp noise floor. Test infrastructure (imports oracle/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import _parity as P  # noqa: E402
from oracle import seg_oracle as O  # noqa: E402
from b200seg.module import B200SegModule  # noqa: E402
from test_gpu_deepv3 import ARCH, WRN_TEST  # noqa: E402


def oracle_step(sd0, images, gts, emulate=True, dev="cuda"):
    P._tf32_off()
    sd = {k: v.clone().to(dev) for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    loss = O.deepv3_forward(O.Ctx(sd, training=True, emulate_bf16=emulate), images.to(dev), gts.to(dev), wcfg=WRN_TEST)
    loss.backward()
    return sd, float(loss)


def run_variant(tag, sd0, images, gts):
    sd_ref, l_ref = oracle_step(sd0, images, gts)
    sd_p, l_p = oracle_step(sd0, P.ulp_perturbed(images), gts)
    sd_32, l_32 = oracle_step(sd0, images, gts, emulate=False)
    net = B200SegModule(ARCH, 19, criterion=None, hcfg=WRN_TEST, use_cuda_graph=False)
    net.load_state_dict(sd0)
    net.wrn_dropout_scale = 0.0
    net = net.cuda().train()
    loss = net({"images": images.cuda(), "gts": gts.cuda()})
    loss.backward()
    torch.cuda.synchronize()
    print("[%s] loss product %.6f oracle(bf16) %.6f oracle(ulp) %.6f oracle(fp32) %.6f" % (tag, float(loss), l_ref, l_p, l_32))
    rows = []
    for name, p in net.named_parameters():
        g = sd_ref[name].grad
        if g is None:
            continue
        c, r = P.cos_rel(p.grad, g)
        cf, rf = P.cos_rel(sd_p[name].grad, g)
        c3, r3 = P.cos_rel(sd_32[name].grad, g)
        rows.append((name, r, c, rf, cf, r3, float(g.norm()), float(p.grad.norm())))
    med = lambda xs: sorted(xs)[len(xs) // 2]
    for part, sel in (("trunk", lambda n: n.startswith("backbone")), ("head", lambda n: not n.startswith("backbone"))):
        rr = [r for r in rows if sel(r[0])]
        print("[%s] %-5s n=%3d  median rel product %.4f floor %.4f fp32 %.4f | max rel product %.4f floor %.4f | median cos product %.4f floor %.4f"
              % (tag, part, len(rr), med([r[1] for r in rr]), med([r[3] for r in rr]), med([r[5] for r in rr]),
                 max(r[1] for r in rr), max(r[3] for r in rr), med([r[2] for r in rr]), med([r[4] for r in rr])))
    rep = {r[0]: (r[2], r[1], r[6]) for r in rows}
    floor = {r[0]: (r[4], r[3]) for r in rows}
    bad, summ = P.check_against_floor(rep, floor, {}, {})
    print("[%s] check_against_floor: %d violations %s %s" % (tag, len(bad), summ, bad[:6]))
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "deepv3_diag_%s.json" % tag), "w"))
    with open(os.path.join(ROOT, "gpurun_out", "deepv3_diag_%s.txt" % tag), "w") as f:
        for row in rows:
            f.write("%-52s %8.4f %8.4f | %8.4f %8.4f | %8.4f | %10.3e %10.3e\n" % row)


def main():
    sd0 = O.synth_state_dict(ARCH, WRN_TEST, seed=3)
    images, gts = O.synth_batch(2, 128, 256, seed=5)
    run_variant("base", sd0, images, gts)
    sd1 = {k: v.clone() for k, v in sd0.items()}
    sd1["aspp.img_conv.1.weight"].zero_()          # the image-pooling branch feeds no gradient to the trunk
    run_variant("imgzero", sd1, images, gts)
    images4, gts4 = O.synth_batch(4, 96, 192, seed=6)
    run_variant("n4", sd0, images4, gts4)
    run_variant("n4_imgzero", sd1, images4, gts4)


if __name__ == "__main__":
    main()
