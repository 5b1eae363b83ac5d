#!/usr/bin/env python
"""Diagnostic (GPU): per-tensor gradient agreement of the deepv3.DeepV3PlusW38 train step with the GPU-run oracle
(bf16-storage emulation) next to the oracle's own one-ulp noise floor. Test infrastructure (imports oracle/)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


def main():
    sd0 = O.synth_state_dict(ARCH, WRN_TEST, seed=3)
    images, gts = O.synth_batch(2, 128, 256, seed=5)
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
    print("loss product %.6f oracle(bf16) %.6f oracle(ulp) %.6f oracle(fp32) %.6f" % (float(loss), l_ref, l_p, l_32))
    rows = []
    for name, p in net.named_parameters():
        g = sd_ref[name].grad
        if g is None:
            continue
        c, r = P.cos_rel(p.grad, g)
        cf, rf = P.cos_rel(sd_p[name].grad, g)
        c3, r3 = P.cos_rel(sd_32[name].grad, g)
        rows.append((name, r, c, rf, cf, r3, float(g.norm()), float(p.grad.norm())))
    print("%-52s %8s %8s | %8s %8s | %8s | %10s %10s" % ("tensor", "rel", "cos", "floor", "cosfl", "rel fp32", "|ref|", "|prod|"))
    for row in rows:
        print("%-52s %8.4f %8.4f | %8.4f %8.4f | %8.4f | %10.3e %10.3e" % row)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "deepv3_diag.json"), "w"))


if __name__ == "__main__":
    main()
