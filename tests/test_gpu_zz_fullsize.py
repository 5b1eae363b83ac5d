"""BASELINE.json's full configuration (HRNet-W48 + OCR + two-scale attention, one 1024x2048 crop) through properties that
need no CPU restatement at that size (the oracle takes minutes per step there; it is compared at reduced width / size in
test_gpu_model.py):

  * idempotence: the captured step replayed twice on the same batch gives the same loss and the same flat gradient, and
    the eager warm-up step agrees with the replays;
  * loss bookkeeping: total = main + OCR_ALPHA * aux (loss/utils + network/ocrnet.py:303-318);
  * softmax cross-entropy is shift invariant per pixel, so the gradient of every CE logit head sums to zero over the
    classes - for the bias (sum over pixels of sum_c dlogit_c) and for every input channel of the weight. This goes
    through the attention blend, both bilinear upsampling adjoints, the bf16 logit gradients, the bias column sum and
    the weight-gradient GEMM at full size;
  * BatchNorm bookkeeping: every layer saw two forward passes per step (0.5x and 1.0x), statistics finite and positive;
  * every parameter receives a finite gradient.
Runs last (file name) because it holds ~40 GB of activations and graph pools while it runs."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
def test_full_size_two_scale_step_properties():
    from b200seg.module import B200SegModule
    torch.manual_seed(0)
    net = B200SegModule("ocrnet.HRNet_Mscale", 19, ocfg=None).cuda().train()
    with torch.no_grad():       # He initialisation of the trunk (the reference's N(0, 1e-3) underflows without a checkpoint)
        for n_, p_ in net.named_parameters():
            if p_.dim() == 4 and n_.startswith("backbone"):
                p_.normal_(0, (2.0 / (p_.shape[1] * p_.shape[2] * p_.shape[3])) ** 0.5)
    g = torch.Generator().manual_seed(11)
    images = torch.randn((1, 3, 1024, 2048), generator=g).cuda()
    gts = torch.randint(0, 19, (1, 1024, 2048), generator=g)
    gts[:, :8] = 255
    gts = gts.cuda()
    drop = net.ocfg["dropout"]
    net.ocfg["dropout"] = 0.0          # identical steps: no fresh Dropout2d mask per call

    def step():
        net.zero_grad(set_to_none=True)
        loss = net({"images": images, "gts": gts})
        loss.backward()
        torch.cuda.synchronize()
        return float(loss), net._flat_grad.clone(), net.last_loss_terms.clone()

    l_eager, g_eager, _ = step()               # eager warm-up
    l_a, g_a, terms = step()                   # capture + first replay
    l_b, g_b, _ = step()                       # second replay
    net.ocfg["dropout"] = drop
    assert all(map(lambda v: v == v and abs(v) < 1e4, (l_eager, l_a, l_b))), (l_eager, l_a, l_b)

    def rel(a, b):
        return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

    assert abs(l_a - l_b) <= 1e-6 * abs(l_a), (l_a, l_b)
    assert rel(g_a, g_b) <= 1e-5
    assert abs(l_eager - l_a) <= 1e-4 * abs(l_a), (l_eager, l_a)
    assert rel(g_eager, g_a) <= 1e-3

    total, main, aux = (float(terms[i]) for i in range(3))
    assert abs(total - (main + net.ocr_alpha * aux)) <= 1e-5 * abs(total), (total, main, aux)
    assert abs(total - l_a) <= 1e-6 * abs(total)

    for head in ("ocr.cls_head", "ocr.aux_head.2"):
        gb = net.get_parameter(head + ".bias").grad.double()
        gw = net.get_parameter(head + ".weight").grad.double().flatten(1)        # [19, Cin]
        assert gb.abs().sum() > 0 and gw.abs().sum() > 0, head
        assert abs(float(gb.sum())) <= 2e-3 * float(gb.abs().sum()), (head, gb)
        col = gw.sum(0).abs().sum() / gw.abs().sum()
        assert float(col) <= 2e-3, (head, float(col))

    n_nonfinite = 0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        n_nonfinite += int((~torch.isfinite(p.grad)).sum())
    assert n_nonfinite == 0
    assert float(g_a.norm()) > 0

    sd = net.state_dict()
    nbt = {int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")}
    assert nbt == {6}, nbt                      # 3 steps x 2 scale passes (the reference's BN modules run twice per step)
    for k, v in sd.items():
        if k.endswith("running_var"):
            assert bool(torch.isfinite(v).all()) and float(v.min()) > 0, k
        elif k.endswith("running_mean"):
            assert bool(torch.isfinite(v).all()), k
