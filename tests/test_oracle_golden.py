"""Pins oracle/seg_oracle.py against outputs of the UNMODIFIED reference (tests/golden/reference_outputs.pt, produced
by tests/golden/make_golden.py in the build container). CPU only; runs here and on the GPU box."""
import pytest
import torch

from oracle import seg_oracle as O

torch.set_num_threads(8)

RTOL = 2e-4   # fp32 reductions in a different association order (the reference itself varies 1e-4 with thread count)


def close(a, b, rtol=RTOL, atol=None):
    a, b = a.double(), b.double()
    scale = b.abs().max().item() + 1e-30
    err = (a - b).abs().max().item()
    assert err <= rtol * scale + (atol or 0.0), "max err %.3e vs scale %.3e" % (err, scale)


def run_train(arch, loss_name, hcfg, sd_seed, batch, res):
    sd = O.synth_state_dict(arch, hcfg, seed=sd_seed)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    ctx = O.Ctx(sd, training=True)
    images, gts = O.synth_batch(batch, res[0], res[1], seed=5 if hcfg is O.HRNET_W16_TEST else 7)
    crit = O.criterion_ce if loss_name == "ce" else O.criterion_rmi
    if arch == "ocrnet.HRNet_Mscale":
        loss = O.mscale_two_scale(ctx, images, gts, criterion=crit, hcfg=hcfg)
    elif arch == "ocrnet.HRNet":
        loss = O.ocrnet_forward(ctx, images, gts, criterion=crit, hcfg=hcfg)
    else:
        loss = O.basic_forward(ctx, images, gts, criterion=crit, hcfg=hcfg)
    loss.backward()
    return sd, loss, images


@pytest.mark.parametrize("arch,loss_name", [("ocrnet.HRNet_Mscale", "ce"), ("ocrnet.HRNet_Mscale", "rmi"),
                                            ("ocrnet.HRNet", "ce"), ("ocrnet.HRNet", "rmi"), ("basic.HRNet", "ce")])
def test_w16_train_step_matches_reference(golden, arch, loss_name):
    g = golden["%s/%s/w16" % (arch, loss_name)]
    sd, loss, images = run_train(arch, loss_name, O.HRNET_W16_TEST, 3, 2, (64, 128))
    close(loss.detach(), g["loss"], rtol=1e-5)
    for name, ref in g["grads"].items():
        close(O.sample_like(sd[name].grad), ref, rtol=2e-3 if loss_name == "rmi" else 5e-4)
    close(sd["backbone.bn1.running_mean"].detach(), g["running_mean_bb_bn1"], atol=1e-6)
    close(sd["backbone.stage4.0.branches.3.0.bn2.running_var"].detach(), g["running_var_s4"])
    if loss_name != "ce":
        return
    # eval mode continues from the state the training step left behind (running stats updated), like the reference run
    sd = {k: v.detach() for k, v in sd.items()}
    ctx = O.Ctx(sd, training=False)
    with torch.no_grad():
        if arch == "ocrnet.HRNet_Mscale":
            o2 = O.mscale_two_scale(ctx, images, hcfg=O.HRNET_W16_TEST)
            for k, ref in g["eval_two_scale"].items():
                close(O.sample_like(o2[k]), ref)
            o3 = O.mscale_nscale(ctx, images, [0.5, 1.0, 2.0], hcfg=O.HRNET_W16_TEST)
            assert sorted(o3.keys()) == g["eval_three_scale_keys"]
            for k, ref in g["eval_three_scale"].items():
                close(O.sample_like(o3[k]), ref)
        elif arch == "ocrnet.HRNet":
            o = O.ocrnet_forward(ctx, images, hcfg=O.HRNET_W16_TEST)
            close(O.sample_like(o["pred"]), g["eval"]["pred"])
        else:
            o = O.basic_forward(ctx, images, hcfg=O.HRNET_W16_TEST)
            close(O.sample_like(o["pred"]), g["eval"]["pred"])


def test_loss_heads_match_reference(golden):
    g = golden["loss_heads"]
    gen = torch.Generator().manual_seed(11)
    logits = torch.randn((2, 19, 48, 80), generator=gen, requires_grad=True)
    _, gts = O.synth_batch(2, 48, 80, seed=12)
    ce = O.ce_loss(logits, gts)
    (gce,) = torch.autograd.grad(ce, logits)
    close(ce.detach(), g["ce"], rtol=1e-6)
    close(O.sample_like(gce), g["ce_grad"], rtol=1e-5)
    rmi = O.rmi_loss(logits, gts, do_rmi=True)
    (grmi,) = torch.autograd.grad(rmi, logits)
    close(rmi.detach(), g["rmi"], rtol=1e-5)
    close(O.sample_like(grmi), g["rmi_grad"], rtol=1e-3)
    close(O.rmi_loss(logits, gts, do_rmi=False).detach(), g["bce"], rtol=1e-6)


def test_w48_state_dict_and_step_match_reference(golden):
    g = golden["ocrnet.HRNet_Mscale/ce/w48"]
    sd, loss, images = run_train("ocrnet.HRNet_Mscale", "ce", O.HRNET_W48, 0, 1, (64, 128))
    assert len(sd) == g["nkeys"] == 1903                       # 1903 state-dict keys
    nparams = sum(v.numel() for k, v in sd.items() if v.requires_grad)
    assert nparams == g["nparams"] == 72143430                 # 72.14 M parameters
    close(loss.detach(), g["loss"], rtol=1e-5)
    close(O.sample_like(sd["backbone.conv1.weight"].grad), g["grad_conv1"], rtol=2e-3)
    close(O.sample_like(sd["ocr.cls_head.weight"].grad), g["grad_cls"], rtol=5e-4)
    close(O.sample_like(sd["scale_attn.conv2.weight"].grad), g["grad_attn"], rtol=5e-4)
    sd = {k: v.detach() for k, v in sd.items()}
    with torch.no_grad():
        o2 = O.mscale_two_scale(O.Ctx(sd, training=False), images)
    close(O.sample_like(o2["pred"]), g["eval_pred_sample"])
    close(O.sample_like(o2["attn_05x"]), g["eval_attn_sample"])


def test_deepv3_wrn38_matches_reference():
    """SURVEY §8(f) row f2 (BASELINE config 4): the DeepLabV3+ / WideResNet-38 restatement against the unmodified
    reference (tests/golden/make_golden_deepv3.py): state_dict registration order, train loss, sampled gradients,
    running statistics, eval prediction."""
    import os
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_deepv3.pt"),
                   map_location="cpu")
    arch = "deepv3.DeepV3PlusW38"
    sd = O.synth_state_dict(arch, O.WRN38, seed=4)
    assert list(sd.keys()) == g["keys"] and len(g["keys"]) == 268
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    assert [k for k, v in sd.items() if v.requires_grad] == g["param_order"]
    assert sum(v.numel() for v in sd.values() if v.requires_grad) == g["nparams"] == 137103936
    images, gts = O.synth_batch(2, 64, 128, seed=6)
    loss = O.deepv3_forward(O.Ctx(sd, training=True), images, gts)
    loss.backward()
    close(loss.detach(), g["loss"], rtol=1e-5)
    for name, ref in g["grads"].items():
        close(O.sample_like(sd[name].grad), ref, rtol=1e-3)
    close(sd["backbone.mod5.block3.convs.bn2.0.running_var"].detach(), g["running_var_mod5"])
    sd = {k: v.detach() for k, v in sd.items()}
    with torch.no_grad():
        o = O.deepv3_forward(O.Ctx(sd, training=False), images)
    close(O.sample_like(o["pred"]), g["eval_pred"])


# ----------------------------------------------------------------------------------------------- mscale.HRNet (MscaleBasic)
@pytest.fixture(scope="module")
def golden_mscale():
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return torch.load(os.path.join(here, "golden", "reference_mscale.pt"), map_location="cpu")


@pytest.mark.parametrize("loss_name,sup", [("ce", 0.0), ("ce", 0.05), ("rmi", 0.05)])
def test_mscale_basic_train_step_matches_reference(golden_mscale, loss_name, sup):
    """arch 'mscale.HRNet' (network/mscale.py:450-475, two_scale_forward :182-220) against the imported reference."""
    g = golden_mscale["mscale.HRNet/%s/sup%g/w16" % (loss_name, sup)]
    hcfg = O.HRNET_W16_TEST
    sd = O.synth_state_dict("mscale.HRNet", hcfg, seed=3)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    images, gts = O.synth_batch(2, 64, 128, seed=5)
    crit = O.criterion_ce if loss_name == "ce" else O.criterion_rmi
    loss = O.mscale_basic_two_scale(O.Ctx(sd, training=True), images, gts, criterion=crit, hcfg=hcfg,
                                    supervised_mscale_wt=sup)
    loss.backward()
    close(loss.detach(), g["loss"], rtol=1e-5)
    for name, ref in g["grads"].items():
        close(O.sample_like(sd[name].grad), ref, rtol=2e-3 if loss_name == "rmi" else 5e-4)
    close(sd["backbone.bn1.running_mean"].detach(), g["running_mean_bb_bn1"], atol=1e-6)
    close(sd["scale_attn.bn1.running_var"].detach(), g["running_var_attn"])
    if "eval_two_scale" not in g:
        return
    sd = {k: v.detach() for k, v in sd.items()}
    ctx = O.Ctx(sd, training=False)
    with torch.no_grad():
        o2 = O.mscale_basic_two_scale(ctx, images, hcfg=hcfg)
        for k, ref in g["eval_two_scale"].items():
            close(O.sample_like(o2[k]), ref)
        o3 = O.mscale_basic_nscale(ctx, images, [0.5, 1.0, 2.0], hcfg=hcfg)
        assert sorted(o3.keys()) == g["eval_three_scale_keys"]
        for k, ref in g["eval_three_scale"].items():
            close(O.sample_like(o3[k]), ref)


def test_mscale_basic_state_dict_order_matches_reference(golden_mscale):
    """The B200 module registers the tensors of mscale.HRNet under the reference's names, in its order."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "semantic-segmentation_b200"))
    from b200seg import arch as A
    names = [n for n, _s, _k in A.tensor_specs("mscale.HRNet")]
    assert names == golden_mscale["state_dict_keys_w48"]
    nparams = 0
    for _n, shp, kind in A.tensor_specs("mscale.HRNet"):
        if kind in ("conv_w", "conv_b", "bn_w", "bn_b"):
            k = 1
            for s in shp:
                k *= s
            nparams += k
    assert nparams == golden_mscale["nparams_w48"]
