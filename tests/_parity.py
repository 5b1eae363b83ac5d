"""Shared helpers of the whole-step parity tests: run ONE training step of the B200 module and of the oracle (on the
GPU through stock PyTorch fp32, TF32 off, bf16-storage emulation) on identical weights and inputs, and report
matched-precision agreement figures (loss, per-tensor gradient cosine / relative L2, running statistics, eval maps).

Test infrastructure only (imports oracle/); used by tests/test_gpu_model.py, tests/test_gpu_zz_fullsize.py and
tests/diag/gpu_parity_report.py."""
import torch


def _tf32_off():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False


def oracle_train_step(O, arch, hcfg, sd0, images, gts, sup_wt=0.0, crit=None, emulate_bf16=True, device="cuda"):
    """-> (sd with .grad on every parameter and updated running statistics, loss float)."""
    _tf32_off()
    sd = {k: v.clone().to(device) for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    ctx = O.Ctx(sd, training=True, emulate_bf16=emulate_bf16)
    crit = crit or O.criterion_ce
    images, gts = images.to(device), gts.to(device)
    if arch == "mscale.HRNet":
        loss = O.mscale_basic_two_scale(ctx, images, gts, criterion=crit, hcfg=hcfg, supervised_mscale_wt=sup_wt)
    elif arch == "ocrnet.HRNet_Mscale":
        loss = O.mscale_two_scale(ctx, images, gts, criterion=crit, hcfg=hcfg, supervised_mscale_wt=sup_wt)
    elif arch == "ocrnet.HRNet":
        loss = O.ocrnet_forward(ctx, images, gts, criterion=crit, hcfg=hcfg)
    elif arch == "deepv3.DeepV3PlusW38":
        loss = O.deepv3_forward(ctx, images, gts, criterion=crit, wcfg=hcfg)
    else:
        loss = O.basic_forward(ctx, images, gts, criterion=crit, hcfg=hcfg)
    loss.backward()
    return sd, float(loss)


def product_train_step(B200SegModule, O, arch, hcfg, sd0, images, gts, sup_wt=0.0, criterion=None, graph=False):
    ocfg = dict(O.OCR_CFG)
    ocfg["dropout"] = 0.0
    net = B200SegModule(arch, 19, criterion=criterion, hcfg=hcfg, ocfg=ocfg, supervised_mscale_wt=sup_wt,
                        use_cuda_graph=graph)
    net.load_state_dict(sd0)
    net = net.cuda().train()
    loss = net({"images": images.cuda(), "gts": gts.cuda()})
    loss.backward()
    torch.cuda.synchronize()
    return net, float(loss)


def cos_rel(a, b):
    a, b = a.double().flatten(), b.double().flatten().to(a.device)
    na, nb = float(a.norm()), float(b.norm())
    if na == 0.0 or nb == 0.0:
        return (1.0 if na == nb else 0.0), (0.0 if na == nb else 1.0)
    return float((a * b).sum() / (na * nb)), float((a - b).norm() / nb)


def grad_report(net, sd_ref):
    """name -> (cosine, relative L2, |ref|) over every parameter whose oracle gradient is non-zero."""
    rep = {}
    for name, p in net.named_parameters():
        g_ref = sd_ref[name].grad
        if g_ref is None or float(g_ref.abs().max()) < 1e-12:
            continue
        c, r = cos_rel(p.grad, g_ref)
        rep[name] = (c, r, float(g_ref.norm()))
    return rep


def running_report(net, sd_ref):
    rep = {}
    sd = net.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            ref = sd_ref[k].to(v.device)
            rep[k] = float((v - ref).abs().max() / (ref.abs().max() + 1e-6))
    return rep


# ----------------------------------------------------------------------------------------------- noise floor
# Whole-network gradients of a batch-statistics-BN + ReLU network are not a smooth function of the arithmetic: a bf16
# rounding that falls the other way flips ReLU gates downstream, and every flipped gate changes the back-propagated
# signal of its unit by O(1). Measured on the oracle itself (fp64, CPU): an input perturbation of relative size eps moves
# per-tensor gradients by ~sqrt(0.4 eps) per layer, i.e. 30-100 % for eps = one bf16 ulp over the ~100-300 layers of the
# model - for ANY two implementations that round differently (the reference's own autocast run included).
# The end-to-end gradient criterion is therefore relative to that floor: the B200 path must be as close to the oracle as
# the oracle is to itself when its input moves by one bf16 ulp (op- and block-level tests carry the tight tolerances).
def ulp_perturbed(images, seed=99):
    g = torch.Generator().manual_seed(seed)
    return images * (1.0 + 2.0 ** -9 * torch.randn(images.shape, generator=g))


def noise_floor(O, arch, hcfg, sd0, images, gts, sd_ref, sup_wt=0.0, crit=None):
    """-> ({name: (cos, rel)} of oracle(images) vs oracle(ulp-perturbed images), {running stat: rel})."""
    sd_p, _ = oracle_train_step(O, arch, hcfg, sd0, ulp_perturbed(images), gts, sup_wt, crit)
    floor = {}
    for name, v in sd_ref.items():
        if v.grad is not None and float(v.grad.abs().max()) >= 1e-12:
            floor[name] = cos_rel(sd_p[name].grad, v.grad)
    run = {k: float((sd_p[k] - v).abs().max() / (v.abs().max() + 1e-6)) for k, v in sd_ref.items()
           if k.endswith("running_mean") or k.endswith("running_var")}
    return floor, run


def check_against_floor(rep, floor, run_rep, run_floor, slack=2.0, abs_slack=0.1, med_slack=1.25, outliers=0.02):
    """rep / floor: name -> (cos, rel[, norm]). Criteria: the median relative distance to the oracle stays within
    med_slack of the floor's median; per tensor rel <= slack * floor + abs_slack for all but a fraction `outliers` of the
    tensors (gate-flip noise is lumpy for small tensors: a 2-sigma tensor is expected among hundreds) and never beyond
    1.6 (= anti-correlated). Returns (violations, summary); empty violations = pass."""
    bad, over = [], []
    names = [n for n in rep if n in floor]
    assert len(names) >= 0.9 * len(rep)
    for n in names:
        if rep[n][1] > 1.6 and rep[n][1] > slack * floor[n][1] + abs_slack:
            bad.append(("grad-hard", n, rep[n][1], floor[n][1]))
        elif rep[n][1] > slack * floor[n][1] + abs_slack:
            over.append(("grad", n, rep[n][1], floor[n][1]))
    if len(over) > max(1, int(outliers * len(names))):
        bad.extend(over)
    med = lambda xs: sorted(xs)[len(xs) // 2]
    m_po, m_fl = med([rep[n][1] for n in names]), med([floor[n][1] for n in names])
    if m_po > med_slack * m_fl + 0.02:
        bad.append(("grad-median", "", m_po, m_fl))
    run_over = [("running", k, v, run_floor.get(k)) for k, v in run_rep.items()
                if v > slack * run_floor.get(k, 0.0) + 2e-2]
    if len(run_over) > max(1, int(outliers * max(1, len(run_rep)))):
        bad.extend(run_over)
    return bad, dict(median_rel=m_po, median_rel_floor=m_fl, outliers=len(over), running_outliers=len(run_over),
                     median_cos=med([rep[n][0] for n in names]), median_cos_floor=med([floor[n][0] for n in names]))
