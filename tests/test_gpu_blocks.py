"""Teacher-forced block-level parity (forward AND backward) of the B200 executor against the oracle run with bf16
storage emulation (oracle.seg_oracle.Ctx(emulate_bf16=True)): both sides see identical bf16-exact inputs and upstream
gradients, so differences are limited to fp32 summation order and the rare 1-ulp bf16 rounding flip it causes.
Whole-network fp32-vs-bf16 comparisons of a randomly initialised batch-stat-BN network are chaotic (SURVEY.md §7
hard part 1) and therefore only checked statistically in test_gpu_model.py."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class Harness:
    def __init__(self, arch="ocrnet.HRNet_Mscale", sharpen_aux=1.0):
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        from oracle import seg_oracle as O
        from b200seg.module import B200SegModule
        from b200seg.engine import Engine, Act
        from b200seg import model as M, raw
        self.O, self.M, self.raw, self.Act = O, M, raw, Act
        self.hcfg = O.HRNET_W16_TEST
        sd0 = O.synth_state_dict(arch, self.hcfg, seed=3)
        if sharpen_aux != 1.0:
            # With random weights the per-class soft regions are almost uniform, every class context collapses onto the
            # mean feature and the 38-sample BatchNorm of f_object / f_down amplifies rounding noise ~50x. Sharper
            # auxiliary logits give distinct class contexts (what a trained network has) and a well-conditioned test.
            sd0["ocr.aux_head.2.weight"] = sd0["ocr.aux_head.2.weight"] * sharpen_aux
        self.sd = {k: v.clone().cuda() for k, v in sd0.items()}
        for k, v in self.sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
        self.ctx = O.Ctx(self.sd, training=True, emulate_bf16=True)
        self.ocfg = dict(O.OCR_CFG)
        self.ocfg["dropout"] = 0.0
        net = B200SegModule(arch, 19, hcfg=self.hcfg, ocfg=self.ocfg, use_cuda_graph=False)
        net.load_state_dict(sd0)
        self.net = net.cuda().train()
        net._ensure_device_state()
        net._repack()
        tensors = {k: v.detach() for k, v in net._tensors().items()}
        self.grads = net._engine_grads("hi")        # conv weights: kernel-layout [O][taps][I] accumulators
        self.E = Engine(tensors, self.grads, net._packed, True, torch.ones((2, 512), device="cuda"))

    def rand_bf16(self, shape, seed, scale=1.0):
        g = torch.Generator(device="cuda").manual_seed(seed)
        return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16).float()

    def act(self, t_nchw):
        return self.Act(t_nchw.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))

    @staticmethod
    def nhwc(t_nchw):
        return t_nchw.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)

    @staticmethod
    def check(a, b, name, tol=0.02, cos_min=0.999):
        a, b = a.double().flatten(), b.double().flatten()
        rel = float((a - b).norm() / (b.norm() + 1e-30))
        c = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        assert rel <= tol and c >= cos_min, "%s: relL2 %.4f cos %.5f" % (name, rel, c)

    def check_param_grads(self, prefix, tol=0.02):
        n = 0
        for name, v in self.sd.items():
            if name.startswith(prefix) and v.grad is not None and name in self.grads:
                if name in ("ocr.conv3x3_ocr.0.bias", "ocr.aux_head.0.bias"):
                    continue   # conv bias in front of a training-mode BN: exactly zero here, float noise in the oracle
                g = self.grads[name]
                if v.grad.dim() == 4:
                    g = self.raw.ohwi_to_oihw(g, v.grad.shape[2])
                self.check(g, v.grad, "grad " + name, tol, 0.998)
                n += 1
        assert n > 0


def test_basic_block_and_bottleneck():
    h = Harness()
    O, M, E = h.O, h.M, h.E
    x = h.rand_bf16((2, 16, 16, 32), 1).abs().requires_grad_(True)
    out = O.basic_block(h.ctx, "backbone.stage2.0.branches.0.0", x)
    dz = h.rand_bf16(out.shape, 2)
    out.backward(dz)
    xa = h.act(x)
    za = M.basic_block(E, "backbone.stage2.0.branches.0.0", xa)
    h.check(za.t.permute(0, 3, 1, 2), out, "basic block fwd", 0.005)
    za.grad = h.nhwc(dz)
    E.run_backward()
    h.check(xa.grad.permute(0, 3, 1, 2), x.grad, "basic block dx")
    h.check_param_grads("backbone.stage2.0.branches.0.0")

    x = h.rand_bf16((2, 64, 16, 32), 3).abs().requires_grad_(True)
    out = O.bottleneck(h.ctx, "backbone.layer1.0", x, True)
    dz = h.rand_bf16(out.shape, 4)
    out.backward(dz)
    xa = h.act(x)
    za = M.bottleneck(E, "backbone.layer1.0", xa, True)
    h.check(za.t.permute(0, 3, 1, 2), out, "bottleneck fwd", 0.005)
    za.grad = h.nhwc(dz)
    E.run_backward()
    h.check(xa.grad.permute(0, 3, 1, 2), x.grad, "bottleneck dx")
    h.check_param_grads("backbone.layer1.0")


def test_hr_module_four_branches():
    h = Harness()
    O, M, E = h.O, h.M, h.E
    chans = h.hcfg["stage4"]["num_channels"]
    xs = [h.rand_bf16((2, c, 16 >> i, 32 >> i), 10 + i).abs().requires_grad_(True) for i, c in enumerate(chans)]
    outs = O.hr_module(h.ctx, "backbone.stage4.0", xs, [1, 1, 1, 1])
    dzs = [h.rand_bf16(o.shape, 20 + i) for i, o in enumerate(outs)]
    torch.autograd.backward(outs, dzs)
    xas = [h.act(x) for x in xs]
    zas = M.hr_module(E, "backbone.stage4.0", xas, [1, 1, 1, 1])
    for i in range(4):
        h.check(zas[i].t.permute(0, 3, 1, 2), outs[i], "hr_module out%d" % i, 0.01, 0.9995)
        zas[i].grad = h.nhwc(dzs[i])
    E.run_backward()
    for i in range(4):
        h.check(xas[i].grad.permute(0, 3, 1, 2), xs[i].grad, "hr_module dx%d" % i, 0.03, 0.999)
    h.check_param_grads("backbone.stage4.0", 0.03)


def test_spatial_gather_fwd_bwd():
    """SpatialGather glue (network/ocr_utils.py:34-46) in isolation: teacher-forced feats / aux logits."""
    h = Harness()
    M, E = h.M, h.E
    from b200seg.engine import HeadRec
    n, hh, ww, C, K = 2, 16, 32, 512, 19
    feats = h.rand_bf16((n, C, hh, ww), 50).abs().requires_grad_(True)
    logits = (h.rand_bf16((n, K, hh, ww), 51) * 4).float().requires_grad_(True)
    q = h.ctx.q
    pr = q(F.softmax(logits.reshape(n, K, -1), dim=2))
    ctxv = q(torch.matmul(pr, q(feats).reshape(n, C, -1).permute(0, 2, 1)))      # n x K x C
    d_ctx = h.rand_bf16(ctxv.shape, 52)
    ctxv.backward(d_ctx)
    fa = h.act(feats)
    aux = HeadRec()
    buf = torch.zeros((n, hh, ww, 20), device="cuda")
    buf[..., :K] = logits.detach().permute(0, 2, 3, 1)
    aux.logits, aux.dlogits = buf[..., :K], None
    proxy = M.spatial_gather(E, fa, aux, K)
    h.check(proxy.t.view(n, K, C), ctxv, "gather fwd", 0.005, 0.9999)
    proxy.grad = d_ctx.view(n, K, 1, C).to(torch.bfloat16)
    E.run_backward()
    h.check(fa.grad.permute(0, 3, 1, 2), feats.grad, "gather d feats", 0.01, 0.9999)
    h.check(aux.dlogits[..., :K].permute(0, 3, 1, 2), logits.grad, "gather d logits", 0.02, 0.9995)


def test_object_attention_core_fwd_bwd():
    """ObjectAttentionBlock core (network/ocr_utils.py:100-113) in isolation: teacher-forced q / k / v."""
    h = Harness()
    M, E = h.M, h.E
    n, hh, ww, C, K = 2, 16, 32, 256, 19
    qv = h.rand_bf16((n, C, hh, ww), 60).abs().requires_grad_(True)
    kv = h.rand_bf16((n, C, K, 1), 61).abs().requires_grad_(True)
    vv = h.rand_bf16((n, C, K, 1), 62).abs().requires_grad_(True)
    qq = h.ctx.q
    query = qv.reshape(n, C, -1).permute(0, 2, 1)
    sim = qq(F.softmax((C ** -0.5) * torch.matmul(query, kv.reshape(n, C, -1)), dim=-1))
    out = qq(torch.matmul(sim, vv.reshape(n, C, -1).permute(0, 2, 1)))         # n x P x C
    d_out = h.rand_bf16(out.shape, 63)
    out.backward(d_out)
    qa, ka, va = h.act(qv), h.act(kv), h.act(vv)
    ctx = M.object_attention(E, qa, ka, va, K)
    h.check(ctx.t.reshape(n, hh * ww, C), out, "attention fwd", 0.005, 0.9999)
    ctx.grad = d_out.view(n, hh, ww, C).to(torch.bfloat16)
    E.run_backward()
    h.check(qa.grad.permute(0, 3, 1, 2), qv.grad, "attention dq", 0.02, 0.9995)
    h.check(ka.grad.permute(0, 3, 1, 2), kv.grad, "attention dk", 0.02, 0.9995)
    h.check(va.grad.permute(0, 3, 1, 2), vv.grad, "attention dv", 0.02, 0.9995)


def test_ocr_block_forward_and_aux_path():
    """Whole OCR block: forward plumbing (concat views, dropout scale, heads) and the aux-head gradient path, which
    does not cross the noise-amplifying 38-sample BatchNorms of f_object / f_down."""
    h = Harness(sharpen_aux=6.0)
    O, M, E = h.O, h.M, h.E
    high = sum(h.hcfg["stage4"]["num_channels"])
    x = h.rand_bf16((2, high, 16, 32), 30).abs().requires_grad_(True)
    cls, aux, mid = O.ocr_block(h.ctx, "ocr", x, h.ocfg)
    d_aux = h.rand_bf16(aux.shape, 32, 0.1)
    aux.backward(d_aux)
    xa = h.act(x)
    cls_r, aux_r, mid_a = M.ocr_block(E, xa, h.ocfg)
    h.check(aux_r.logits.permute(0, 3, 1, 2), aux, "aux logits", 0.005)
    h.check(mid_a.t.permute(0, 3, 1, 2), mid, "ocr feats", 0.03, 0.9995)
    h.check(cls_r.logits.permute(0, 3, 1, 2), cls, "cls logits", 0.03, 0.9995)
    pad = lambda t: F.pad(t.permute(0, 2, 3, 1), (0, 32 - t.shape[1])).contiguous().to(torch.bfloat16)
    aux_r.dlogits = pad(d_aux)
    cls_r.dlogits = torch.zeros((2, 16, 32, 32), dtype=torch.bfloat16, device="cuda")
    mid_a.grad = torch.zeros((2, 16, 32, 512), dtype=torch.bfloat16, device="cuda")
    E.run_backward()
    h.check(xa.grad.permute(0, 3, 1, 2), x.grad, "ocr block dx (aux path)", 0.02, 0.9995)
    h.check_param_grads("ocr.aux_head", 0.02)


def test_attention_head():
    h = Harness()
    O, M, E = h.O, h.M, h.E
    x = h.rand_bf16((2, 512, 16, 32), 40).abs().requires_grad_(True)
    a = O.attn_head(h.ctx, "scale_attn", x)
    d_a = h.rand_bf16(a.shape, 41, 0.1)
    a.backward(d_a)
    xa = h.act(x)
    rec = M.attn_head(E, xa)
    h.check(torch.sigmoid(rec.logits.permute(0, 3, 1, 2)), a, "attn", 0.005)
    s = torch.sigmoid(rec.logits)
    dl = torch.zeros((2, 16, 32, 8), dtype=torch.bfloat16, device="cuda")
    dl[..., :1] = (d_a.permute(0, 2, 3, 1) * s * (1 - s)).to(torch.bfloat16)
    rec.dlogits = dl
    E.run_backward()
    h.check(xa.grad.permute(0, 3, 1, 2), x.grad, "attn head dx", 0.03)
    h.check_param_grads("scale_attn", 0.03)


def test_backbone_forward_statistics():
    """Full backbone: chaotic amplification allowed, but the fused concat/upsample plumbing must stay consistent."""
    h = Harness()
    O, M, E = h.O, h.M, h.E
    images, _ = O.synth_batch(2, 64, 128, seed=5)
    images = images.cuda()
    with torch.no_grad():
        feats = O.hrnet_forward(h.ctx, "backbone", images, h.hcfg)
    x16 = h.Act(h.raw.image_prep(images, 64, 128), needs_grad=False)
    cat = M.hrnet_forward(E, x16, h.hcfg)
    a, b = cat.t.permute(0, 3, 1, 2).double().flatten(), feats.double().flatten()
    rel = float((a - b).norm() / b.norm())
    assert rel < 0.06, rel
