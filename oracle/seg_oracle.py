"""CPU oracle for the HRNet-OCR-MScale hot path — TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 *restatement* of the reference algorithm (NVIDIA/semantic-segmentation, mounted at
/root/reference during development), written functionally over a ``state_dict`` that uses the reference's own
parameter names.  It is imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` arm; the product path (``semantic-segmentation_b200/``) never imports it.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is pinned against the
reference ITSELF: ``tests/golden/make_golden.py`` imports the reference modules in the build container, feeds them
the deterministic weights/inputs of ``synth_state_dict``/``synth_batch`` below and stores their outputs in
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks this file against those vectors (on CPU, also on the
GPU box where /root/reference does not exist).  The floating-point arithmetic of conv / batch-norm / bilinear /
softmax / inverse / cholesky lives in PyTorch (2.11.0 in this image), the same third-party dependency the reference
calls.

Every function cites the reference file:line it follows.
"""
import math
import zlib

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------- configuration
# config.py:161-190 (cfg.MODEL.OCR_EXTRA) — HRNetV2-W48
HRNET_W48 = dict(
    stem=64,
    stage1=dict(block="BOTTLENECK", num_blocks=[4], num_channels=[64]),
    stage2=dict(num_modules=1, num_blocks=[4, 4], num_channels=[48, 96]),
    stage3=dict(num_modules=4, num_blocks=[4, 4, 4], num_channels=[48, 96, 192]),
    stage4=dict(num_modules=3, num_blocks=[4, 4, 4, 4], num_channels=[48, 96, 192, 384]),
)
# A narrow variant with the same topology, used only to keep committed golden vectors small.
HRNET_W16_TEST = dict(
    stem=64,
    stage1=dict(block="BOTTLENECK", num_blocks=[2], num_channels=[32]),
    stage2=dict(num_modules=1, num_blocks=[1, 1], num_channels=[16, 32]),
    stage3=dict(num_modules=1, num_blocks=[1, 1, 1], num_channels=[16, 32, 64]),
    stage4=dict(num_modules=1, num_blocks=[1, 1, 1, 1], num_channels=[16, 32, 64, 128]),
)
OCR_CFG = dict(mid_channels=512, key_channels=256, num_classes=19, segattn_bot_ch=256, dropout=0.05)  # config.py:157-159,130
BN_EPS = 1e-5
BN_MOMENTUM = 0.1          # hrnetv2.py:25 and torch default
ALIGN_CORNERS = False      # config.py:127


class _RoundBF16(torch.autograd.Function):
    """Round-trip through bf16 in forward AND backward: emulates a tensor (and its gradient) being stored in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundGradBF16(torch.autograd.Function):
    """Identity forward, bf16 round-trip of the gradient (fp32 logits whose gradients are stored in bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundFwdOnly(torch.autograd.Function):
    """bf16 round-trip in forward, exact (fp32) gradient: conv weights are read as bf16, their gradients are fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class Ctx:
    """Execution context: the state dict, train/eval flag and the dropout policy.

    ``emulate_bf16=False`` (default) is the reference algorithm in fp32, pinned by the golden vectors.
    ``emulate_bf16=True`` additionally rounds to bf16 at exactly the points where the B200 path stores bf16
    (conv operands and outputs, activations after BN/ReLU/fuse, class-probability operands, activation gradients);
    all arithmetic stays fp32. It is the matched-precision comparison target for end-to-end GPU tests, because an
    fp32-vs-bf16 comparison of a randomly initialised 300-layer batch-stat-BN network is chaotic."""

    def __init__(self, sd, training=True, drop_mask_fn=None, emulate_bf16=False):
        self.sd = sd
        self.training = training
        self.drop_mask_fn = drop_mask_fn   # callable(shape)->mask or None (None => dropout disabled, p=0 comparison mode)
        self.emulate_bf16 = emulate_bf16

    def q(self, x):
        return _RoundBF16.apply(x) if self.emulate_bf16 else x

    def qg(self, x):
        return _RoundGradBF16.apply(x) if self.emulate_bf16 else x


# ----------------------------------------------------------------------------------------------- primitives
def conv(ctx, name, x, stride=1, padding=0, dilation=1):
    """nn.Conv2d call sites, e.g. hrnetv2.py:31-34; bias present only where the reference leaves bias=True."""
    w = ctx.sd[name + ".weight"]
    b = ctx.sd.get(name + ".bias")
    if ctx.emulate_bf16:
        w = _RoundFwdOnly.apply(w)
    return F.conv2d(x, w, b, stride=stride, padding=padding, dilation=dilation)


def bn(ctx, name, x):
    """Norm2d -> torch.nn.BatchNorm2d (mynn.py:18-24). Train: biased batch variance to normalise, running stats
    updated in place with momentum 0.1 and the unbiased variance."""
    sd = ctx.sd
    if ctx.training and (name + ".num_batches_tracked") in sd:
        sd[name + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=ctx.training, momentum=BN_MOMENTUM, eps=BN_EPS)


def bilinear(x, size):
    """F.interpolate(mode='bilinear', align_corners=cfg.MODEL.ALIGN_CORNERS) — mynn.py:42-48,70-84, hrnetv2.py:246-249."""
    return F.interpolate(x, size=tuple(int(s) for s in size), mode="bilinear", align_corners=ALIGN_CORNERS)


def resize_x(x, scale):
    """mynn.ResizeX (mynn.py:102-114): scale_factor + recompute_scale_factor=True == size=floor(in*scale)."""
    return F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=ALIGN_CORNERS,
                         recompute_scale_factor=True)


# ----------------------------------------------------------------------------------------------- HRNet backbone
def basic_block(ctx, p, x):
    """hrnetv2.BasicBlock.forward (hrnetv2.py:50-66); HRNet branches never use a downsample here."""
    q = ctx.q
    out = q(F.relu(bn(ctx, p + ".bn1", q(conv(ctx, p + ".conv1", x, 1, 1)))))
    out = bn(ctx, p + ".bn2", q(conv(ctx, p + ".conv2", out, 1, 1)))
    return q(F.relu(out + x))


def bottleneck(ctx, p, x, has_downsample):
    """hrnetv2.Bottleneck.forward (hrnetv2.py:86-106), 1x1 -> 3x3 -> 1x1 (x4) with optional 1x1+BN residual."""
    q = ctx.q
    out = q(F.relu(bn(ctx, p + ".bn1", q(conv(ctx, p + ".conv1", x)))))
    out = q(F.relu(bn(ctx, p + ".bn2", q(conv(ctx, p + ".conv2", out, 1, 1)))))
    out = bn(ctx, p + ".bn3", q(conv(ctx, p + ".conv3", out)))
    res = x
    if has_downsample:
        res = bn(ctx, p + ".downsample.1", q(conv(ctx, p + ".downsample.0", x)))
    return q(F.relu(out + res))


def hr_module(ctx, p, xs, num_blocks):
    """HighResolutionModule.forward (hrnetv2.py:230-254) with fuse layers of _make_fuse_layers (hrnetv2.py:181-225)."""
    nb = len(xs)
    xs = list(xs)
    for i in range(nb):
        for k in range(num_blocks[i]):
            xs[i] = basic_block(ctx, "%s.branches.%d.%d" % (p, i, k), xs[i])
    outs = []
    for i in range(nb):
        y = None
        for j in range(nb):
            fp = "%s.fuse_layers.%d.%d" % (p, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = bn(ctx, fp + ".1", ctx.q(conv(ctx, fp + ".0", xs[j])))
                t = bilinear(t, xs[i].shape[-2:])
            else:
                t = xs[j]
                for k in range(i - j):
                    t = bn(ctx, "%s.%d.1" % (fp, k), ctx.q(conv(ctx, "%s.%d.0" % (fp, k), t, 2, 1)))
                    if k != i - j - 1:
                        t = ctx.q(F.relu(t))
            y = t if y is None else y + t
        outs.append(ctx.q(F.relu(y)))
    return outs


def hrnet_forward(ctx, p, x, cfg=HRNET_W48):
    """HighResolutionNet.forward (hrnetv2.py:399-449). Returns the concatenated high-resolution features."""
    q = ctx.q
    x = q(x)   # the B200 path reads the image as bf16
    x = q(F.relu(bn(ctx, p + ".bn1", q(conv(ctx, p + ".conv1", x, 2, 1)))))
    x = q(F.relu(bn(ctx, p + ".bn2", q(conv(ctx, p + ".conv2", x, 2, 1)))))
    s1 = cfg["stage1"]
    for k in range(s1["num_blocks"][0]):   # _make_layer (hrnetv2.py:353-368): 1x1+BN residual iff channel count changes
        x = bottleneck(ctx, "%s.layer1.%d" % (p, k), x,
                       has_downsample=(k == 0 and s1["num_channels"][0] * 4 != cfg["stem"]))
    pre = [s1["num_channels"][0] * 4]
    ys = [x]
    for si, key in ((1, "stage2"), (2, "stage3"), (3, "stage4")):
        sc = cfg[key]
        chans = sc["num_channels"]
        tp = "%s.transition%d" % (p, si)
        xs = []
        for i in range(len(chans)):
            if i < len(pre):
                if chans[i] != pre[i]:   # hrnetv2.py:324-336 (3x3 s1 + BN + ReLU)
                    xs.append(q(F.relu(bn(ctx, "%s.%d.1" % (tp, i), q(conv(ctx, "%s.%d.0" % (tp, i), ys[i], 1, 1))))))
                else:
                    xs.append(ys[i])
            else:                        # hrnetv2.py:338-349: new branch from the LAST previous branch, 3x3 s2 chain
                t = ys[-1]
                for j in range(i + 1 - len(pre)):
                    t = q(F.relu(bn(ctx, "%s.%d.%d.1" % (tp, i, j), q(conv(ctx, "%s.%d.%d.0" % (tp, i, j), t, 2, 1)))))
                xs.append(t)
        for m in range(sc["num_modules"]):
            xs = hr_module(ctx, "%s.%s.%d" % (p, key, m), xs, sc["num_blocks"])
        ys = xs
        pre = chans
    size = ys[0].shape[-2:]
    feats = torch.cat([ys[0]] + [q(bilinear(t, size)) for t in ys[1:]], 1)   # hrnetv2.py:438-447
    return feats


# ----------------------------------------------------------------------------------------------- OCR head
def bn_relu(ctx, name, x):
    """network.utils.BNReLU (utils.py:314-317): Sequential(Norm2d, ReLU) -> sub-module '0' is the BN."""
    return ctx.q(F.relu(bn(ctx, name + ".0", ctx.q(x))))


def spatial_gather(feats, probs, ctx=None):
    """SpatialGather_Module.forward (ocr_utils.py:34-46), scale = 1."""
    q = ctx.q if ctx is not None else (lambda t: t)
    n, k = probs.shape[:2]
    c = feats.shape[1]
    pr = q(F.softmax(probs.reshape(n, k, -1), dim=2))
    ft = feats.reshape(n, c, -1).permute(0, 2, 1)
    ctxv = q(torch.matmul(pr, ft))                    # n x k x c
    return ctxv.permute(0, 2, 1).unsqueeze(3)         # n x c x k x 1


def object_attention(ctx, p, x, proxy, key_ch):
    """ObjectAttentionBlock.forward (ocr_utils.py:95-119), scale = 1 (no pooling)."""
    n, _, h, w = x.shape
    q = bn_relu(ctx, p + ".f_pixel.1", conv(ctx, p + ".f_pixel.0", x))
    q = bn_relu(ctx, p + ".f_pixel.3", conv(ctx, p + ".f_pixel.2", q))
    k = bn_relu(ctx, p + ".f_object.1", conv(ctx, p + ".f_object.0", proxy))
    k = bn_relu(ctx, p + ".f_object.3", conv(ctx, p + ".f_object.2", k))
    v = bn_relu(ctx, p + ".f_down.1", conv(ctx, p + ".f_down.0", proxy))
    q = q.reshape(n, key_ch, -1).permute(0, 2, 1)
    k = k.reshape(n, key_ch, -1)
    v = v.reshape(n, key_ch, -1).permute(0, 2, 1)
    sim = ctx.q(F.softmax((key_ch ** -0.5) * torch.matmul(q, k), dim=-1))
    c = ctx.q(torch.matmul(sim, v)).permute(0, 2, 1).contiguous().reshape(n, key_ch, h, w)
    return bn_relu(ctx, p + ".f_up.1", conv(ctx, p + ".f_up.0", c))


def ocr_block(ctx, p, feats_in, ocfg=OCR_CFG):
    """OCR_block.forward (ocrnet.py:85-91) + SpatialOCR_Module.forward (ocr_utils.py:149-158)."""
    feats = bn_relu(ctx, p + ".conv3x3_ocr.1", conv(ctx, p + ".conv3x3_ocr.0", feats_in, 1, 1))
    aux = ctx.qg(conv(ctx, p + ".aux_head.2", bn_relu(ctx, p + ".aux_head.1", conv(ctx, p + ".aux_head.0", feats_in))))
    context = spatial_gather(feats, aux, ctx)
    dp = p + ".ocr_distri_head"
    oc = object_attention(ctx, dp + ".object_context_block", feats, context, ocfg["key_channels"])
    out = bn_relu(ctx, dp + ".conv_bn_dropout.1", conv(ctx, dp + ".conv_bn_dropout.0", torch.cat([oc, feats], 1)))
    if ctx.training and ctx.drop_mask_fn is not None:   # nn.Dropout2d(0.05): per-(n,c) mask scaled by 1/(1-p)
        mask = ctx.drop_mask_fn((out.shape[0], out.shape[1], 1, 1)).to(out.dtype)
        out = out * mask / (1.0 - ocfg["dropout"])
    cls = ctx.qg(conv(ctx, p + ".cls_head", out))
    return cls, aux, out


def attn_head(ctx, p, x):
    """make_attn_head new-arch (utils.py:343-367): 3x3-BN-ReLU, 3x3-BN-ReLU, 1x1, Sigmoid; all bias-free."""
    q = ctx.q
    x = q(F.relu(bn(ctx, p + ".bn0", q(conv(ctx, p + ".conv0", x, 1, 1)))))
    x = q(F.relu(bn(ctx, p + ".bn1", q(conv(ctx, p + ".conv1", x, 1, 1)))))
    return torch.sigmoid(ctx.qg(conv(ctx, p + ".conv2", x)))


def seg_head(ctx, p, x):
    """make_seg_head (utils.py:320-329), used by basic.HRNet (basic.py:38-64)."""
    q = ctx.q
    x = q(F.relu(bn(ctx, p + ".1", q(conv(ctx, p + ".0", x, 1, 1)))))
    x = q(F.relu(bn(ctx, p + ".4", q(conv(ctx, p + ".3", x, 1, 1)))))
    return ctx.qg(conv(ctx, p + ".6", x))


# ----------------------------------------------------------------------------------------------- losses
def ce_loss(logits, target, ignore_index=255):
    """CrossEntropyLoss2d.forward (loss/utils.py:133-134)."""
    return F.nll_loss(F.log_softmax(logits, dim=1), target, ignore_index=ignore_index, reduction="mean")


def rmi_loss(logits, labels, do_rmi=True, num_classes=19, radius=3, pool=4, pos_alpha=5e-4, clip_min=1e-6,
             weight_lambda=0.5):
    """RMILoss.forward_sigmoid + rmi_lower_bound (loss/rmi.py:82-215), map_get_pairs (rmi_utils.py:15-56),
    log_det_by_cholesky (rmi_utils.py:95-107). lambda_way=1, pool_way=1 (avg), _IS_SUM=1."""
    mask = labels < num_classes
    onehot = F.one_hot(labels.long() * mask.long(), num_classes=num_classes).float()
    maskf = mask.float()
    onehot = onehot * maskf.unsqueeze(3)
    logits_flat = logits.permute(0, 2, 3, 1).contiguous().view(-1, num_classes)
    valid = maskf.sum()
    bce = F.binary_cross_entropy_with_logits(logits_flat, onehot.view(-1, num_classes),
                                             weight=maskf.view(-1, 1), reduction="sum") / (valid + 1.0)
    if not do_rmi:
        return bce
    probs = torch.sigmoid(logits) * maskf.unsqueeze(1) + clip_min
    la = F.avg_pool2d(onehot.permute(0, 3, 1, 2), kernel_size=pool, stride=pool, padding=pool // 2)
    pr = F.avg_pool2d(probs, kernel_size=pool, stride=pool, padding=pool // 2)
    n, c, h, w = la.shape
    nh, nw = h - (radius - 1), w - (radius - 1)
    la_ns, pr_ns = [], []
    for y in range(radius):
        for x in range(radius):
            la_ns.append(la[:, :, y:y + nh, x:x + nw])
            pr_ns.append(pr[:, :, y:y + nh, x:x + nw])
    half_d = radius * radius
    la_v = torch.stack(la_ns, 2).reshape(n, c, half_d, -1).double()
    pr_v = torch.stack(pr_ns, 2).reshape(n, c, half_d, -1).double()
    eye = torch.eye(half_d, dtype=torch.float64, device=logits.device)[None, None]
    la_v = la_v - la_v.mean(dim=3, keepdim=True)
    la_cov = la_v @ la_v.transpose(2, 3)
    pr_v = pr_v - pr_v.mean(dim=3, keepdim=True)
    pr_cov = pr_v @ pr_v.transpose(2, 3)
    pr_cov_inv = torch.inverse(pr_cov + eye * pos_alpha)
    la_pr = la_v @ pr_v.transpose(2, 3)
    appro = la_cov - (la_pr @ pr_cov_inv) @ la_pr.transpose(-2, -1)
    chol = torch.linalg.cholesky(appro + eye * pos_alpha)
    rmi_now = 0.5 * 2.0 * torch.sum(torch.log(torch.diagonal(chol, dim1=-2, dim2=-1) + 1e-8), dim=-1)
    per_class = rmi_now.view(-1, num_classes).mean(dim=0).float() / float(half_d)
    rmi = per_class.sum()
    return weight_lambda * bce + rmi * (1 - weight_lambda)


def criterion_ce(logits, gts, do_rmi=None):
    return ce_loss(logits, gts)


def criterion_rmi(logits, gts, do_rmi=True):
    return rmi_loss(logits, gts, do_rmi=do_rmi)


# ----------------------------------------------------------------------------------------------- networks
def mscale_fwd(ctx, x, hcfg=HRNET_W48, ocfg=OCR_CFG):
    """MscaleOCR._fwd (ocrnet.py:170-183): backbone -> OCR -> attention head -> 3 bilinear Upsample to input size."""
    size = x.shape[2:]
    feats = hrnet_forward(ctx, "backbone", x, hcfg)
    cls, aux, mid = ocr_block(ctx, "ocr", feats, ocfg)
    attn = attn_head(ctx, "scale_attn", mid)
    return dict(cls_out=bilinear(cls, size), aux_out=bilinear(aux, size), logit_attn=bilinear(attn, size),
                cls_q=cls, aux_q=aux, attn_q=attn)


def mscale_basic_fwd(ctx, x, hcfg=HRNET_W48):
    """MscaleBasic._fwd (mscale.py:463-470): trunk features -> attention head and seg head, both scale_as()'d to the
    input size. Same dict keys as mscale_fwd (no auxiliary head)."""
    size = x.shape[2:]
    feats = hrnet_forward(ctx, "backbone", x, hcfg)
    attn = attn_head(ctx, "scale_attn", feats)
    cls = seg_head(ctx, "cls_head", feats)
    return dict(cls_out=bilinear(cls, size), aux_out=None, logit_attn=bilinear(attn, size), cls_q=cls, attn_q=attn)


def mscale_basic_two_scale(ctx, images, gts=None, criterion=criterion_ce, hcfg=HRNET_W48, lo_scale=0.5,
                           supervised_mscale_wt=0.0):
    """MscaleBase.two_scale_forward (mscale.py:182-220) for arch 'mscale.HRNet' (MscaleBasic, mscale.py:450-475)."""
    x_lo = resize_x(images, lo_scale)
    lo = mscale_basic_fwd(ctx, x_lo, hcfg)
    hi = mscale_basic_fwd(ctx, images, hcfg)
    pred_05x, p_1x, attn = lo["cls_out"], hi["cls_out"], lo["logit_attn"]
    size = p_1x.shape[2:]
    joint_pred = bilinear(attn * pred_05x, size) + (1 - bilinear(attn, size)) * p_1x
    if ctx.training:
        loss = criterion(joint_pred, gts)
        if supervised_mscale_wt:
            loss = loss + supervised_mscale_wt * criterion(bilinear(pred_05x, size), gts, do_rmi=False)
            loss = loss + supervised_mscale_wt * criterion(p_1x, gts, do_rmi=False)
        return loss
    return dict(pred=joint_pred, pred_05x=pred_05x, pred_10x=p_1x, attn_05x=attn)


def mscale_basic_nscale(ctx, images, scales, hcfg=HRNET_W48):
    """MscaleBase.nscale_forward (mscale.py:114-180), eval branch."""
    assert 1.0 in scales
    pred = None
    out = {}
    for s in sorted(scales, reverse=True):
        o = mscale_basic_fwd(ctx, resize_x(images, s), hcfg)
        p, attn = o["cls_out"], o["logit_attn"]
        out[fmt_scale("pred", s)] = p
        if s != 2.0:
            out[fmt_scale("attn", s)] = attn
        if pred is None:
            pred = p
        elif s >= 1.0:
            pred = attn * p + (1 - attn) * bilinear(pred, p.shape[2:])
        else:
            p = bilinear(attn * p, pred.shape[2:])
            attn = bilinear(attn, pred.shape[2:])
            pred = p + (1 - attn) * pred
    out["pred"] = pred
    return out


def mscale_two_scale(ctx, images, gts=None, criterion=criterion_ce, hcfg=HRNET_W48, ocfg=OCR_CFG, lo_scale=0.5,
                     ocr_alpha=0.4, ocr_aux_rmi=False, supervised_mscale_wt=0.0):
    """MscaleOCR.two_scale_forward (ocrnet.py:264-327); MscaleBase twin at mscale.py:182-220."""
    x_lo = resize_x(images, lo_scale)
    lo = mscale_fwd(ctx, x_lo, hcfg, ocfg)
    hi = mscale_fwd(ctx, images, hcfg, ocfg)
    pred_05x, pred_10x = lo["cls_out"], hi["cls_out"]
    size = pred_10x.shape[2:]
    attn = lo["logit_attn"]
    p_lo = bilinear(attn * pred_05x, size)
    aux_lo = bilinear(attn * lo["aux_out"], size)
    attn_up = bilinear(attn, size)
    joint_pred = p_lo + (1 - attn_up) * pred_10x
    joint_aux = aux_lo + (1 - attn_up) * hi["aux_out"]
    if ctx.training:
        aux_loss = criterion(joint_aux, gts, do_rmi=ocr_aux_rmi)
        main_loss = criterion(joint_pred, gts, do_rmi=True)
        loss = ocr_alpha * aux_loss + main_loss
        if supervised_mscale_wt:
            loss = loss + supervised_mscale_wt * criterion(bilinear(pred_05x, size), gts, do_rmi=False)
            loss = loss + supervised_mscale_wt * criterion(pred_10x, gts, do_rmi=False)
        return loss
    return dict(pred=joint_pred, pred_05x=pred_05x, pred_10x=pred_10x, attn_05x=attn)


def fmt_scale(prefix, scale):
    """utils/misc.fmt_scale (misc.py:503-513): 'pred_0.5x' style keys."""
    return "%s_%sx" % (prefix, str(float(scale)))


def mscale_nscale(ctx, images, scales, hcfg=HRNET_W48, ocfg=OCR_CFG):
    """MscaleOCR.nscale_forward (ocrnet.py:185-262), eval branch."""
    assert 1.0 in scales
    pred = aux = None
    out = {}
    for s in sorted(scales, reverse=True):
        x = resize_x(images, s)
        o = mscale_fwd(ctx, x, hcfg, ocfg)
        cls, attn, auxo = o["cls_out"], o["logit_attn"], o["aux_out"]
        out[fmt_scale("pred", s)] = cls
        if s != 2.0:
            out[fmt_scale("attn", s)] = attn
        if pred is None:
            pred, aux = cls, auxo
        elif s >= 1.0:
            pred = bilinear(pred, cls.shape[2:])
            pred = attn * cls + (1 - attn) * pred
            aux = bilinear(aux, cls.shape[2:])
            aux = attn * auxo + (1 - attn) * aux
        else:
            cls = bilinear(attn * cls, pred.shape[2:])
            auxo = bilinear(attn * auxo, pred.shape[2:])
            attn = bilinear(attn, pred.shape[2:])
            pred = cls + (1 - attn) * pred
            aux = auxo + (1 - attn) * aux
    out["pred"] = pred
    return out


def ocrnet_forward(ctx, images, gts=None, criterion=criterion_ce, hcfg=HRNET_W48, ocfg=OCR_CFG, ocr_alpha=0.4,
                   ocr_aux_rmi=False):
    """OCRNet.forward (ocrnet.py:104-122): single scale, arch 'ocrnet.HRNet'."""
    feats = hrnet_forward(ctx, "backbone", images, hcfg)
    cls, aux, _ = ocr_block(ctx, "ocr", feats, ocfg)
    size = images.shape[2:]
    aux, cls = bilinear(aux, size), bilinear(cls, size)
    if ctx.training:
        return ocr_alpha * criterion(aux, gts, do_rmi=ocr_aux_rmi) + criterion(cls, gts)
    return dict(pred=cls)


def basic_forward(ctx, images, gts=None, criterion=criterion_ce, hcfg=HRNET_W48):
    """basic.Basic.forward (basic.py:50-64): arch 'basic.HRNet' (BASELINE config 1)."""
    feats = hrnet_forward(ctx, "backbone", images, hcfg)
    pred = bilinear(seg_head(ctx, "seg_head", feats), images.shape[2:])
    if ctx.training:
        return criterion(pred, gts)
    return dict(pred=pred)




# ----------------------------------------------------------------------------------------------- DeepLabV3+ / WRN-38 (§8 f2)
# network/wider_resnet.py:270-435 (WiderResNetA2, dilation=True, structure 38) and :398-435 (wrn38 wrapper: no bn_out)
WRN38 = dict(structure=[3, 3, 6, 3, 1, 1],
             channels=[(128, 128), (256, 256), (512, 512), (512, 1024), (512, 1024, 2048), (1024, 2048, 4096)])
ASPP_RATES = (12, 24, 36)      # utils.py:176-183: rates (6, 12, 18) doubled at output stride 8
ASPP_DIM = 256                 # deepv3.py:50-53 bottleneck_ch


def _wrn_block_plan(wcfg=WRN38):
    """-> list of (module name, block name, in_ch, channels tuple, stride, dilation, dropout p) in forward order
    (wider_resnet.py:313-345 with dilation=True)."""
    plan = []
    in_ch = 64
    for mod_id, num in enumerate(wcfg["structure"]):
        for block_id in range(num):
            dil = 2 if mod_id == 3 else (4 if mod_id > 3 else 1)
            stride = 2 if (block_id == 0 and mod_id == 2) else 1
            drop = 0.3 if mod_id == 4 else (0.5 if mod_id == 5 else None)
            plan.append(("mod%d" % (mod_id + 2), "block%d" % (block_id + 1), in_ch, wcfg["channels"][mod_id], stride,
                         dil, drop))
            in_ch = wcfg["channels"][mod_id][-1]
    return plan


def wrn_block(ctx, p, x, in_ch, channels, stride, dil, drop):
    """IdentityResidualBlock.forward (wider_resnet.py:170-183): pre-activation BN-ReLU, optional 1x1 projection of the
    ACTIVATED input, 2x(3x3) or 1x1-3x3-1x1 body (dilated), Dropout2d before the last conv (mod6/mod7)."""
    q = ctx.q
    need_proj = stride != 1 or in_ch != channels[-1]
    a = q(F.relu(bn(ctx, p + ".bn1.0", x)))
    shortcut = conv(ctx, p + ".proj_conv", a, stride) if need_proj else x
    c = p + ".convs"

    def dropout(t):
        if drop is not None and ctx.training and ctx.drop_mask_fn is not None:
            m = ctx.drop_mask_fn((t.shape[0], t.shape[1], 1, 1)).to(t.dtype)
            return t * m / (1.0 - drop)
        return t

    if len(channels) == 2:
        t = q(conv(ctx, c + ".conv1", a, stride, dil, dil))
        t = q(F.relu(bn(ctx, c + ".bn2.0", t)))
        t = conv(ctx, c + ".conv2", dropout(t), 1, dil, dil)
    else:
        t = q(conv(ctx, c + ".conv1", a, stride))
        t = q(F.relu(bn(ctx, c + ".bn2.0", t)))
        t = q(conv(ctx, c + ".conv2", t, 1, dil, dil))
        t = q(F.relu(bn(ctx, c + ".bn3.0", t)))
        t = conv(ctx, c + ".conv3", dropout(t))
    return q(t + shortcut)


def wrn38_forward(ctx, p, x, wcfg=WRN38):
    """wrn38.forward (wider_resnet.py:423-435): -> (s2 features [/2], s4 features [/4], final [/8], all pre-activation)."""
    x = ctx.q(conv(ctx, p + ".mod1.conv1", x, 1, 1))
    feats = {}
    for mod, blk, in_ch, channels, stride, dil, drop in _wrn_block_plan(wcfg):
        if blk == "block1" and mod in ("mod2", "mod3"):
            x = F.max_pool2d(x, 3, stride=2, padding=1)          # pool2 / pool3
        x = wrn_block(ctx, "%s.%s.%s" % (p, mod, blk), x, in_ch, channels, stride, dil, drop)
        feats[mod] = x
    return feats["mod2"], feats["mod3"], x


def aspp(ctx, p, x):
    """AtrousSpatialPyramidPoolingModule.forward (utils.py:204-216): image pooling branch first, then 1x1 and the three
    dilated 3x3 branches, concatenated."""
    q = ctx.q
    img = F.adaptive_avg_pool2d(x, 1)
    img = q(F.relu(bn(ctx, p + ".img_conv.1", q(conv(ctx, p + ".img_conv.0", img)))))
    out = [bilinear(img, x.shape[2:])]
    out.append(q(F.relu(bn(ctx, p + ".features.0.1", q(conv(ctx, p + ".features.0.0", x))))))
    for i, r in enumerate(ASPP_RATES):
        f = "%s.features.%d" % (p, i + 1)
        out.append(q(F.relu(bn(ctx, f + ".1", q(conv(ctx, f + ".0", x, 1, r, r))))))
    return torch.cat(out, 1)


def deepv3_forward(ctx, images, gts=None, criterion=criterion_ce, wcfg=WRN38):
    """DeepV3Plus.forward (deepv3.py:75-96), arch 'deepv3.DeepV3PlusW38' (BASELINE config 4)."""
    q = ctx.q
    s2, _s4, final = wrn38_forward(ctx, "backbone", images, wcfg)
    a = aspp(ctx, "aspp", final)
    conv_aspp = q(conv(ctx, "bot_aspp", a))
    conv_s2 = q(conv(ctx, "bot_fine", s2))
    cat = torch.cat([conv_s2, bilinear(conv_aspp, s2.shape[2:])], 1)
    t = q(F.relu(bn(ctx, "final.1", q(conv(ctx, "final.0", cat, 1, 1)))))
    t = q(F.relu(bn(ctx, "final.4", q(conv(ctx, "final.3", t, 1, 1)))))
    out = bilinear(ctx.qg(conv(ctx, "final.6", t)), images.shape[2:])
    if ctx.training:
        return criterion(out, gts)
    return dict(pred=out)


def _deepv3_names(wcfg=WRN38, num_classes=19):
    """(name, shape) of every tensor of DeepV3PlusW38's state_dict, in registration order (deepv3.py:46-65,
    wider_resnet.py:303-350, utils.py:171-202)."""
    out = []

    def conv_(name, o, i, k):
        out.append((name + ".weight", (o, i, k, k)))

    def bn_(name, c):
        out.extend([(name + ".weight", (c,)), (name + ".bias", (c,)), (name + ".running_mean", (c,)),
                    (name + ".running_var", (c,)), (name + ".num_batches_tracked", ())])

    conv_("backbone.mod1.conv1", 64, 3, 3)
    for mod, blk, in_ch, ch, stride, dil, drop in _wrn_block_plan(wcfg):
        b = "backbone.%s.%s" % (mod, blk)
        bn_(b + ".bn1.0", in_ch)
        if len(ch) == 2:
            conv_(b + ".convs.conv1", ch[0], in_ch, 3); bn_(b + ".convs.bn2.0", ch[0])
            conv_(b + ".convs.conv2", ch[1], ch[0], 3)
        else:
            conv_(b + ".convs.conv1", ch[0], in_ch, 1); bn_(b + ".convs.bn2.0", ch[0])
            conv_(b + ".convs.conv2", ch[1], ch[0], 3); bn_(b + ".convs.bn3.0", ch[1])
            conv_(b + ".convs.conv3", ch[2], ch[1], 1)
        if stride != 1 or in_ch != ch[-1]:
            conv_(b + ".proj_conv", ch[-1], in_ch, 1)
    high = wcfg["channels"][-1][-1]
    s2_ch = wcfg["channels"][0][-1]
    conv_("aspp.features.0.0", ASPP_DIM, high, 1); bn_("aspp.features.0.1", ASPP_DIM)
    for i in range(len(ASPP_RATES)):
        conv_("aspp.features.%d.0" % (i + 1), ASPP_DIM, high, 3); bn_("aspp.features.%d.1" % (i + 1), ASPP_DIM)
    conv_("aspp.img_conv.0", ASPP_DIM, high, 1); bn_("aspp.img_conv.1", ASPP_DIM)
    conv_("bot_fine", 48, s2_ch, 1)
    conv_("bot_aspp", 256, ASPP_DIM * (2 + len(ASPP_RATES)), 1)
    conv_("final.0", 256, 256 + 48, 3); bn_("final.1", 256)
    conv_("final.3", 256, 256, 3); bn_("final.4", 256)
    conv_("final.6", num_classes, 256, 1)
    return out


# ----------------------------------------------------------------------------------------------- state dict synthesis
def _conv_names(hcfg, ocfg, arch):
    """Enumerate (name, shape) of every tensor of the reference state_dict for the given arch, in no particular
    order. Mirrors the constructors hrnetv2.py:263-315, ocrnet.py:46-83, utils.py:320-367."""
    out = []

    def conv_(name, o, i, k, bias=False):
        out.append((name + ".weight", (o, i, k, k)))
        if bias:
            out.append((name + ".bias", (o,)))

    def bn_(name, c):
        out.extend([(name + ".weight", (c,)), (name + ".bias", (c,)), (name + ".running_mean", (c,)),
                    (name + ".running_var", (c,)), (name + ".num_batches_tracked", ())])

    p = "backbone"
    conv_(p + ".conv1", 64, 3, 3); bn_(p + ".bn1", 64)
    conv_(p + ".conv2", 64, 64, 3); bn_(p + ".bn2", 64)
    s1 = hcfg["stage1"]
    planes = s1["num_channels"][0]
    inpl = 64
    for k in range(s1["num_blocks"][0]):
        b = "%s.layer1.%d" % (p, k)
        conv_(b + ".conv1", planes, inpl, 1); bn_(b + ".bn1", planes)
        conv_(b + ".conv2", planes, planes, 3); bn_(b + ".bn2", planes)
        conv_(b + ".conv3", planes * 4, planes, 1); bn_(b + ".bn3", planes * 4)
        if k == 0 and inpl != planes * 4:
            conv_(b + ".downsample.0", planes * 4, inpl, 1); bn_(b + ".downsample.1", planes * 4)
        inpl = planes * 4
    pre = [planes * 4]
    for si, key in ((1, "stage2"), (2, "stage3"), (3, "stage4")):
        sc = hcfg[key]
        chans = sc["num_channels"]
        tp = "%s.transition%d" % (p, si)
        for i in range(len(chans)):
            if i < len(pre):
                if chans[i] != pre[i]:
                    conv_("%s.%d.0" % (tp, i), chans[i], pre[i], 3); bn_("%s.%d.1" % (tp, i), chans[i])
            else:
                for j in range(i + 1 - len(pre)):
                    oc = chans[i] if j == i - len(pre) else pre[-1]
                    conv_("%s.%d.%d.0" % (tp, i, j), oc, pre[-1], 3); bn_("%s.%d.%d.1" % (tp, i, j), oc)
        nb = len(chans)
        for m in range(sc["num_modules"]):
            mp = "%s.%s.%d" % (p, key, m)
            for i in range(nb):
                for k in range(sc["num_blocks"][i]):
                    b = "%s.branches.%d.%d" % (mp, i, k)
                    conv_(b + ".conv1", chans[i], chans[i], 3); bn_(b + ".bn1", chans[i])
                    conv_(b + ".conv2", chans[i], chans[i], 3); bn_(b + ".bn2", chans[i])
            for i in range(nb):
                for j in range(nb):
                    fp = "%s.fuse_layers.%d.%d" % (mp, i, j)
                    if j > i:
                        conv_(fp + ".0", chans[i], chans[j], 1); bn_(fp + ".1", chans[i])
                    elif j < i:
                        for k in range(i - j):
                            oc = chans[i] if k == i - j - 1 else chans[j]
                            conv_("%s.%d.0" % (fp, k), oc, chans[j], 3); bn_("%s.%d.1" % (fp, k), oc)
        pre = chans
    high = sum(pre)
    if arch in ("ocrnet.HRNet", "ocrnet.HRNet_Mscale"):
        mid, key_ch, ncls = ocfg["mid_channels"], ocfg["key_channels"], ocfg["num_classes"]
        o = "ocr"
        conv_(o + ".conv3x3_ocr.0", mid, high, 3, True); bn_(o + ".conv3x3_ocr.1.0", mid)
        ob = o + ".ocr_distri_head.object_context_block"
        conv_(ob + ".f_pixel.0", key_ch, mid, 1); bn_(ob + ".f_pixel.1.0", key_ch)
        conv_(ob + ".f_pixel.2", key_ch, key_ch, 1); bn_(ob + ".f_pixel.3.0", key_ch)
        conv_(ob + ".f_object.0", key_ch, mid, 1); bn_(ob + ".f_object.1.0", key_ch)
        conv_(ob + ".f_object.2", key_ch, key_ch, 1); bn_(ob + ".f_object.3.0", key_ch)
        conv_(ob + ".f_down.0", key_ch, mid, 1); bn_(ob + ".f_down.1.0", key_ch)
        conv_(ob + ".f_up.0", mid, key_ch, 1); bn_(ob + ".f_up.1.0", mid)
        cb = o + ".ocr_distri_head.conv_bn_dropout"
        conv_(cb + ".0", mid, 2 * mid, 1); bn_(cb + ".1.0", mid)
        conv_(o + ".cls_head", ncls, mid, 1, True)
        conv_(o + ".aux_head.0", high, high, 1, True); bn_(o + ".aux_head.1.0", high)
        conv_(o + ".aux_head.2", ncls, high, 1, True)
        if arch == "ocrnet.HRNet_Mscale":
            bot = ocfg["segattn_bot_ch"]
            a = "scale_attn"
            conv_(a + ".conv0", bot, mid, 3); bn_(a + ".bn0", bot)
            conv_(a + ".conv1", bot, bot, 3); bn_(a + ".bn1", bot)
            conv_(a + ".conv2", 1, bot, 1)
    elif arch == "basic.HRNet":
        bot, ncls = ocfg["segattn_bot_ch"], ocfg["num_classes"]
        s = "seg_head"
        conv_(s + ".0", bot, high, 3); bn_(s + ".1", bot)
        conv_(s + ".3", bot, bot, 3); bn_(s + ".4", bot)
        conv_(s + ".6", ncls, bot, 1)
    elif arch == "mscale.HRNet":     # MscaleBasic.__init__ (mscale.py:450-461)
        bot, ncls = ocfg["segattn_bot_ch"], ocfg["num_classes"]
        s = "cls_head"
        conv_(s + ".0", bot, high, 3); bn_(s + ".1", bot)
        conv_(s + ".3", bot, bot, 3); bn_(s + ".4", bot)
        conv_(s + ".6", ncls, bot, 1)
        a = "scale_attn"
        conv_(a + ".conv0", bot, high, 3); bn_(a + ".bn0", bot)
        conv_(a + ".conv1", bot, bot, 3); bn_(a + ".bn1", bot)
        conv_(a + ".conv2", 1, bot, 1)
    else:
        raise ValueError(arch)
    return out


def synth_state_dict(arch="ocrnet.HRNet_Mscale", hcfg=HRNET_W48, ocfg=OCR_CFG, seed=0, dtype=torch.float32):
    """Deterministic, reference-independent weights: every tensor is drawn from a torch CPU generator seeded by
    crc32(name) ^ seed, so the GPU box regenerates exactly what the golden script fed to the reference. Conv weights
    use a fan-in scaled normal (well-conditioned activations, unlike the reference's N(0, 1e-3) default init),
    BN affine parameters are perturbed around (1, 0), running stats start at (0, 1) like a fresh BatchNorm2d."""
    sd = {}
    names = _deepv3_names(hcfg if "structure" in hcfg else WRN38, ocfg["num_classes"]) \
        if arch == "deepv3.DeepV3PlusW38" else _conv_names(hcfg, ocfg, arch)
    for name, shape in names:
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
        if name.endswith("num_batches_tracked"):
            t = torch.zeros((), dtype=torch.long)
        elif name.endswith("running_mean"):
            t = torch.zeros(shape, dtype=dtype)
        elif name.endswith("running_var"):
            t = torch.ones(shape, dtype=dtype)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            t = (torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)).to(dtype)
        elif ".bn" in name or name.split(".")[-2].isdigit() and name.endswith(".weight") and len(shape) == 1:
            t = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
        else:
            t = (0.1 * torch.randn(shape, generator=g)).to(dtype)
        sd[name] = t
    # make BN weights ~1 and BN biases ~0 regardless of naming quirks above
    for name in list(sd):
        if name.endswith(".running_mean"):
            base = name[: -len(".running_mean")]
            g = torch.Generator().manual_seed((zlib.crc32((base + "#affine").encode()) ^ seed) & 0x7FFFFFFF)
            c = sd[name].shape[0]
            sd[base + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
            sd[base + ".bias"] = (0.1 * torch.randn(c, generator=g)).to(dtype)
    return sd


def synth_batch(n, h, w, seed=0, num_classes=19, ignore_rows=8):
    """SURVEY §8d synthetic inputs: images ~ N(0,1), labels uniform over classes with the top rows = 255."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn((n, 3, h, w), generator=g)
    gts = torch.randint(0, num_classes, (n, h, w), generator=g)
    gts[:, : min(ignore_rows, h // 4)] = 255
    return images, gts


def clone_sd(sd):
    return {k: v.clone() for k, v in sd.items()}


def sample_like(t, target=4096):
    """Deterministic strided sample used by the golden fixtures: stride = the smallest odd number that brings the
    element count under `target` (odd so it does not alias with power-of-two image widths)."""
    flat = t.detach().flatten()
    stride = max(1, flat.numel() // target)
    if stride % 2 == 0:
        stride += 1
    return flat[::stride].clone()
