"""Architecture description shared by the module constructor and the executor.

``tensor_specs(arch, hcfg, ocfg)`` lists, in the reference's registration order, every parameter/buffer of the
reference module for ``arch`` (names/shapes as in SURVEY.md §8b "State": 1903 keys / 72.14 M parameters for
``ocrnet.HRNet_Mscale``), so checkpoints and optimizer state move between the two implementations by name and
position. Constructors followed: network/hrnetv2.py:263-315 (HighResolutionNet.__init__), network/ocrnet.py:46-83
(OCR_block), network/ocr_utils.py:61-93,129-147, network/utils.py:320-367 (seg / attention heads).
"""

# config.py:161-190 — cfg.MODEL.OCR_EXTRA (HRNetV2-W48)
HRNET_W48 = dict(
    stem=64,
    stage1=dict(num_blocks=[4], num_channels=[64]),
    stage2=dict(num_modules=1, num_blocks=[4, 4], num_channels=[48, 96]),
    stage3=dict(num_modules=4, num_blocks=[4, 4, 4], num_channels=[48, 96, 192]),
    stage4=dict(num_modules=3, num_blocks=[4, 4, 4, 4], num_channels=[48, 96, 192, 384]),
)
# config.py:157-159 (OCR), :130 (SEGATTN_BOT_CH), network/ocrnet.py:63 (dropout)
OCR_DEFAULT = dict(mid_channels=512, key_channels=256, num_classes=19, segattn_bot_ch=256, dropout=0.05)

ARCHS = ("ocrnet.HRNet_Mscale", "ocrnet.HRNet", "basic.HRNet")


def hrnet_cfg_from_reference_cfg(cfg):
    """Read the HRNet widths from a reference-style cfg (cfg.MODEL.OCR_EXTRA) when one is importable."""
    ex = cfg.MODEL.OCR_EXTRA
    out = dict(stem=64, stage1=dict(num_blocks=list(ex.STAGE1.NUM_BLOCKS), num_channels=list(ex.STAGE1.NUM_CHANNELS)))
    for key, ref in (("stage2", ex.STAGE2), ("stage3", ex.STAGE3), ("stage4", ex.STAGE4)):
        out[key] = dict(num_modules=ref.NUM_MODULES, num_blocks=list(ref.NUM_BLOCKS),
                        num_channels=list(ref.NUM_CHANNELS))
    return out


def high_level_channels(hcfg):
    return sum(hcfg["stage4"]["num_channels"])


def tensor_specs(arch, hcfg=HRNET_W48, ocfg=OCR_DEFAULT):
    """-> list of (name, shape, kind) with kind in {'conv_w','conv_b','bn_w','bn_b','bn_rm','bn_rv','bn_nbt'}."""
    out = []

    def conv(name, o, i, k, bias=False):
        out.append((name + ".weight", (o, i, k, k), "conv_w"))
        if bias:
            out.append((name + ".bias", (o,), "conv_b"))

    def bn(name, c):
        out.append((name + ".weight", (c,), "bn_w"))
        out.append((name + ".bias", (c,), "bn_b"))
        out.append((name + ".running_mean", (c,), "bn_rm"))
        out.append((name + ".running_var", (c,), "bn_rv"))
        out.append((name + ".num_batches_tracked", (), "bn_nbt"))

    p = "backbone"
    conv(p + ".conv1", 64, 3, 3); bn(p + ".bn1", 64)
    conv(p + ".conv2", 64, 64, 3); bn(p + ".bn2", 64)
    s1 = hcfg["stage1"]
    planes, inpl = s1["num_channels"][0], hcfg["stem"]
    for k in range(s1["num_blocks"][0]):
        b = "%s.layer1.%d" % (p, k)
        conv(b + ".conv1", planes, inpl, 1); bn(b + ".bn1", planes)
        conv(b + ".conv2", planes, planes, 3); bn(b + ".bn2", planes)
        conv(b + ".conv3", planes * 4, planes, 1); bn(b + ".bn3", planes * 4)
        if k == 0 and inpl != planes * 4:
            conv(b + ".downsample.0", planes * 4, inpl, 1); bn(b + ".downsample.1", planes * 4)
        inpl = planes * 4
    pre = [planes * 4]
    for si, key in ((1, "stage2"), (2, "stage3"), (3, "stage4")):
        sc = hcfg[key]
        ch = sc["num_channels"]
        tp = "%s.transition%d" % (p, si)
        for i in range(len(ch)):
            if i < len(pre):
                if ch[i] != pre[i]:
                    conv("%s.%d.0" % (tp, i), ch[i], pre[i], 3); bn("%s.%d.1" % (tp, i), ch[i])
            else:
                for j in range(i + 1 - len(pre)):
                    oc = ch[i] if j == i - len(pre) else pre[-1]
                    conv("%s.%d.%d.0" % (tp, i, j), oc, pre[-1], 3); bn("%s.%d.%d.1" % (tp, i, j), oc)
        nb = len(ch)
        for m in range(sc["num_modules"]):
            mp = "%s.%s.%d" % (p, key, m)
            for i in range(nb):
                for k in range(sc["num_blocks"][i]):
                    b = "%s.branches.%d.%d" % (mp, i, k)
                    conv(b + ".conv1", ch[i], ch[i], 3); bn(b + ".bn1", ch[i])
                    conv(b + ".conv2", ch[i], ch[i], 3); bn(b + ".bn2", ch[i])
            for i in range(nb):
                for j in range(nb):
                    fp = "%s.fuse_layers.%d.%d" % (mp, i, j)
                    if j > i:
                        conv(fp + ".0", ch[i], ch[j], 1); bn(fp + ".1", ch[i])
                    elif j < i:
                        for k in range(i - j):
                            oc = ch[i] if k == i - j - 1 else ch[j]
                            conv("%s.%d.0" % (fp, k), oc, ch[j], 3); bn("%s.%d.1" % (fp, k), oc)
        pre = ch
    high = sum(pre)
    if arch in ("ocrnet.HRNet", "ocrnet.HRNet_Mscale"):
        mid, key_ch, ncls = ocfg["mid_channels"], ocfg["key_channels"], ocfg["num_classes"]
        o = "ocr"
        conv(o + ".conv3x3_ocr.0", mid, high, 3, True); bn(o + ".conv3x3_ocr.1.0", mid)
        ob = o + ".ocr_distri_head.object_context_block"
        conv(ob + ".f_pixel.0", key_ch, mid, 1); bn(ob + ".f_pixel.1.0", key_ch)
        conv(ob + ".f_pixel.2", key_ch, key_ch, 1); bn(ob + ".f_pixel.3.0", key_ch)
        conv(ob + ".f_object.0", key_ch, mid, 1); bn(ob + ".f_object.1.0", key_ch)
        conv(ob + ".f_object.2", key_ch, key_ch, 1); bn(ob + ".f_object.3.0", key_ch)
        conv(ob + ".f_down.0", key_ch, mid, 1); bn(ob + ".f_down.1.0", key_ch)
        conv(ob + ".f_up.0", mid, key_ch, 1); bn(ob + ".f_up.1.0", mid)
        cb = o + ".ocr_distri_head.conv_bn_dropout"
        conv(cb + ".0", mid, 2 * mid, 1); bn(cb + ".1.0", mid)
        conv(o + ".cls_head", ncls, mid, 1, True)
        conv(o + ".aux_head.0", high, high, 1, True); bn(o + ".aux_head.1.0", high)
        conv(o + ".aux_head.2", ncls, high, 1, True)
        if arch == "ocrnet.HRNet_Mscale":
            bot = ocfg["segattn_bot_ch"]
            a = "scale_attn"
            conv(a + ".conv0", bot, mid, 3); bn(a + ".bn0", bot)
            conv(a + ".conv1", bot, bot, 3); bn(a + ".bn1", bot)
            conv(a + ".conv2", 1, bot, 1)
    elif arch == "basic.HRNet":
        bot, ncls = ocfg["segattn_bot_ch"], ocfg["num_classes"]
        s = "seg_head"
        conv(s + ".0", bot, high, 3); bn(s + ".1", bot)
        conv(s + ".3", bot, bot, 3); bn(s + ".4", bot)
        conv(s + ".6", ncls, bot, 1)
    else:
        raise ValueError("unsupported arch %r (hot path covers %s)" % (arch, ", ".join(ARCHS)))
    return out
