"""Forward intermediates: B200 engine vs CPU oracle on one scale pass (training-mode BN)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
import torch.nn.functional as F

from oracle import seg_oracle as O
from b200seg.module import B200SegModule
from b200seg.engine import Engine, Act
from b200seg import model as M, raw

arch = "ocrnet.HRNet_Mscale"
hcfg = O.HRNET_W16_TEST
sd0 = O.synth_state_dict(arch, hcfg, seed=3)
images, gts = O.synth_batch(2, 64, 128, seed=5)


def rel(a, b, name):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item()
    sc = b.abs().max().item()
    l2 = ((a - b).norm() / (b.norm() + 1e-30)).item()
    print("%-28s shape %-22s maxerr %.3e scale %.3e relL2 %.4f" % (name, tuple(b.shape), err, sc, l2), flush=True)


# ---- oracle
ctx = O.Ctx(O.clone_sd(sd0), training=True, emulate_bf16=True)
with torch.no_grad():
    feats_o = O.hrnet_forward(ctx, "backbone", images, hcfg)
    p = "ocr"
    f_o = O.bn_relu(ctx, p + ".conv3x3_ocr.1", O.conv(ctx, p + ".conv3x3_ocr.0", feats_o, 1, 1))
    aux_o = O.conv(ctx, p + ".aux_head.2", O.bn_relu(ctx, p + ".aux_head.1", O.conv(ctx, p + ".aux_head.0", feats_o)))
    context_o = O.spatial_gather(f_o, aux_o, ctx)
    dp = p + ".ocr_distri_head"
    oc_o = O.object_attention(ctx, dp + ".object_context_block", f_o, context_o, 256)
    out_o = O.bn_relu(ctx, dp + ".conv_bn_dropout.1", O.conv(ctx, dp + ".conv_bn_dropout.0", torch.cat([oc_o, f_o], 1)))
    cls_o = O.conv(ctx, p + ".cls_head", out_o)
    attn_o = O.attn_head(ctx, "scale_attn", out_o)

# ---- product
ocfg = dict(O.OCR_CFG)
ocfg["dropout"] = 0.0
net = B200SegModule(arch, 19, hcfg=hcfg, ocfg=ocfg, use_cuda_graph=False)
net.load_state_dict(sd0)
net = net.cuda().train()
net._ensure_device_state()
net._repack()
tensors = {k: v.detach() for k, v in net._tensors().items()}
grads = dict(net._grad_views)
grads["backbone.conv1.weight"] = torch.zeros((64, 16, 3, 3), device="cuda")
E = Engine(tensors, grads, net._packed, True, torch.ones((2, 512), device="cuda"))
x16 = Act(raw.image_prep(images.cuda(), 64, 128), needs_grad=False)
cat = M.hrnet_forward(E, x16, hcfg)
nchw = lambda t: t.permute(0, 3, 1, 2)
rel(nchw(cat.t), feats_o, "hrnet concat feats")
for c0, c1, nm in ((0, 16, "branch0"), (16, 48, "branch1 up"), (48, 112, "branch2 up"), (112, 240, "branch3 up")):
    rel(nchw(cat.t[..., c0:c1]), feats_o[:, c0:c1], "  " + nm)
cls, aux, ocr_feats = M.ocr_block(E, cat, ocfg)
rel(nchw(aux.logits), aux_o, "aux logits")
rel(nchw(ocr_feats.t), out_o, "ocr feats")
rel(nchw(cls.logits), cls_o, "cls logits")
attn = M.attn_head(E, ocr_feats)
rel(torch.sigmoid(nchw(attn.logits)), attn_o, "attn")
