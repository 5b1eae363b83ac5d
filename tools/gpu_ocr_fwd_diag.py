import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from test_gpu_blocks import Harness

h = Harness(sharpen_aux=float(sys.argv[1]) if len(sys.argv) > 1 else 6.0)
O, M, E = h.O, h.M, h.E
ctx = h.ctx
high = sum(h.hcfg["stage4"]["num_channels"])
x = h.rand_bf16((2, high, 16, 32), 30).abs()
def rep(a, b, name):
    a, b = a.double().flatten(), b.double().flatten()
    rel = float((a - b).norm() / (b.norm() + 1e-30))
    print("%-20s relL2 %.5f maxerr %.3e scale %.3e" % (name, rel, float((a-b).abs().max()), float(b.abs().max())))
with torch.no_grad():
    p = "ocr"
    f_o = O.bn_relu(ctx, p + ".conv3x3_ocr.1", O.conv(ctx, p + ".conv3x3_ocr.0", x, 1, 1))
    aux_o = O.conv(ctx, p + ".aux_head.2", O.bn_relu(ctx, p + ".aux_head.1", O.conv(ctx, p + ".aux_head.0", x)))
    n, k = aux_o.shape[:2]
    pr = ctx.q(F.softmax(aux_o.reshape(n, k, -1), dim=2))
    context_o = O.spatial_gather(f_o, aux_o, ctx)          # n x c x k x 1
    ob = p + ".ocr_distri_head.object_context_block"
    q_o = O.bn_relu(ctx, ob + ".f_pixel.1", O.conv(ctx, ob + ".f_pixel.0", f_o))
    q_o = O.bn_relu(ctx, ob + ".f_pixel.3", O.conv(ctx, ob + ".f_pixel.2", q_o))
    k_o = O.bn_relu(ctx, ob + ".f_object.1", O.conv(ctx, ob + ".f_object.0", context_o))
    k_o = O.bn_relu(ctx, ob + ".f_object.3", O.conv(ctx, ob + ".f_object.2", k_o))
    v_o = O.bn_relu(ctx, ob + ".f_down.1", O.conv(ctx, ob + ".f_down.0", context_o))
    qq = q_o.reshape(n, 256, -1).permute(0, 2, 1)
    kk = k_o.reshape(n, 256, -1)
    vv = v_o.reshape(n, 256, -1).permute(0, 2, 1)
    sim_o = ctx.q(F.softmax((256 ** -0.5) * torch.matmul(qq, kk), dim=-1))
    c_o = ctx.q(torch.matmul(sim_o, vv))     # n x P x 256
    up_o = O.bn_relu(ctx, ob + ".f_up.1", O.conv(ctx, ob + ".f_up.0", c_o.permute(0, 2, 1).reshape(n, 256, 16, 32)))
E.debug = {}
xa = h.act(x)
cls_r, aux_r, mid_a = M.ocr_block(E, xa, h.ocfg)
d = E.debug
nchw = lambda t: t.permute(0, 3, 1, 2)
rep(nchw(d["feats"]), f_o, "feats")
rep(nchw(aux_r.logits), aux_o, "aux logits")
rep(d["probs"][..., :19].permute(0, 2, 1), pr, "probs")
rep(d["proxy"].view(2, 19, 512).permute(0, 2, 1), context_o.squeeze(3), "proxy/context")
rep(nchw(d["q"]), q_o, "q")
rep(d["k"].view(2, 19, 256).permute(0, 2, 1), k_o.squeeze(3), "k")
rep(d["v"].view(2, 19, 256).permute(0, 2, 1), v_o.squeeze(3), "v")
rep(torch.stack([s[:, :19] for s in d["sim"]]), sim_o, "sim")
rep(d["ctx"].reshape(2, 512, 256), c_o, "attn context")
rep(nchw(d["up"]), up_o, "f_up out")
