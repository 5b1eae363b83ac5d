"""Teacher-forced block-level forward check: each B200 block gets the (bf16-exact) oracle input of that block."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import seg_oracle as O
from b200seg.module import B200SegModule
from b200seg.engine import Engine, Act
from b200seg import model as M, raw

arch = "ocrnet.HRNet_Mscale"
hcfg = O.HRNET_W16_TEST
sd0 = O.synth_state_dict(arch, hcfg, seed=3)
images, gts = O.synth_batch(2, 64, 128, seed=5)
images = images.cuda()


def rel(a, b, name):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    sc = b.abs().max().item()
    l2 = ((a - b).norm() / (b.norm() + 1e-30)).item()
    nbad = int(((a - b).abs() > 0.02 * sc).sum())
    print("%-34s shape %-20s maxerr %.3e scale %.3e relL2 %.5f n>2%% %d" % (name, tuple(b.shape), err, sc, l2, nbad),
          flush=True)


sd = {k: v.clone().cuda() for k, v in sd0.items()}
ctx = O.Ctx(sd, training=True, emulate_bf16=True)
q = ctx.q
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
act = lambda t: Act(t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
nchw = lambda a: a.t.permute(0, 3, 1, 2)

with torch.no_grad():
    p = "backbone"
    x0 = q(images)
    y1 = q(O.conv(ctx, p + ".conv1", x0, 2, 1))
    x16 = Act(raw.image_prep(images, 64, 128), needs_grad=False)
    rec = E.conv_stats(x16, p + ".conv1", p + ".bn1", 3, stride=2)
    rel(rec.y.permute(0, 3, 1, 2), y1, "stem conv1 raw")
    x1 = q(F.relu(O.bn(ctx, p + ".bn1", y1)))
    z1 = E.bn_act(rec, relu=True)
    rel(nchw(z1), x1, "stem conv1+bn+relu")
    x2 = q(F.relu(O.bn(ctx, p + ".bn2", q(O.conv(ctx, p + ".conv2", x1, 2, 1)))))
    z2 = E.conv_bn(act(x1), p + ".conv2", p + ".bn2", 3, stride=2)
    rel(nchw(z2), x2, "stem conv2 (s2)")
    b0 = O.bottleneck(ctx, p + ".layer1.0", x2, True)
    rel(nchw(M.bottleneck(E, p + ".layer1.0", act(x2), True)), b0, "layer1.0 bottleneck+ds")
    b1 = O.bottleneck(ctx, p + ".layer1.1", b0, False)
    rel(nchw(M.bottleneck(E, p + ".layer1.1", act(b0), False)), b1, "layer1.1 bottleneck")
    t0 = q(F.relu(O.bn(ctx, p + ".transition1.0.1", q(O.conv(ctx, p + ".transition1.0.0", b1, 1, 1)))))
    rel(nchw(E.conv_bn(act(b1), p + ".transition1.0.0", p + ".transition1.0.1", 3)), t0, "transition1.0")
    t1 = q(F.relu(O.bn(ctx, p + ".transition1.1.0.1", q(O.conv(ctx, p + ".transition1.1.0.0", b1, 2, 1)))))
    rel(nchw(E.conv_bn(act(b1), p + ".transition1.1.0.0", p + ".transition1.1.0.1", 3, stride=2)), t1, "transition1.1 (s2)")
    bb = O.basic_block(ctx, p + ".stage2.0.branches.0.0", t0)
    rel(nchw(M.basic_block(E, p + ".stage2.0.branches.0.0", act(t0))), bb, "stage2 basic block")
    ys = O.hr_module(ctx, p + ".stage2.0", [t0, t1], [1, 1])
    # reset BN running stats effects don't matter for train-mode outputs
    zs = M.hr_module(E, p + ".stage2.0", [act(t0), act(t1)], [1, 1])
    rel(nchw(zs[0]), ys[0], "stage2 module out0")
    rel(nchw(zs[1]), ys[1], "stage2 module out1")
