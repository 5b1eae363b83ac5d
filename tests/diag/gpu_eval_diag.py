"""Layer-by-layer eval-mode comparison of the B200 path against the bf16-emulating oracle (debug aid).
usage: python tests/diag/gpu_eval_diag.py [arch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
from oracle import seg_oracle as O
from b200seg.module import B200SegModule
from b200seg import engine as EN, model as M

arch = sys.argv[1] if len(sys.argv) > 1 else "ocrnet.HRNet"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
hcfg = O.HRNET_W16_TEST
sd0 = O.synth_state_dict(arch, hcfg, seed=3)
g = torch.Generator().manual_seed(9)
for k in sd0:
    if k.endswith("running_mean"):
        sd0[k] = 0.1 * torch.randn(sd0[k].shape, generator=g)
    elif k.endswith("running_var"):
        sd0[k] = 0.5 + torch.rand(sd0[k].shape, generator=g)
if len(sys.argv) > 2 and sys.argv[2] == "cond":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_model import _condition_eval_weights
    _condition_eval_weights(sd0)
images, _ = O.synth_batch(2, 64, 128, seed=5)
sd = {k: v.clone().cuda() for k, v in sd0.items()}
ctx = O.Ctx(sd, training=False, emulate_bf16=True)

ref = {}
o_conv, o_bn = O.conv, O.bn
def conv_hook(c, name, x, stride=1, padding=0):
    y = o_conv(c, name, x, stride, padding)
    ref["conv:" + name] = y
    return y
def bn_hook(c, name, x):
    y = o_bn(c, name, x)
    ref["bn:" + name] = y
    return y
O.conv, O.bn = conv_hook, bn_hook
o_sg = O.spatial_gather
def sg_hook(feats, probs, c=None):
    y = o_sg(feats, probs, c)
    ref["gather"] = y            # n x c x k x 1
    ref["gather_feats"] = feats
    return y
O.spatial_gather = sg_hook
with torch.no_grad():
    if arch == "ocrnet.HRNet":
        r = O.ocrnet_forward(ctx, images.cuda(), hcfg=hcfg)
    elif arch == "basic.HRNet":
        r = O.basic_forward(ctx, images.cuda(), hcfg=hcfg)
    else:
        r = O.mscale_two_scale(ctx, images.cuda(), hcfg=hcfg)

ours = {}
E_conv_stats, E_conv_head, E_bn_act = EN.Engine.conv_stats, EN.Engine.conv_head, EN.Engine.bn_act
def cs(self, x, cname, bname, ksize, stride=1, bias=False):
    rec = E_conv_stats(self, x, cname, bname, ksize, stride, bias)
    ours.setdefault("conv:" + cname, rec.y.float().permute(0, 3, 1, 2))
    return rec
def ch(self, x, cname, bias=True, ld=20):
    rec = E_conv_head(self, x, cname, bias, ld)
    ours.setdefault("conv:" + cname, rec.logits.float().permute(0, 3, 1, 2))
    return rec
EN.Engine.conv_stats, EN.Engine.conv_head = cs, ch
from b200seg import raw as RAW
m_sg, r_ss = M.spatial_gather, RAW.spatial_softmax_fwd
def sg2(E, feats, aux, K):
    pr = m_sg(E, feats, aux, K)
    ours["gather"] = pr.t.float().permute(0, 3, 1, 2)      # n,K,1,C -> n,C,K,1
    ours["gather_feats"] = feats.t.float().permute(0, 3, 1, 2)
    return pr
def ss2(logits, K):
    p_ = r_ss(logits, K)
    ours["probs"] = p_.float()
    return p_
M.spatial_gather, RAW.spatial_softmax_fwd = sg2, ss2
net = B200SegModule(arch, 19, hcfg=hcfg)
net.load_state_dict(sd0)
net = net.cuda().eval()
out = net({"images": images.cuda()})
n = 0
for k, b in ref.items():
    if not k.startswith("conv:") or k not in ours:
        continue
    a = ours[k]
    if a.shape != b.shape:
        if a.numel() == b.numel():
            a = a.reshape(b.shape) if a.shape[1:] == b.shape[1:] else a.permute(0, 2, 1, 3).reshape(b.shape) if False else a
        if a.shape != b.shape:
            print("%-70s shape %s vs %s" % (k, tuple(a.shape), tuple(b.shape)))
            continue
    rel = float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
    flag = " <<<<" if rel > 0.02 else ""
    if rel > 0.008 or "ocr" in k or "seg_head" in k or "scale_attn" in k:
        print("%-70s rel %.4f |ref| %.3e%s" % (k, rel, float(b.abs().max()), flag))
import torch.nn.functional as F
for k in ("gather", "gather_feats"):
    a, b = ours[k].double(), ref[k].double()
    print("%-20s rel %.4f  |ref| %.3e  mean ours %.5f ref %.5f" % (k, float((a - b).norm() / b.norm()), float(b.abs().max()), float(a.mean()), float(b.mean())))
aux = ref["conv:ocr.aux_head.2"]
n_, k_ = aux.shape[:2]
pr = F.softmax(aux.reshape(n_, k_, -1), dim=2).permute(0, 2, 1)          # n, P, K
po = ours["probs"][..., :19].double()
print("probs rel %.5f ; sum over pixels ours %s ref %s" % (float((po - pr.double()).norm() / pr.double().norm()), po.sum(1)[0, :4].tolist(), pr.sum(1)[0, :4].tolist()))
print("probs pad cols max", float(ours["probs"][..., 19:].abs().max()))
ft = ref["gather_feats"].reshape(n_, ref["gather_feats"].shape[1], -1).permute(0, 2, 1).double()   # n,P,C
mine = torch.matmul(po.permute(0, 2, 1), ours["gather_feats"].reshape(n_, ft.shape[2], -1).permute(0, 2, 1).double())  # n,K,C
print("gather recomputed from OUR probs/feats vs our kernel: rel %.5f" % float((mine.permute(0, 2, 1).unsqueeze(3) - ours["gather"].double()).norm() / mine.norm()))
for k in r:
    a, b = out[k].double(), r[k].double()
    print("OUT %-20s rel %.4f" % (k, float((a - b).norm() / (b.norm() + 1e-30))))
