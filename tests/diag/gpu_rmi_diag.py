"""RMI head debug: pooled maps, pooled-probability gradient and final logit gradient against torch autograd (fp64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch, torch.nn.functional as F
from b200seg import raw
n, H, W = 2, 64, 96
hq, wq = H // 4, W // 4
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.randn(s, generator=g, device="cuda")
hi_cls = torch.zeros((n, hq, wq, 20), device="cuda"); hi_aux = torch.zeros((n, hq, wq, 20), device="cuda")
hi_cls[..., :19], hi_aux[..., :19] = mk(n, hq, wq, 19) * 2, mk(n, hq, wq, 19) * 2
gts = torch.randint(0, 19, (n, H, W), generator=g, device="cuda"); gts[:, :5] = 255; gts[:, 20:30, 40:50] = 3
d = raw.mscale_desc(n, H, W, hq, wq, 0, 0, 0, 0, 2, 1.0, 0.4, 0.0, loss_kind=1)
leaf = hi_cls[..., :19].permute(0, 3, 1, 2).clone().requires_grad_(True)
logits = F.interpolate(leaf, size=(H, W), mode="bilinear", align_corners=False)
mask = gts < 19
onehot = F.one_hot(gts * mask, 19).float() * mask.unsqueeze(3)
probs = torch.sigmoid(logits) * mask.unsqueeze(1) + 1e-6
la = F.avg_pool2d(onehot.permute(0, 3, 1, 2), 4, 4, 2)
pr = F.avg_pool2d(probs, 4, 4, 2)
pr.retain_grad()
hp, wp = la.shape[2:]
nh, nw = hp - 2, wp - 2
la_v = torch.stack([la[:, :, y:y + nh, x:x + nw] for y in range(3) for x in range(3)], 2).reshape(n, 19, 9, -1).double()
pr_v = torch.stack([pr[:, :, y:y + nh, x:x + nw] for y in range(3) for x in range(3)], 2).reshape(n, 19, 9, -1).double()
eye = torch.eye(9, dtype=torch.float64, device="cuda")[None, None]
la_v = la_v - la_v.mean(3, keepdim=True); pr_v = pr_v - pr_v.mean(3, keepdim=True)
pc = pr_v @ pr_v.transpose(2, 3); lp = la_v @ pr_v.transpose(2, 3); lc = la_v @ la_v.transpose(2, 3)
appro = lc - (lp @ torch.inverse(pc + eye * 5e-4)) @ lp.transpose(-2, -1)
rmi = torch.sum(torch.log(torch.diagonal(torch.linalg.cholesky(appro + eye * 5e-4), dim1=-2, dim2=-1) + 1e-8), -1)
loss_rmi = 0.5 * (rmi.view(-1, 19).mean(0).float() / 9.0).sum()
loss_rmi.backward()
dpr, terms = raw.rmi_head(d, gts, hi_cls, None)
# our pooled maps: recompute via kernel call
L = raw.lib(); import ctypes
prp = torch.empty((n, hp, wp, 20), device="cuda"); lap = torch.empty((n, hp, wp, 20), device="cuda")
assert L.b200seg_rmi_pool(ctypes.byref(d), raw.ptr(gts), raw.ptr(hi_cls), None, raw.ptr(prp), raw.ptr(lap), raw.stream_ptr()) == 0
def rep(a, b, name):
    a, b = a.double(), b.double()
    print("%-28s rel %.3e  max|ref| %.3e  max err %.3e" % (name, float((a - b).norm() / (b.norm() + 1e-30)), float(b.abs().max()), float((a - b).abs().max())))
rep(prp[..., :19].permute(0, 3, 1, 2), pr.detach(), "pr_pool")
rep(lap[..., :19].permute(0, 3, 1, 2), la, "la_pool")
rep(terms.sum(), loss_rmi.detach(), "rmi term")
rep(dpr[..., :19].permute(0, 3, 1, 2), pr.grad, "d loss / d pr_pool")
print("sample ours", dpr[0, 5, 5, :4].tolist(), "ref", pr.grad[0, :4, 5, 5].tolist())
print("sample ours", dpr[0, 0, 0, :4].tolist(), "ref", pr.grad[0, :4, 0, 0].tolist())
# ---- per-pixel gradient of the full criterion on head 0 (single scale: Ghi[:, :19] == d loss / d full-res logits)
from oracle import seg_oracle as O
leaf2 = hi_cls[..., :19].permute(0, 3, 1, 2).clone().requires_grad_(True)
full = F.interpolate(leaf2, size=(H, W), mode="bilinear", align_corners=False)
full.retain_grad()
tot = O.rmi_loss(full, gts, do_rmi=True)
tot.backward()
inv = raw.count_valid(gts, plus_one=True)
d1 = raw.mscale_desc(n, H, W, hq, wq, 0, 0, 0, 0, 1, 1.0, 0.4, 0.0, loss_kind=1)
dpr1, terms1 = raw.rmi_head(d1, gts, hi_cls, None)
loss, g_hi, g_lo, g_sup = raw.mscale_loss_fwd(d1, gts, inv, hi_cls, None, None, None, dpr1, terms1)
G = g_hi.view(n, H, W, 40)[..., :19].float().permute(0, 3, 1, 2)
rep(loss[0], tot.detach(), "total loss")
rep(G, full.grad, "per-pixel grad (bce+rmi)")
loss_b, g_b, _, _ = raw.mscale_loss_fwd(d1, gts, inv, hi_cls, None, None, None, None, None)
Gb = g_b.view(n, H, W, 40)[..., :19].float().permute(0, 3, 1, 2)
fb = F.interpolate(leaf2.detach(), size=(H, W), mode="bilinear", align_corners=False).requires_grad_(True)
(0.5 * O.rmi_loss(fb, gts, do_rmi=False)).backward()
rep(Gb, fb.grad, "per-pixel grad (bce only)")
rep(G - Gb, full.grad - fb.grad, "per-pixel grad (rmi only)")
idx = (G - full.grad).abs().flatten().argmax().item()
print("worst index", idx, "ours", G.flatten()[idx].item(), "ref", full.grad.flatten()[idx].item(), "bce ours", Gb.flatten()[idx].item(), "bce ref", fb.grad.flatten()[idx].item())
d_cls, _ = raw.mscale_hi_bwd(d1, g_hi)
rep(d_cls[..., :19].float().permute(0, 3, 1, 2), leaf2.grad, "d hi cls")
nn_ = idx // (19 * H * W); r_ = idx % (19 * H * W); cc_ = r_ // (H * W); r2 = r_ % (H * W); Y_ = r2 // W; X_ = r2 % W
ci_, cj_ = (Y_ + 2) // 4, (X_ + 2) // 4
pz = torch.sigmoid(full.detach()[nn_, cc_, Y_, X_]).item()
print("pixel", nn_, cc_, Y_, X_, "cell", ci_, cj_, "label", int(gts[nn_, Y_, X_]))
print("dpr1", dpr1[nn_, ci_, cj_, cc_].item(), "dpr(first call)", dpr[nn_, ci_, cj_, cc_].item(), "pr.grad", pr.grad[nn_, cc_, ci_, cj_].item(), "p(1-p)", pz * (1 - pz))
print("expected rmi px grad", pr.grad[nn_, cc_, ci_, cj_].item() / 16 * pz * (1 - pz), "ours", (G - Gb)[nn_, cc_, Y_, X_].item(), "ref", (full.grad - fb.grad)[nn_, cc_, Y_, X_].item())
rep(dpr1, dpr, "dpr second call vs first")
# ---- manual graph on exactly the same full-resolution logits as the oracle call above
z = full.detach().clone().requires_grad_(True)
probs = torch.sigmoid(z) * mask.unsqueeze(1) + 1e-6
pr2 = F.avg_pool2d(probs, 4, 4, 2); pr2.retain_grad()
pv = torch.stack([pr2[:, :, y:y + nh, x:x + nw] for y in range(3) for x in range(3)], 2).reshape(n, 19, 9, -1).double()
lv = torch.stack([la[:, :, y:y + nh, x:x + nw] for y in range(3) for x in range(3)], 2).reshape(n, 19, 9, -1).double()
lv = lv - lv.mean(3, keepdim=True); pv = pv - pv.mean(3, keepdim=True)
pc2 = pv @ pv.transpose(2, 3); lp2 = lv @ pv.transpose(2, 3); lc2 = lv @ lv.transpose(2, 3)
ap2 = lc2 - (lp2 @ torch.inverse(pc2 + eye * 5e-4)) @ lp2.transpose(-2, -1)
rm2 = torch.sum(torch.log(torch.diagonal(torch.linalg.cholesky(ap2 + eye * 5e-4), dim1=-2, dim2=-1) + 1e-8), -1)
(0.5 * (rm2.view(-1, 19).mean(0).float() / 9).sum()).backward()
rep(z.grad, full.grad - fb.grad, "manual px grad vs oracle px grad")
rep(pr2.grad, pr.grad, "manual pr.grad (2nd) vs (1st)")
z3 = full.detach().clone().requires_grad_(True)
t3 = O.rmi_loss(z3, gts, do_rmi=True); t3.backward()
rep(z3.grad, full.grad, "oracle grad leaf vs non-leaf")
print("full is contiguous", full.is_contiguous(), "full.grad contiguous", full.grad.is_contiguous(), full.grad.shape, full.grad.stride())
