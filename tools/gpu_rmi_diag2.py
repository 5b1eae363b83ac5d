import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from oracle import seg_oracle as O
torch.manual_seed(0)
for dev in ("cpu", "cuda"):
    torch.manual_seed(0)
    n, H, W = 2, 64, 96
    leaf = (torch.randn(n, 19, H // 4, W // 4) * 2).to(dev)
    gts = torch.randint(0, 19, (n, H, W)); gts[:, :5] = 255; gts[:, 20:30, 40:50] = 3
    gts = gts.to(dev)
    full = F.interpolate(leaf, size=(H, W), mode="bilinear", align_corners=False).requires_grad_(True)
    tot = O.rmi_loss(full, gts, do_rmi=True); tot.backward()
    fb = full.detach().clone().requires_grad_(True)
    (0.5 * O.rmi_loss(fb, gts, do_rmi=False)).backward()
    ref = full.grad - fb.grad
    z = full.detach().clone().requires_grad_(True)
    mask = (gts < 19)
    probs = torch.sigmoid(z) * mask.unsqueeze(1) + 1e-6
    pr = F.avg_pool2d(probs, 4, 4, 2); pr.retain_grad()
    onehot = F.one_hot(gts * mask, 19).float() * mask.unsqueeze(3)
    la = F.avg_pool2d(onehot.permute(0, 3, 1, 2), 4, 4, 2)
    hp, wp = la.shape[2:]; nh, nw = hp - 2, wp - 2
    la_v = torch.stack([la[:, :, y:y + nh, x:x + nw] for y in range(3) for x in range(3)], 2).reshape(n, 19, 9, -1).double()
    pr_v = torch.stack([pr[:, :, y:y + nh, x:x + nw] for y in range(3) for x in range(3)], 2).reshape(n, 19, 9, -1).double()
    eye = torch.eye(9, dtype=torch.float64, device=dev)[None, None]
    la_v = la_v - la_v.mean(3, keepdim=True); pr_v = pr_v - pr_v.mean(3, keepdim=True)
    pc = pr_v @ pr_v.transpose(2, 3); lp = la_v @ pr_v.transpose(2, 3); lc = la_v @ la_v.transpose(2, 3)
    inv = torch.inverse(pc + eye * 5e-4); inv.retain_grad()
    appro = lc - (lp @ inv) @ lp.transpose(-2, -1); appro.retain_grad()
    ch = torch.linalg.cholesky(appro + eye * 5e-4)
    rmi = torch.sum(torch.log(torch.diagonal(ch, dim1=-2, dim2=-1) + 1e-8), -1)
    loss = 0.5 * (rmi.view(-1, 19).mean(0).float() / 9).sum()
    loss.backward()
    p = torch.sigmoid(z.detach())
    ci = (torch.arange(H, device=dev) + 2) // 4; cj = (torch.arange(W, device=dev) + 2) // 4
    form = pr.grad[:, :, ci][:, :, :, cj] / 16 * p * (1 - p) * mask.unsqueeze(1)
    print(dev, "manual-autograd vs oracle:", (z.grad - ref).abs().max().item(), "formula vs manual-autograd:", (form - z.grad).abs().max().item(), "scale", ref.abs().max().item())
    torch.save(dict(zg=z.grad.cpu(), ref=ref.cpu(), prg=pr.grad.cpu(), ag=appro.grad.cpu(), ig=inv.grad.cpu(), loss=loss.item(), tot=tot.item()), os.path.join(ROOT, "gpurun_out", "rmi_%s.pt" % dev))
a = torch.load(os.path.join(ROOT, "gpurun_out", "rmi_cpu.pt")); b = torch.load(os.path.join(ROOT, "gpurun_out", "rmi_cuda.pt"))
for k in ("zg", "ref", "prg", "ag", "ig"):
    print(k, "cpu vs cuda max diff", (a[k] - b[k]).abs().max().item(), "scale", a[k].abs().max().item())
print("loss", a["loss"], b["loss"], a["tot"], b["tot"])
