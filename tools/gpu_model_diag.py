"""Per-parameter gradient agreement (cosine / relative norm) between the B200 module and the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch

from oracle import seg_oracle as O
from b200seg.module import B200SegModule

arch = sys.argv[1] if len(sys.argv) > 1 else "ocrnet.HRNet_Mscale"
hcfg = O.HRNET_W16_TEST
sd0 = O.synth_state_dict(arch, hcfg, seed=3)
images, gts = O.synth_batch(2, 64, 128, seed=5)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
EMU = os.environ.get("EMULATE_BF16", "1") == "1"
sd = {k: v.clone().cuda() for k, v in sd0.items()}
for k, v in sd.items():
    if v.is_floating_point() and "running" not in k:
        v.requires_grad_(True)
ctx = O.Ctx(sd, training=True, emulate_bf16=EMU)
if arch == "ocrnet.HRNet_Mscale":
    loss = O.mscale_two_scale(ctx, images.cuda(), gts.cuda(), hcfg=hcfg)
elif arch == "ocrnet.HRNet":
    loss = O.ocrnet_forward(ctx, images.cuda(), gts.cuda(), hcfg=hcfg)
else:
    loss = O.basic_forward(ctx, images.cuda(), gts.cuda(), hcfg=hcfg)
loss.backward()
sd = {k: v.cpu() if v.grad is None else v for k, v in sd.items()}
ocfg = dict(O.OCR_CFG)
ocfg["dropout"] = 0.0
net = B200SegModule(arch, 19, hcfg=hcfg, ocfg=ocfg, use_cuda_graph=False)
net.load_state_dict(sd0)
net = net.cuda().train()
l2 = net({"images": images.cuda(), "gts": gts.cuda()})
l2.backward()
torch.cuda.synchronize()
print("loss oracle %.6f b200 %.6f terms %s" % (float(loss), float(l2), net.last_loss_terms.tolist()))
for name, p in net.named_parameters():
    g_ref = sd[name].grad
    if g_ref is None:
        print("%-70s oracle grad None, ours |g| %.3e" % (name, p.grad.norm().item()))
        continue
    a, b = p.grad.cpu().double().flatten(), g_ref.cpu().double().flatten()
    c = float((a @ b) / (a.norm() * b.norm() + 1e-30))
    print("%-70s cos %.4f  |ours|/|ref| %.3f  |ref| %.3e" % (name, c, float(a.norm() / (b.norm() + 1e-30)), float(b.norm())))
