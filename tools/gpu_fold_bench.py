"""Times the end-of-step gradient fold in isolation (after a real step, again on cleared accumulators, single accumulator)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch
from b200seg.module import B200SegModule
from b200seg import raw
net = B200SegModule("ocrnet.HRNet_Mscale", 19, use_cuda_graph=False).cuda().train()
images = torch.randn(1, 3, 256, 512, device="cuda")
gts = torch.randint(0, 19, (1, 256, 512), device="cuda")
loss = net({"images": images, "gts": gts})
torch.cuda.synchronize()
def t(fn, n=3):
    out = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1) * 1000))
    return out
g = torch.Generator(device="cuda").manual_seed(0)
net._acc_hi.normal_(generator=g); net._acc_lo.normal_(generator=g)
print("fold both (random data)      us:", t(lambda: raw.grad_fold(net._flat_grad, net._acc_hi, net._acc_lo, net._fold_table)))
print("fold both (zeros)            us:", t(lambda: raw.grad_fold(net._flat_grad, net._acc_hi, net._acc_lo, net._fold_table)))
print("fold hi only                 us:", t(lambda: raw.grad_fold(net._flat_grad, net._acc_hi, None, net._fold_table)))
print("fold both, no clear          us:", t(lambda: raw.grad_fold(net._flat_grad, net._acc_hi, net._acc_lo, net._fold_table, clear=False)))
print("memset one accumulator       us:", t(lambda: net._acc_hi.zero_()))
print("accum_f32 reference          us:", t(lambda: raw.accum_f32(net._flat_grad, net._acc_hi)))
print("blocks", net._fold_table["n_blocks"], "numel", net._flat_grad.numel())
