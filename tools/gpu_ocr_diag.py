import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from test_gpu_blocks import Harness

h = Harness(sharpen_aux=6.0)
O, M, E = h.O, h.M, h.E
high = sum(h.hcfg["stage4"]["num_channels"])
x = h.rand_bf16((2, high, 16, 32), 30).abs().requires_grad_(True)
cls, aux, mid = O.ocr_block(h.ctx, "ocr", x, h.ocfg)
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
d_cls = h.rand_bf16(cls.shape, 31, 0.1) * (mode in ("all", "cls"))
d_aux = h.rand_bf16(aux.shape, 32, 0.1) * (mode in ("all", "aux"))
d_mid = h.rand_bf16(mid.shape, 33, 0.1) * (mode in ("all", "mid"))
torch.autograd.backward([cls, aux, mid], [d_cls, d_aux, d_mid])
xa = h.act(x)
cls_r, aux_r, mid_a = M.ocr_block(E, xa, h.ocfg)
pad = lambda t: F.pad(t.permute(0, 2, 3, 1), (0, 32 - t.shape[1])).contiguous().to(torch.bfloat16)
cls_r.dlogits, aux_r.dlogits = pad(d_cls), pad(d_aux)
mid_a.grad = h.nhwc(d_mid)
E.run_backward()
def rep(a, b, name):
    a, b = a.double().flatten(), b.double().flatten()
    rel = float((a - b).norm() / (b.norm() + 1e-30)); c = float((a @ b) / (a.norm() * b.norm() + 1e-30))
    print("%-70s relL2 %.4f cos %.5f |ref| %.3e" % (name, rel, c, float(b.norm())))
print("mode", mode)
rep(xa.grad.permute(0, 3, 1, 2), x.grad, "dx")
for name, v in h.sd.items():
    if name.startswith("ocr.") and v.grad is not None and name in h.grads:
        rep(h.grads[name], v.grad, name)
