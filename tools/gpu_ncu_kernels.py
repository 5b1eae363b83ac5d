"""Launches the HBM-bound hot kernels of the 48-channel / 256x512 branch once each (after one warm-up launch) for an
`ncu --set full` capture:  conv3x3_halo (forward, BN statistics), its data gradient, wgrad_igemm + wgrad_reduce,
bn_apply, bn_bwd_reduce, bn_bwd_apply (deferred BatchNorm finalisation, the training default), and the three kernels of
the device input pipeline on a 1024x2048 frame.

  ncu --set full --clock-control none --import-source on -o gpurun_out/r2_kernels \
      -k regex:"conv3x3_halo|wgrad_|bn_bwd_|bn_apply|aug_" python tools/gpu_ncu_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import torch  # noqa: E402

from b200seg import raw  # noqa: E402

h, w, c = 256, 512, 48
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((1, h, w, c), generator=g, device="cuda").to(torch.bfloat16)
dz = torch.randn((1, h, w, c), generator=g, device="cuda").to(torch.bfloat16)
wt = torch.randn((c, c, 3, 3), generator=g, device="cuda") * 0.05
w_f, w_d = raw.pack_weight(wt)
gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
dgamma, dbeta = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
dw = torch.zeros((c, 9, c), device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
cpad = (c + 15) // 16 * 16
for it in range(2):
    # the training default: deferred BatchNorm finalisation (statistics cells folded by the consuming apply passes)
    cells_f = torch.zeros(2 * cpad, dtype=torch.float64, device="cuda")
    cells_b = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    par = torch.empty((4, c), device="cuda")
    flush.zero_()                      # cold L2 for the measured launch, like inside a step whose working set is > 4 GB
    y = raw.conv2d_fwd_cells(x, w_f, None, 1, cells_f)
    flush.zero_()
    z = raw.bn_apply_cells(y, cells_f, par, gamma, beta, 1e-5, 0.1, x, None, True, batch_out=torch.zeros(2 * c, device="cuda"))
    flush.zero_()
    dy = raw.bn_bwd(dz, z, None, y, par[2], par[3], gamma, dgamma, dbeta, g_out=torch.empty_like(dz), cells=cells_b)
    flush.zero_()
    raw.conv2d_wgrad(x, dy, dw, c, 3, 1)
    flush.zero_()
    dx = raw.conv2d_dgrad(dy, w_d, (1, h, w, c), 3, 1)
torch.cuda.synchronize()
# the device input pipeline on a Cityscapes-size frame
import random  # noqa: E402

import numpy as np  # noqa: E402

from b200seg import augment as AUG  # noqa: E402

t = AUG.DeviceTrainTransform((1024, 2048))
img_d = torch.randint(0, 256, (1024, 2048, 3), dtype=torch.uint8, device="cuda")       # a decoded frame's worth of bytes
mask_d = torch.randint(0, 34, (1024, 2048), dtype=torch.uint8, device="cuda")
random.seed(0)
np.random.seed(0)
for it in range(2):
    p = t.draw(2048, 1024)
    flush.zero_()
    t(img_d, mask_d, params=p)
torch.cuda.synchronize()
print("done")
