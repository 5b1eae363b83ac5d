"""Op-level parity of the kernels added for DeepLabV3+/WRN-38 against plain PyTorch fp32: dilated
convolutions (forward, data gradient, weight gradient), the residual-sum epilogue with batch statistics, the 1x1
stride-2 projection, 3x3/2 max pooling, channel statistics, per-image spatial sums and the pixel broadcast."""
import pytest
import torch
import torch.nn.functional as F

from test_gpu_ops import BF16_TOL, _setup, close, rnd

pytestmark = pytest.mark.gpu

DIL_CASES = [
    # n, h, w, cin, cout, k, stride, dilation
    (1, 24, 40, 64, 64, 3, 1, 2), (2, 16, 32, 128, 96, 3, 1, 4), (1, 16, 32, 256, 256, 3, 1, 12),
    (1, 16, 32, 512, 256, 3, 1, 24), (1, 16, 32, 512, 256, 3, 1, 36), (1, 33, 47, 64, 128, 3, 1, 2),
    (1, 32, 64, 128, 256, 1, 2, 1), (2, 17, 31, 64, 128, 1, 2, 1),
]


@pytest.mark.parametrize("case", DIL_CASES)
def test_dilated_and_projection_convs(case):
    raw = _setup()
    n, h, w, cin, cout, k, s, dil = case
    pad = dil if k == 3 else 0
    x = rnd((n, h, w, cin), 1)
    wt = rnd((cout, cin, k, k), 2, scale=(cin * k * k) ** -0.5, dtype=torch.float32).contiguous()
    w_f, w_d = raw.pack_weight(wt)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, stride=s, padding=pad, dilation=dil)
    y, stats = raw.conv2d_fwd(x, w_f, None, stride=s, emit_stats=True, dilation=dil)
    assert tuple(y.shape) == (n, ref.shape[2], ref.shape[3], cout)
    close(y, ref.permute(0, 2, 3, 1), BF16_TOL, "fwd")
    buf, grid, cpad = stats
    tot = buf[: grid * 2 * cpad].view(grid, 2, cpad).sum(0)[:, :cout]
    yr = y.float().reshape(-1, cout)
    close(tot[0], yr.sum(0), 1e-3 * max(1.0, yr.shape[0] ** 0.5), "stats sum")
    close(tot[1], (yr * yr).sum(0), 1e-3, "stats sumsq")
    ho, wo = y.shape[1:3]
    dy = rnd((n, ho, wo, cout), 4)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dx = raw.conv2d_dgrad(dy, w_d, (n, h, w, cin), k, s, dilation=dil)
    close(dx, xr.grad.permute(0, 2, 3, 1), BF16_TOL, "dgrad")
    add = rnd((n, h, w, cin), 5)
    dx2 = raw.conv2d_dgrad(dy, w_d, (n, h, w, cin), k, s, addend=add.clone(), dilation=dil)
    close(dx2, xr.grad.permute(0, 2, 3, 1) + add.float(), BF16_TOL, "dgrad+addend")
    dw = torch.zeros((cout, k * k, cin), dtype=torch.float32, device="cuda")
    raw.conv2d_wgrad(x, dy, dw, cout, k, s, dilation=dil)
    close(raw.ohwi_to_oihw(dw, k), wr.grad, 2e-3, "wgrad")


@pytest.mark.parametrize("case", [(1, 16, 32, 64, 64, 3, 1), (2, 24, 40, 128, 128, 3, 2), (1, 16, 32, 256, 512, 1, 1),
                                  (1, 16, 32, 128, 128, 3, 4)])
def test_residual_sum_epilogue_with_statistics(case):
    """y = conv(x) + shortcut in the convolution epilogue, statistics of the STORED sum (IdentityResidualBlock)."""
    raw = _setup()
    n, h, w, cin, cout, k, dil = case
    x = rnd((n, h, w, cin), 1)
    short = rnd((n, h, w, cout), 6)
    wt = rnd((cout, cin, k, k), 2, scale=(cin * k * k) ** -0.5, dtype=torch.float32).contiguous()
    w_f, _ = raw.pack_weight(wt)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), None, padding=dil if k == 3 else 0,
                   dilation=dil).permute(0, 2, 3, 1) + short.float()
    y, stats = raw.conv2d_fwd(x, w_f, None, emit_stats=True, dilation=dil, addend=short)
    close(y, ref, BF16_TOL, "conv+addend")
    buf, grid, cpad = stats
    tot = buf[: grid * 2 * cpad].view(grid, 2, cpad).sum(0)[:, :cout]
    yr = y.float().reshape(-1, cout)
    close(tot[0], yr.sum(0), 1e-3 * max(1.0, yr.shape[0] ** 0.5), "stats sum")
    close(tot[1], (yr * yr).sum(0), 1e-3, "stats sumsq")
    # writing into a channel slice of a wider buffer (decoder / ASPP concatenations)
    wide = torch.zeros((n, h, w, cout + 64), dtype=torch.bfloat16, device="cuda")
    raw.conv2d_fwd(x, w_f, None, dilation=dil, addend=short, out=wide[..., 64:])
    assert torch.equal(wide[..., 64:], y) and float(wide[..., :64].abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(1, 32, 64, 64), (2, 33, 47, 128), (1, 17, 9, 256)])
def test_maxpool_3x3_stride2(shape):
    raw = _setup()
    n, h, w, c = shape
    # a coarse value grid forces ties inside windows: both implementations route the gradient to the FIRST maximum in
    # (kh, kw) order (ATen's rule), so the comparison stays element-wise
    x = (rnd(shape, 3, scale=4.0).float().round() / 2).to(torch.bfloat16)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(xr, 3, stride=2, padding=1)
    y = raw.maxpool3x3s2(x)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))
    dy = rnd(tuple(y.shape), 4)
    dx = raw.maxpool3x3s2_bwd(x, dy)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    close(dx, xr.grad.permute(0, 2, 3, 1), BF16_TOL, "maxpool backward")
    acc = rnd(tuple(x.shape), 7)
    dx2 = raw.maxpool3x3s2_bwd(x, dy, out=acc.clone(), accumulate=True)
    close(dx2, dx.float() + acc.float(), BF16_TOL, "accumulate")


@pytest.mark.parametrize("shape", [(2, 16, 32, 64), (1, 33, 47, 128), (2, 1, 1, 4096), (1, 64, 128, 256)])
def test_channel_stats_spatial_sum_and_broadcast(shape):
    raw = _setup()
    n, h, w, c = shape
    x = rnd(shape, 1)
    if c <= 2048:                       # the statistics kernel serves the pooled trunk maps (64 / 128 channels)
        buf, grid, cpad = raw.channel_stats(x)
        tot = buf[: grid * 2 * cpad].view(grid, 2, cpad).sum(0)
        xr = x.float().reshape(-1, c)
        close(tot[0], xr.sum(0), 1e-3 * max(1.0, xr.shape[0] ** 0.5), "sum")
        close(tot[1], (xr * xr).sum(0), 1e-3, "sumsq")
    s = raw.spatial_sum(x, 1.0 / (h * w))
    assert tuple(s.shape) == (n, 1, 1, c)
    close(s.view(n, c), x.float().mean((1, 2)), BF16_TOL, "spatial mean")
    acc = rnd((n, 1, 1, c), 2)
    s2 = raw.spatial_sum(x, 0.5, out=acc.clone(), accumulate=True)
    close(s2.view(n, c), 0.5 * x.float().sum((1, 2)) + acc.float().view(n, c), BF16_TOL, "spatial sum accumulate")
    v = rnd((n, 1, 1, c), 3)
    wide = torch.zeros((n, h, w, c + 32), dtype=torch.bfloat16, device="cuda")
    o = raw.broadcast_pixels(v, h, w, out=wide[..., 32:])
    assert torch.equal(o, v.expand(n, h, w, c)) and float(wide[..., :32].abs().max()) == 0.0
    base = rnd(shape, 4)
    o2 = raw.broadcast_pixels(v, h, w, out=base.clone(), scale=0.25, accumulate=True)
    close(o2, base.float() + 0.25 * v.float(), BF16_TOL, "broadcast accumulate")
