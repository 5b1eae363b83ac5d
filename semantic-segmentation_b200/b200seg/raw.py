"""Functional (no-autograd) wrappers: torch tensors in, C-ABI calls on the current CUDA stream, torch tensors out.

Tensors are NHWC ("pixels x channels") bf16 unless stated. PyTorch only provides device memory and streams here.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvDesc, check, lib, ptr, stream_ptr


def pack_weight(w_oihw, want_dgrad=True):
    """fp32 OIHW master weight -> (bf16 [O][k*k][I], bf16 [I][k*k flipped][O])."""
    assert w_oihw.is_cuda and w_oihw.dtype == torch.float32 and w_oihw.is_contiguous()
    o, i, k, _ = w_oihw.shape
    w_f = torch.empty((o, k * k, i), dtype=torch.bfloat16, device=w_oihw.device)
    w_d = torch.empty((i, k * k, o), dtype=torch.bfloat16, device=w_oihw.device) if want_dgrad else None
    check(lib().b200seg_pack_weight(ptr(w_oihw), o, i, k, ptr(w_f), ptr(w_d), stream_ptr()), "pack_weight")
    return w_f, w_d


def _conv_desc(x, cout, ksize, stride, y_ld, out_fp32, has_bias, emit_stats, force_kc=0, cin=None):
    n, h, w, _ = x.shape
    d = ConvDesc()
    d.n, d.h, d.w = n, h, w
    d.cin = x.shape[3] if cin is None else cin
    d.cout = cout
    d.ksize, d.stride, d.pad = ksize, stride, (1 if ksize == 3 else 0)
    d.x_ld = x.stride(2)
    d.y_ld = y_ld
    d.out_fp32, d.has_bias, d.emit_stats, d.reserved = int(out_fp32), int(has_bias), int(emit_stats), force_kc
    return d


def conv2d_fwd(x, w_ohwi, bias=None, stride=1, out=None, out_fp32=False, emit_stats=False, force_kc=0,
               direct=False):
    """x: [N,H,W,Cin] bf16 view (channel stride 1, pixel pitch x.stride(2)); w_ohwi: [Cout][k*k][Cin] bf16.
    Returns y [N,Ho,Wo,Cout] (bf16|fp32) and, with emit_stats, the per-CTA partials [grid][2][cout_pad] fp32."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.stride(3) == 1
    cout, taps, cin = w_ohwi.shape
    assert cin == x.shape[3]
    ksize = 3 if taps == 9 else 1
    n, h, w, _ = x.shape
    pad = 1 if ksize == 3 else 0
    ho = (h + 2 * pad - ksize) // stride + 1
    wo = (w + 2 * pad - ksize) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=x.device)
    assert out.stride(3) == 1
    d = _conv_desc(x, cout, ksize, stride, out.stride(2), out_fp32, bias is not None, emit_stats, force_kc)
    if direct:
        check(lib().b200seg_conv2d_fwd_direct(ctypes.byref(d), ptr(x), ptr(w_ohwi), ptr(bias), ptr(out),
                                              stream_ptr()), "conv2d_fwd_direct")
        return out
    stats = None
    grid = ctypes.c_int32(0)
    if emit_stats:
        nelem = lib().b200seg_conv2d_stats_elems(ctypes.byref(d))
        stats = torch.empty(nelem, dtype=torch.float32, device=x.device)
    check(lib().b200seg_conv2d_fwd(ctypes.byref(d), ptr(x), ptr(w_ohwi), ptr(bias), ptr(out), ptr(stats),
                                   ctypes.byref(grid), stream_ptr()), "conv2d_fwd")
    if emit_stats:
        cout_pad = nelem // (148 * 2)
        return out, stats[: grid.value * 2 * cout_pad].view(grid.value, 2, cout_pad)
    return out
