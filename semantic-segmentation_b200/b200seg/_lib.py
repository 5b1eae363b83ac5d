"""ctypes binding of libb200seg.so (the C ABI declared in include/b200seg.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc`` and loaded from ``b200seg/lib``.
There is deliberately no fallback: if the shared object is missing, importing a kernel-backed op raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200seg.so")

c_int32 = ctypes.c_int32
c_void_p = ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "n", "h", "w", "cin", "cout", "ksize", "stride", "pad", "x_ld", "y_ld",
        "out_fp32", "has_bias", "emit_stats", "reserved")]


class ProbeOperand(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "rows", "cols", "box_cols", "box_rows", "nboxes", "c0", "r0", "dcol", "drow", "smem_stride",
        "swizzle_bytes")]


class ProbeDesc(ctypes.Structure):
    _fields_ = [("a", ProbeOperand), ("b", ProbeOperand)] + [(n, c_int32) for n in (
        "M", "N", "ksteps",
        "a_off", "a_lbo", "a_sbo", "a_layout", "a_base", "a_major", "a_kstep",
        "b_off", "b_lbo", "b_sbo", "b_layout", "b_base", "b_major", "b_kstep")]


_lib = None


class B200SegError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200SegError(
                "libb200seg.so not found at %s - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the hot path)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.b200seg_abi_version.restype = ctypes.c_int
    L.b200seg_build_info.restype = ctypes.c_char_p
    L.b200seg_conv2d_stats_elems.restype = ctypes.c_size_t
    L.b200seg_conv2d_stats_elems.argtypes = [ctypes.POINTER(ConvDesc)]
    L.b200seg_conv2d_fwd.restype = ctypes.c_int
    L.b200seg_conv2d_fwd.argtypes = [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     ctypes.POINTER(c_int32), c_void_p]
    L.b200seg_conv2d_fwd_direct.restype = ctypes.c_int
    L.b200seg_conv2d_fwd_direct.argtypes = [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p]
    L.b200seg_pack_weight.restype = ctypes.c_int
    L.b200seg_pack_weight.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    L.b200seg_umma_probe.restype = ctypes.c_int
    L.b200seg_umma_probe.argtypes = [ctypes.POINTER(ProbeDesc), c_void_p, c_void_p, c_void_p, c_void_p]


def check(rc, what):
    if rc != 0:
        raise B200SegError("%s failed with status %d" % (what, rc))


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
