"""Overlay of the reference's network/mscale.py for the HRNet trunk (network/mscale.py:450-475): the arch string
``mscale.HRNet`` = MscaleBasic (seg head + scale-attention head straight on the HRNet features, hierarchical two-scale
training forward / n-scale evaluation of MscaleBase, :114-231). The DeepLab-style trunks of that file (DeeperW38 ...)
are outside the hot path."""
from ._factory import build


def HRNet(num_classes, criterion, s2s4=None):
    """network/mscale.py:473-475."""
    return build("mscale.HRNet", num_classes, criterion)
