"""network.deepv3 factory (network/deepv3.py:165-170) backed by the sm_100a kernels."""
from ._factory import build


def DeepV3PlusW38(num_classes, criterion):
    return build("deepv3.DeepV3PlusW38", num_classes, criterion)
