"""Import-only stub: transforms/transforms.py:36-37,44 imports skimage unconditionally although only the optional
--gblur/--bblur/--jointwtborder paths use it."""
from . import filters, restoration, segmentation, morphology  # noqa: F401
