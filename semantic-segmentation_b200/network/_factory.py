"""Shared constructor: reads the reference cfg keys the original modules read (SURVEY.md §5 'Config / flags')."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from b200seg import arch as A            # noqa: E402
from b200seg.module import B200SegModule  # noqa: E402


def build(arch, num_classes, criterion):
    kw = {}
    try:
        from config import cfg
        if not arch.startswith("deepv3."):
            kw["hcfg"] = A.hrnet_cfg_from_reference_cfg(cfg)
        ocfg = dict(A.OCR_DEFAULT)
        ocfg["mid_channels"] = cfg.MODEL.OCR.MID_CHANNELS
        ocfg["key_channels"] = cfg.MODEL.OCR.KEY_CHANNELS
        ocfg["segattn_bot_ch"] = cfg.MODEL.SEGATTN_BOT_CH
        kw["ocfg"] = ocfg
        kw["lo_scale"] = cfg.MODEL.MSCALE_LO_SCALE
        kw["ocr_alpha"] = cfg.LOSS.OCR_ALPHA
        kw["supervised_mscale_wt"] = cfg.LOSS.SUPERVISED_MSCALE_WT
        kw["ignore_index"] = cfg.DATASET.IGNORE_LABEL
        kw["n_scales"] = cfg.MODEL.N_SCALES
        if cfg.MODEL.ALIGN_CORNERS:
            raise NotImplementedError("align_corners=True is not covered by the B200 resampling kernels")
        if cfg.MODEL.OCR_ASPP or cfg.MODEL.MSCALE_OLDARCH or not cfg.MODEL.MSCALE_INNER_3x3 or cfg.MODEL.MSCALE_DROPOUT:
            raise NotImplementedError("only the default OCR / attention-head configuration is on the hot path")
        if cfg.LOSS.OCR_AUX_RMI:
            raise NotImplementedError("OCR_AUX_RMI is not covered")
    except ImportError:
        pass
    return B200SegModule(arch, num_classes=num_classes, criterion=criterion, **kw)
