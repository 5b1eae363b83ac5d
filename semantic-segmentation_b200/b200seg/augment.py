"""Device-side training input pipeline: the reference's per-sample transform chain

    RandomSizeAndCrop (transforms/joint_transforms.py:433-472: PIL BICUBIC resize of the image, NEAREST of the mask,
    RandomCrop with ImageOps.expand padding :143-182) -> RandomHorizontallyFlip (:276-281) -> ColorJitter
    (transforms/transforms.py:297-362: ImageEnhance brightness / contrast / saturation + HSV hue shift, random order)
    -> ToTensor -> Normalize (datasets/__init__.py:102-104) and MaskToTensor, after the label-id remap of
    datasets/base_loader.py:177-181

as three sm_100a kernels (csrc/augment_kernels.cu) on the DECODED uint8 frame, so a 1024x2048 crop costs microseconds on
the GPU instead of a PIL worker process per GPU. Results are bit-exact with PIL 12 on the uint8 stages and with
torchvision's ToTensor / Normalize arithmetic on the fp32 stage (tests/test_gpu_augment.py): the host side below draws
the random parameters with the reference's own generator calls in the reference's order (`random` for the geometry,
`numpy.random` for the colour jitter), builds PIL's fixed-point resampling tables exactly as Pillow's
precompute_coeffs / normalize_coeffs_8bpc do (double arithmetic, 22 fractional bits) and hands them to the kernels.
There is no CPU image path here: the tables are a few KB of integers, the pixels never touch the host."""
import math
import random

import numpy as np
import torch

from . import raw

PRECISION_BITS = 32 - 8 - 2      # Pillow's 8 bits-per-channel resampler
OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_HUE = 0, 1, 2, 3


# ----------------------------------------------------------------------------------------------- Pillow's tables
def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_tables(in_size, out_size, lo=0, hi=None):
    """Pillow's ImagingResample coefficient tables (BICUBIC, 8 bpc) of output positions [lo, hi): int32 kk[hi-lo, ksize]
    (22-bit fixed point, rows zero padded) and int32 bounds[hi-lo, 2] = (first source index, tap count)."""
    hi = out_size if hi is None else hi
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((hi - lo, ksize), dtype=np.int32)
    bounds = np.zeros((hi - lo, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(lo, hi):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            kk[xx - lo, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx - lo] = (xmin, xmax)
    return kk, bounds


def nearest_table(in_size, out_size, lo=0, hi=None):
    """Source index of every NEAREST output position in [lo, hi): Pillow's affine scaler accumulates xo += in/out in
    double starting at in/out * 0.5 (the accumulated value, not (x + 0.5) * scale, decides ties)."""
    hi = out_size if hi is None else hi
    a0 = float(in_size) / out_size
    xo = a0 * 0.5
    tab = np.zeros(hi - lo, dtype=np.int32)
    for x in range(hi):
        if x >= lo:
            tab[x - lo] = min(max(int(math.floor(xo)), 0), in_size - 1)
        xo += a0
    return tab


# ----------------------------------------------------------------------------------------------- parameters
class AugParams:
    """One sample's random decisions, drawn like the reference draws them."""
    __slots__ = ("scale", "rs_w", "rs_h", "pad_x", "pad_y", "x1", "y1", "flip", "ops")

    def __repr__(self):
        return "AugParams(%s)" % ", ".join("%s=%r" % (k, getattr(self, k)) for k in self.__slots__)


def draw_params(w, h, crop_hw, scale_min=0.5, scale_max=2.0, color_aug=0.25, centroid=None, pre_size=None,
                full_size=False, rng=random, nprng=np.random):
    """Consumes `random` / `numpy.random` exactly like RandomSizeAndCrop.__call__ -> RandomCrop.__call__ (nopad=False,
    TRANSLATE_AUG_FIX off) -> RandomHorizontallyFlip -> ColorJitter.get_params do for one sample."""
    p = AugParams()
    scale_amt = rng.uniform(scale_min, scale_max)                      # joint_transforms.py:446
    if pre_size is not None:                                           # :448-456
        scale_amt *= pre_size / (w if w > h else h)
    th, tw = (h, w) if full_size else crop_hw                          # :458-459 (crop.size = (h, w) of the input)
    p.scale = scale_amt
    p.rs_w, p.rs_h = int(w * scale_amt), int(h * scale_amt)            # :461
    if centroid is not None:
        centroid = [int(c * scale_amt) for c in centroid]              # :463-464
    rw, rh = p.rs_w, p.rs_h
    p.pad_x = p.pad_y = 0
    p.x1 = p.y1 = 0
    if not (rw == tw and rh == th):                                    # RandomCrop.__call__ :143-182
        if th > rh:
            p.pad_y = (th - rh) // 2 + 1
        if tw > rw:
            p.pad_x = (tw - rw) // 2 + 1
        rw, rh = rw + 2 * p.pad_x, rh + 2 * p.pad_y
        if centroid is not None:                                       # crop_in_image :104-126
            c_x, c_y = centroid
            x1 = rng.randint(c_x - tw, c_x)
            p.x1 = min(rw - tw, max(0, x1))
            y1 = rng.randint(c_y - th, c_y)
            p.y1 = min(rh - th, max(0, y1))
        else:
            p.x1 = 0 if rw == tw else rng.randint(0, rw - tw)
            p.y1 = 0 if rh == th else rng.randint(0, rh - th)
    p.flip = rng.random() < 0.5                                        # :278
    ops = []
    if color_aug and color_aug > 0:                                    # transforms.py:326-348
        ops.append((OP_BRIGHTNESS, float(nprng.uniform(max(0, 1 - color_aug), 1 + color_aug))))
        ops.append((OP_CONTRAST, float(nprng.uniform(max(0, 1 - color_aug), 1 + color_aug))))
        ops.append((OP_SATURATION, float(nprng.uniform(max(0, 1 - color_aug), 1 + color_aug))))
        ops.append((OP_HUE, float(nprng.uniform(-color_aug, color_aug))))
        nprng.shuffle(ops)
    p.ops = [(int(k), float(f)) for k, f in ops]
    return p


# ----------------------------------------------------------------------------------------------- the transform
class DeviceTrainTransform:
    """image uint8 [H, W, 3] + label-id mask uint8 [H, W] (CUDA tensors) -> (fp32 [3, th, tw] normalised image, int64
    [th, tw] train-id labels), the outputs of the reference's train_joint_transform_list + train_input_transform +
    target_transform (datasets/__init__.py:72-108) for the same random decisions."""

    def __init__(self, crop_hw, scale_min=0.5, scale_max=2.0, color_aug=0.25, ignore_label=255,
                 mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), id_to_trainid=None, pre_size=None,
                 full_size=False):
        self.crop_hw = tuple(crop_hw)
        self.scale_min, self.scale_max, self.color_aug = scale_min, scale_max, color_aug
        self.ignore_label = int(ignore_label)
        self.mean, self.std = tuple(float(m) for m in mean), tuple(float(s) for s in std)
        self.pre_size, self.full_size = pre_size, full_size
        lut = np.arange(256, dtype=np.uint8)
        if id_to_trainid:                      # datasets/base_loader.py:177-181 (applied to the decoded label ids)
            for k, v in id_to_trainid.items():
                if 0 <= int(k) < 256:
                    lut[int(k)] = int(v) & 0xFF
        self._lut_host = torch.from_numpy(lut)
        self._lut = None

    def draw(self, w, h, centroid=None, rng=random, nprng=np.random):
        return draw_params(w, h, self.crop_hw, self.scale_min, self.scale_max, self.color_aug, centroid, self.pre_size,
                           self.full_size, rng, nprng)

    def tables(self, p, w, h):
        """Resampling tables of the crop window only (host integers, a few KB)."""
        th, tw = (h, w) if self.full_size else self.crop_hw
        # output column c (before the flip) shows padded-resized column x1 + c, i.e. resized column x1 + c - pad_x
        lo_x, hi_x = max(0, p.x1 - p.pad_x), min(p.rs_w, p.x1 - p.pad_x + tw)
        lo_y, hi_y = max(0, p.y1 - p.pad_y), min(p.rs_h, p.y1 - p.pad_y + th)
        kh, bh = bicubic_tables(w, p.rs_w, lo_x, max(lo_x, hi_x))
        kv, bv = bicubic_tables(h, p.rs_h, lo_y, max(lo_y, hi_y))
        nx = nearest_table(w, p.rs_w, lo_x, max(lo_x, hi_x))
        ny = nearest_table(h, p.rs_h, lo_y, max(lo_y, hi_y))
        return dict(kh=kh, bh=bh, kv=kv, bv=bv, nx=nx, ny=ny, lo_x=lo_x, lo_y=lo_y, n_x=max(0, hi_x - lo_x),
                    n_y=max(0, hi_y - lo_y), th=th, tw=tw)

    def __call__(self, image_u8, mask_u8, params=None, centroid=None, out_image=None, out_label=None):
        assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.dim() == 3 and image_u8.shape[2] == 3
        assert mask_u8.is_cuda and mask_u8.dtype == torch.uint8 and tuple(mask_u8.shape) == tuple(image_u8.shape[:2])
        h, w = image_u8.shape[:2]
        p = params if params is not None else self.draw(w, h, centroid)
        t = self.tables(p, w, h)
        dev = image_u8.device
        if self._lut is None or self._lut.device != dev:
            self._lut = self._lut_host.to(dev)
        img, lab, self.last_rgb_u8 = raw.augment(image_u8.contiguous(), mask_u8.contiguous(), p, t, self._lut,
                                                 self.ignore_label, self.mean, self.std, out_image, out_label)
        return img, lab
