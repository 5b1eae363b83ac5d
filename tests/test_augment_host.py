"""SURVEY §8 row f4 (training input pipeline), host half, no GPU: (1) b200seg.augment.draw_params consumes `random` /
`numpy.random` exactly like the reference's transform classes, and the PIL replay of those decisions
(oracle/augment_oracle.py) reproduces the UNMODIFIED reference's outputs bit for bit (tests/golden/reference_augment.pt);
(2) the resampling tables the product hands to its kernels are Pillow's: applied with integer numpy arithmetic they give
Image.resize(BICUBIC / NEAREST) exactly, also for a sub-window."""
import hashlib
import os
import random
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))

from b200seg import augment as AUG  # noqa: E402
from oracle import augment_oracle as AO  # noqa: E402


def test_draws_and_oracle_chain_reproduce_the_reference():
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "reference_augment.pt"), weights_only=False)
    seen_pad = seen_flip = seen_ops = 0
    for i, (h, w, crop, smin, smax, caug, seed) in enumerate(gold["cases"]):
        img_u8, mask_u8 = AO.synth_frame(h, w, seed)
        random.seed(seed)
        np.random.seed(seed)
        p = AUG.draw_params(w, h, tuple(crop), smin, smax, caug)
        image, label, _ = AO.reference_chain(img_u8, mask_u8, p, tuple(crop), 255, gold["mean"], gold["std"])
        sha = (hashlib.sha256(image.numpy().tobytes()).hexdigest(), hashlib.sha256(label.numpy().tobytes()).hexdigest())
        assert sha == tuple(gold["sha256"][i]), (i, p)
        if i < len(gold["image"]):
            assert torch.equal(image, gold["image"][i]) and torch.equal(label, gold["label"][i])
        seen_pad += bool(p.pad_x or p.pad_y)
        seen_flip += bool(p.flip)
        seen_ops += len(p.ops)
    assert seen_pad >= 1 and seen_flip >= 1 and seen_ops >= 4     # the cases exercise padding, flips and the jitter


def test_centroid_pre_size_and_full_size_draws_reproduce_the_reference():
    """The other arguments of RandomSizeAndCrop: class-uniform centroids (crop_in_image with a centroid,
    joint_transforms.py:104-116), --pre_size (:448-456), --full_crop_training (:458-459)."""
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "reference_augment.pt"), weights_only=False)
    for i, (h, w, crop, smin, smax, caug, seed, centroid, pre_size, full_size) in enumerate(gold["cases2"]):
        img_u8, mask_u8 = AO.synth_frame(h, w, seed)
        random.seed(seed)
        np.random.seed(seed)
        p = AUG.draw_params(w, h, tuple(crop), smin, smax, caug, centroid=centroid, pre_size=pre_size, full_size=full_size)
        image, label, _ = AO.reference_chain(img_u8, mask_u8, p, tuple(crop), 255, gold["mean"], gold["std"],
                                             full_size=full_size)
        sha = (hashlib.sha256(image.numpy().tobytes()).hexdigest(), hashlib.sha256(label.numpy().tobytes()).hexdigest())
        assert tuple(image.shape) == tuple(gold["sha256_2"][i][2]), (i, p)
        assert sha == tuple(gold["sha256_2"][i][:2]), (i, p)


def _apply_bicubic(a, kk, bounds, axis):
    """Pillow's pass along `axis` with the product's tables (integer arithmetic, test only)."""
    a = np.moveaxis(a.astype(np.int64), axis, 0)
    out = np.zeros((kk.shape[0],) + a.shape[1:], dtype=np.uint8)
    for i in range(kk.shape[0]):
        x0, n = bounds[i]
        acc = (1 << 21) + np.tensordot(kk[i, :n].astype(np.int64), a[x0:x0 + n], axes=(0, 0))
        out[i] = np.clip(acc >> 22, 0, 255)
    return np.moveaxis(out, 0, axis)


def test_resampling_tables_are_pillows():
    img_u8, mask_u8 = AO.synth_frame(61, 97, 11)
    im, mk = Image.fromarray(img_u8), Image.fromarray(mask_u8)
    for (ow, oh) in [(97, 61), (48, 30), (49, 31), (130, 77), (193, 121), (60, 100)]:
        ref = np.array(im.resize((ow, oh), Image.BICUBIC))
        kh, bh = AUG.bicubic_tables(97, ow)
        kv, bv = AUG.bicubic_tables(61, oh)
        mine = _apply_bicubic(_apply_bicubic(img_u8, kh, bh, 1), kv, bv, 0)
        assert np.array_equal(mine, ref), (ow, oh)
        # a sub-window of the tables gives the same pixels
        lo, hi = ow // 3, ow - 2
        kh2, bh2 = AUG.bicubic_tables(97, ow, lo, hi)
        assert np.array_equal(kh2, kh[lo:hi]) and np.array_equal(bh2, bh[lo:hi])
        refm = np.array(mk.resize((ow, oh), Image.NEAREST))
        nx, ny = AUG.nearest_table(97, ow), AUG.nearest_table(61, oh)
        assert np.array_equal(mask_u8[ny][:, nx], refm), (ow, oh)
        assert np.array_equal(AUG.nearest_table(97, ow, lo, hi), nx[lo:hi])


def test_window_geometry_covers_padding_and_offsets():
    t = AUG.DeviceTrainTransform((64, 96), color_aug=0.0)
    p = AUG.AugParams()
    p.scale, p.rs_w, p.rs_h, p.pad_x, p.pad_y, p.x1, p.y1, p.flip, p.ops = 0.5, 80, 48, 9, 9, 1, 2, False, []
    tb = t.tables(p, 160, 96)
    # padded frame 98 x 66; crop origin (1, 2): output column c shows resized column c - 8, row r shows r - 7
    assert (tb["lo_x"], tb["n_x"], tb["lo_y"], tb["n_y"]) == (0, 80, 0, 48)
    p.pad_x = p.pad_y = 0
    p.rs_w, p.rs_h, p.x1, p.y1 = 200, 120, 30, 20
    tb = t.tables(p, 160, 96)
    assert (tb["lo_x"], tb["n_x"], tb["lo_y"], tb["n_y"]) == (30, 96, 20, 64)
