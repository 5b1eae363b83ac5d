"""SURVEY §8 row f4 on the GPU: the device input pipeline (b200seg.augment.DeviceTrainTransform, csrc/augment_kernels.cu)
against the PIL / torchvision replay of the reference's transform chain (oracle/augment_oracle.py, itself pinned to the
unmodified reference by tests/test_augment_host.py): the uint8 crop after resize / pad / crop / flip is bit-exact, the labels
are exact, the normalised fp32 image is bit-exact (IEEE division on both sides)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))

pytestmark = pytest.mark.gpu

CITYSCAPES_IDS = {-1: 255, 0: 255, 1: 255, 2: 255, 3: 255, 4: 255, 5: 255, 6: 255, 7: 0, 8: 1, 9: 255, 10: 255, 11: 2,
                  12: 3, 13: 4, 14: 255, 15: 255, 16: 255, 17: 5, 18: 255, 19: 6, 20: 7, 21: 8, 22: 9, 23: 10, 24: 11,
                  25: 12, 26: 13, 27: 14, 28: 15, 29: 255, 30: 255, 31: 16, 32: 17, 33: 18}


def _run(h, w, crop, smin, smax, caug, seed, lut=None, params=None):
    from b200seg import augment as AUG
    from oracle import augment_oracle as AO
    img_u8, mask_u8 = AO.synth_frame(h, w, seed)
    t = AUG.DeviceTrainTransform(crop, smin, smax, caug, id_to_trainid=lut)
    random.seed(seed)
    np.random.seed(seed)
    p = params if params is not None else t.draw(w, h)
    lut_pos = {k: v for k, v in (lut or {}).items() if k >= 0}
    ref_img, ref_lab, ref_u8 = AO.reference_chain(img_u8, mask_u8, p, crop, 255, t.mean, t.std, id_to_trainid=lut_pos)
    img, lab = t(torch.from_numpy(img_u8).cuda(), torch.from_numpy(mask_u8).cuda(), params=p)
    torch.cuda.synchronize()
    got_u8 = t.last_rgb_u8.cpu().numpy()
    assert np.array_equal(got_u8, ref_u8), (p, int((got_u8 != ref_u8).sum()))
    assert torch.equal(lab.cpu(), ref_lab), p
    d = (img.cpu() - ref_img).abs().max().item()
    assert torch.equal(img.cpu(), ref_img), (p, d)
    return p


@pytest.mark.parametrize("case", [(96, 160, (64, 96), 0.5, 2.0, 0.25), (80, 144, (96, 160), 0.5, 1.0, 0.25),
                                  (80, 144, (96, 160), 0.5, 2.0, 0.4), (128, 256, (64, 128), 1.0, 1.0, 0.0),
                                  (64, 128, (64, 128), 1.0, 1.0, 0.25), (100, 180, (72, 120), 0.6, 1.7, 0.25)])
def test_device_chain_matches_pil_replay(case):
    h, w, crop, smin, smax, caug = case
    pads = flips = 0
    for seed in range(1, 9):
        p = _run(h, w, crop, smin, smax, caug, seed)
        pads += bool(p.pad_x or p.pad_y)
        flips += bool(p.flip)
    assert flips >= 1


def test_every_jitter_op_alone_and_extreme_factors():
    from b200seg import augment as AUG
    for kind, factors in ((0, (0.6, 1.0, 1.4)), (1, (0.6, 1.4)), (2, (0.0, 0.6, 1.4)), (3, (-0.5, -0.1, 0.0, 0.37, 0.5))):
        for f in factors:
            p = AUG.AugParams()
            p.scale, p.rs_w, p.rs_h, p.pad_x, p.pad_y, p.x1, p.y1, p.flip = 1.25, 200, 120, 0, 0, 17, 9, kind == 1
            p.ops = [(kind, f)]
            _run(96, 160, (64, 96), 0.5, 2.0, 0.25, 20 + kind, params=p)
    # all four in a fixed order, contrast last and first (the grey level depends on the ops before it)
    for order in ((0, 2, 3, 1), (1, 0, 3, 2)):
        p = AUG.AugParams()
        p.scale, p.rs_w, p.rs_h, p.pad_x, p.pad_y, p.x1, p.y1, p.flip = 0.75, 120, 72, 0, 0, 11, 3, False
        p.ops = [(k, {0: 1.21, 1: 0.83, 2: 1.17, 3: -0.21}[k]) for k in order]
        _run(96, 160, (64, 96), 0.5, 2.0, 0.25, 31, params=p)


def test_full_size_crop_with_label_lookup():
    """BASELINE shape: a 1024x2048 crop of a 1024x2048 Cityscapes-size frame, label ids -> train ids on the device."""
    for seed in (3, 4):
        _run(1024, 2048, (1024, 2048), 0.5, 2.0, 0.25, seed, lut=CITYSCAPES_IDS)


def test_batch_wrapper_draws_in_loader_order():
    """DeviceAugmentedBatches: sample i of a batch gets the i-th draw; the batch dict is what train.py feeds the network."""
    from b200seg import augment as AUG
    from oracle import augment_oracle as AO
    frames = [AO.synth_frame(96, 160, s) for s in (41, 42, 43)]
    imgs = torch.from_numpy(np.stack([f[0] for f in frames])).pin_memory()
    masks = torch.from_numpy(np.stack([f[1] for f in frames])).pin_memory()
    t = AUG.DeviceTrainTransform((64, 96))
    random.seed(7)
    np.random.seed(7)
    batch = next(iter(AUG.DeviceAugmentedBatches([(imgs, masks)], t)))
    torch.cuda.synchronize()
    assert tuple(batch["images"].shape) == (3, 3, 64, 96) and batch["gts"].dtype == torch.int64
    random.seed(7)
    np.random.seed(7)
    for i, (img_u8, mask_u8) in enumerate(frames):
        p = t.draw(160, 96)
        ref_img, ref_lab, _ = AO.reference_chain(img_u8, mask_u8, p, (64, 96), 255, t.mean, t.std)
        assert torch.equal(batch["images"][i].cpu(), ref_img) and torch.equal(batch["gts"][i].cpu(), ref_lab), i
