"""TEST INFRASTRUCTURE (oracle): the reference's training transform chain replayed with Pillow / torch for GIVEN random
decisions - the arithmetic lives in the reference's third-party dependencies (Pillow 12.2, torchvision 0.26 in this image;
unpinned by the reference), so the oracle calls them the way the reference does:

    transforms/joint_transforms.py:433-472  RandomSizeAndCrop: img.resize(BICUBIC), mask.resize(NEAREST)
    transforms/joint_transforms.py:143-182  RandomCrop (nopad=False): ImageOps.expand + crop
    transforms/joint_transforms.py:276-281  RandomHorizontallyFlip: transpose(FLIP_LEFT_RIGHT)
    transforms/transforms.py:192-300        adjust_brightness / contrast / saturation (ImageEnhance), adjust_hue (HSV)
    datasets/__init__.py:102-108            ToTensor, Normalize(mean, std), MaskToTensor
    datasets/base_loader.py:177-181         label id -> train id

Pinned by tests/test_augment_host.py against outputs of the UNMODIFIED reference classes run under the same seeds
(tests/golden/make_golden_augment.py -> tests/golden/reference_augment.pt). Only tests/ and __graft_entry__.smoke() may
import this module; the product (b200seg/augment.py + csrc/augment_kernels.cu) never does."""
import numpy as np
import torch
from PIL import Image, ImageEnhance, ImageOps


def synth_frame(h, w, seed):
    """Deterministic uint8 frame with smooth structure + noise (so resampling matters) and a blocky label-id mask."""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([127 + 100 * np.sin(xx / (7.0 + c) + yy / 11.0) * np.cos(yy / (5.0 + 2 * c)) for c in range(3)], -1)
    img = np.clip(base + g.normal(0, 25, (h, w, 3)), 0, 255).astype(np.uint8)
    mask = (g.integers(0, 34, ((h + 7) // 8, (w + 7) // 8)).repeat(8, 0).repeat(8, 1)[:h, :w]).astype(np.uint8)
    return img, mask


def adjust_hue(img, hue_factor):
    """transforms/transforms.py:252-300."""
    h, s, v = img.convert("HSV").split()
    np_h = np.array(h, dtype=np.uint8)
    # reference: np_h += np.uint8(hue_factor * 255). The NumPy the reference targets casts a negative float like C does
    # (truncate toward zero, wrap modulo 256); NumPy 2 raises OverflowError instead, so the cast is spelled out here
    np_h += np.uint8(int(hue_factor * 255) & 0xFF)
    h = Image.fromarray(np_h, "L")
    return Image.merge("HSV", (h, s, v)).convert("RGB")


def reference_chain(img_u8, mask_u8, p, crop_hw, ignore_label=255, mean=(0.485, 0.456, 0.406),
                    std=(0.229, 0.224, 0.225), id_to_trainid=None, full_size=False):
    """img_u8 [H,W,3] / mask_u8 [H,W] numpy uint8, p: b200seg.augment.AugParams (the random decisions).
    Returns (fp32 tensor [3,th,tw], int64 tensor [th,tw], uint8 crop [th,tw,3] before the colour jitter)."""
    mask_u8 = mask_u8.copy()
    if id_to_trainid:                                   # base_loader.py:177-181 (on a copy, from the original ids)
        src = mask_u8.copy()
        for k, v in id_to_trainid.items():
            mask_u8[src == k] = v
    img, mask = Image.fromarray(img_u8), Image.fromarray(mask_u8)
    w, h = img.size
    th, tw = (h, w) if full_size else crop_hw
    img = img.resize((p.rs_w, p.rs_h), Image.BICUBIC)
    mask = mask.resize((p.rs_w, p.rs_h), Image.NEAREST)
    if not (p.rs_w == tw and p.rs_h == th):
        if p.pad_x or p.pad_y:
            border = (p.pad_x, p.pad_y, p.pad_x, p.pad_y)
            img = ImageOps.expand(img, border=border, fill=(0, 0, 0))
            mask = ImageOps.expand(mask, border=border, fill=ignore_label)
        img = img.crop((p.x1, p.y1, p.x1 + tw, p.y1 + th))
        mask = mask.crop((p.x1, p.y1, p.x1 + tw, p.y1 + th))
    if p.flip:
        img, mask = img.transpose(Image.FLIP_LEFT_RIGHT), mask.transpose(Image.FLIP_LEFT_RIGHT)
    crop_u8 = np.array(img)
    for kind, factor in p.ops:
        if kind == 0:
            img = ImageEnhance.Brightness(img).enhance(factor)
        elif kind == 1:
            img = ImageEnhance.Contrast(img).enhance(factor)
        elif kind == 2:
            img = ImageEnhance.Color(img).enhance(factor)
        else:
            img = adjust_hue(img, factor)
    t = torch.from_numpy(np.array(img)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)     # ToTensor
    m, s = torch.as_tensor(mean, dtype=torch.float32), torch.as_tensor(std, dtype=torch.float32)
    t = t.sub_(m[:, None, None]).div_(s[:, None, None])                                              # Normalize
    lab = torch.from_numpy(np.array(mask, dtype=np.int32)).long()                                     # MaskToTensor
    return t, lab, crop_u8
