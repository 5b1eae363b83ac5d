#!/usr/bin/env python
"""Timing (GPU) of the device input pipeline on a 1024x2048 Cityscapes-size frame -> 1024x2048 crop: device kernels (tables
built on the host per sample, included) next to the PIL / torchvision replay of the reference chain on one host core."""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from b200seg import augment as AUG  # noqa: E402
from oracle import augment_oracle as AO  # noqa: E402


def main():
    img_u8, mask_u8 = AO.synth_frame(1024, 2048, 0)
    t = AUG.DeviceTrainTransform((1024, 2048))
    img_d, mask_d = torch.from_numpy(img_u8).cuda(), torch.from_numpy(mask_u8).cuda()
    random.seed(0)
    np.random.seed(0)
    params = [t.draw(2048, 1024) for _ in range(12)]
    for p in params[:2]:
        t(img_d, mask_d, params=p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for p in params[2:]:
        t(img_d, mask_d, params=p)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 10
    dev = e0.elapsed_time(e1) / 10
    t1 = time.perf_counter()
    for p in params[2:5]:
        AO.reference_chain(img_u8, mask_u8, p, (1024, 2048))
    host = (time.perf_counter() - t1) / 3
    print(json.dumps(dict(device_ms_per_sample=dev, wall_ms_per_sample_incl_host_tables=wall * 1e3,
                          pil_chain_ms_per_sample_one_core=host * 1e3, scales=[round(p.scale, 3) for p in params[2:]])))


if __name__ == "__main__":
    main()
