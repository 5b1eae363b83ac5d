"""Golden vectors for SURVEY.md §8(f) row f2 — DeepLabV3+ / WideResNet-38 (network/deepv3.py, BASELINE config 4) — from the
UNMODIFIED reference on CPU. Run in the build container only:  python tests/golden/make_golden_deepv3.py
Writes tests/golden/reference_deepv3.pt (loss, sampled gradients, sampled eval prediction, state_dict key order)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "semantic-segmentation_b200", "shims"))
sys.path.insert(1, "/root/reference")

import numpy as np
import torch

np.int = int
from oracle import seg_oracle as O  # noqa: E402
from config import cfg  # noqa: E402

torch.set_num_threads(8)
cfg.MODEL.BNFUNC = torch.nn.BatchNorm2d
cfg.MODEL.HRNET_CHECKPOINT = ""
cfg.DATASET.NUM_CLASSES = 19
cfg.DATASET.IGNORE_LABEL = 255
cfg.OPTIONS.TORCH_VERSION = 2.1

import network.wider_resnet as wr  # noqa: E402

_orig_init = wr.wrn38.__init__


def _init_no_ckpt(self, pretrained=True):          # SURVEY §8c shim (8): the pretrained file does not exist offline
    _orig_init(self, pretrained=False)


wr.wrn38.__init__ = _init_no_ckpt
import network.deepv3 as ref_deepv3  # noqa: E402
from loss.utils import CrossEntropyLoss2d  # noqa: E402


def main():
    arch = "deepv3.DeepV3PlusW38"
    net = ref_deepv3.DeepV3PlusW38(num_classes=19, criterion=CrossEntropyLoss2d(ignore_index=255))
    sd = O.synth_state_dict(arch, O.WRN38, seed=4)
    ref_keys = list(net.state_dict().keys())
    assert ref_keys == list(sd.keys()), [k for k in ref_keys if k not in sd][:5] + [k for k in sd if k not in ref_keys][:5]
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    net.load_state_dict(sd)
    for m in net.modules():
        if isinstance(m, (torch.nn.Dropout2d, torch.nn.Dropout)):
            m.p = 0.0
    images, gts = O.synth_batch(2, 64, 128, seed=6)
    net.train()
    loss = net({"images": images, "gts": gts})
    loss.backward()
    named = dict(net.named_parameters())
    gnames = ("backbone.mod1.conv1.weight", "backbone.mod4.block2.convs.conv2.weight", "backbone.mod5.block1.proj_conv.weight",
              "backbone.mod6.block1.convs.conv2.weight", "backbone.mod7.block1.convs.bn3.0.weight",
              "aspp.features.2.0.weight", "aspp.img_conv.0.weight", "bot_fine.weight", "final.6.weight")
    out = dict(loss=loss.detach().clone(), grads={n: O.sample_like(named[n].grad) for n in gnames},
               running_var_mod5=net.state_dict()["backbone.mod5.block3.convs.bn2.0.running_var"].clone(),
               keys=ref_keys, param_order=[n for n, _ in net.named_parameters()],
               nparams=sum(p.numel() for p in net.parameters()))
    net.eval()
    with torch.no_grad():
        o = net({"images": images})
    out["eval_pred"] = O.sample_like(o["pred"])
    print("deepv3 loss", float(loss), "params", out["nparams"], "keys", len(ref_keys))
    torch.save(out, os.path.join(HERE, "reference_deepv3.pt"))
    print("wrote reference_deepv3.pt", os.path.getsize(os.path.join(HERE, "reference_deepv3.pt")))


if __name__ == "__main__":
    main()
