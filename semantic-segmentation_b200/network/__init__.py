"""Drop-in replacement of the reference's ``network`` package for the HRNet-OCR-MScale hot path.

Put ``semantic-segmentation_b200/`` in front of the reference checkout on ``sys.path`` and the
reference's own ``train.py`` keeps calling exactly what it called before:

    net = network.get_net(args, criterion)                      # network/__init__.py:12-23
    net = network.wrap_network_in_dataparallel(net, args.apex)  # network/__init__.py:33-42

``get_model`` keeps the importlib contract (``'network.<module>.<factory>'``, network/__init__.py:45-54); the
``network.ocrnet`` / ``network.basic`` modules of this package return ``B200SegModule`` instances whose parameters
carry the reference's names, so ``loss/optimizer.py``'s restore functions and ``logx.save_model`` checkpoints
interoperate. Architectures outside the hot path (DeepLabV3+, SEResNeXt, ...) are deliberately not provided here.
"""
import importlib

import torch


def _cfg():
    try:
        from config import cfg   # the reference's global config when its checkout is on sys.path
        return cfg
    except Exception:
        return None


def _logx():
    try:
        from runx.logx import logx
        return logx
    except Exception:
        class _L:
            @staticmethod
            def msg(s):
                print(s)
        return _L


def get_net(args, criterion):
    """Same contract as network.get_net (network/__init__.py:12-23): build, log the parameter count, move to CUDA."""
    cfg = _cfg()
    num_classes = cfg.DATASET.NUM_CLASSES if cfg is not None else getattr(args, "num_classes", 19)
    net = get_model(network="network." + args.arch, num_classes=num_classes, criterion=criterion)
    num_params = sum(p.nelement() for p in net.parameters())
    _logx().msg("Model params = {:2.1f}M".format(num_params / 1000000))
    return net.cuda()


def is_gscnn_arch(args):
    return "gscnn" in args.arch


def wrap_network_in_dataparallel(net, use_apex_data_parallel=False):
    """network/__init__.py:33-42. One process per GPU: instead of per-parameter autograd hooks the module all-reduces
    its single flat fp32 gradient buffer over NCCL when ``loss.backward()`` publishes the gradients, so the wrapper is
    only a thin container that keeps the ``module.`` state_dict prefix train.py / logx.save_model expect."""
    if use_apex_data_parallel:
        return FlatGradDataParallel(net)
    if torch.cuda.device_count() <= 1:
        # torch.nn.DataParallel on a single device calls the module directly (network/__init__.py:41): keep the
        # `module.` state_dict prefix the checkpoint code expects, no collective
        return FlatGradDataParallel(net, allreduce=False)
    raise NotImplementedError("single-process nn.DataParallel replication over several GPUs is outside the B200 hot "
                              "path; launch one process per GPU (torch.distributed.launch / torchrun) with --apex")


class FlatGradDataParallel(torch.nn.Module):
    """Stands in for apex.parallel.DistributedDataParallel(net) (collective C1 of SURVEY.md §2b)."""

    def __init__(self, module, allreduce=True):
        super().__init__()
        self.module = module
        module._ddp_allreduce = bool(allreduce)
        if allreduce and torch.distributed.is_available() and torch.distributed.is_initialized():
            # replicate rank 0's initial weights like DDP does at construction
            for t in list(module.parameters()) + list(module.buffers()):
                torch.distributed.broadcast(t.data, src=0)

    def forward(self, *a, **k):
        return self.module(*a, **k)


def get_model(network, num_classes, criterion):
    """network/__init__.py:45-54."""
    module = network[:network.rfind(".")]
    model = network[network.rfind(".") + 1:]
    mod = importlib.import_module(module)
    net_func = getattr(mod, model)
    return net_func(num_classes=num_classes, criterion=criterion)
