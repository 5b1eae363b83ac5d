import torch

SyncBatchNorm = torch.nn.SyncBatchNorm


class DistributedDataParallel(torch.nn.parallel.DistributedDataParallel):
    """apex.parallel.DistributedDataParallel(net) call shape (network/__init__.py:38-39) on top of torch DDP."""

    def __init__(self, module, **kwargs):
        kwargs.pop("delay_allreduce", None)
        dev = next(module.parameters()).device
        if dev.type == "cuda":
            super().__init__(module, device_ids=[dev.index], output_device=dev.index)
        else:
            super().__init__(module)
