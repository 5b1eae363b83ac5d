"""world_size-2 gloo test of the data-parallel host logic (collective C1): every rank computes gradients on its own
shard into one flat fp32 buffer, the buffer is averaged across ranks when the gradients are published, and parameters
stay in sync after the optimizer step. Runs on CPU tensors (the flat-buffer logic is device agnostic)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200seg.module import allreduce_mean_
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(37))
    flat = torch.zeros(64)
    x = torch.arange(37, dtype=torch.float32) * (rank + 1)       # a different shard per rank
    flat[:37] = 2 * (w.detach() - x)                              # d/dw sum (w - x)^2
    allreduce_mean_(flat)
    expect = 2 * (w.detach() - torch.arange(37, dtype=torch.float32) * (1 + 2) / 2)
    ok = torch.allclose(flat[:37], expect, atol=1e-6)
    w.grad = flat[:37]
    torch.optim.SGD([w], lr=0.1).step()
    gathered = [torch.zeros(37) for _ in range(world)]
    dist.all_gather(gathered, w.detach())
    ok = ok and torch.equal(gathered[0], gathered[1])
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_single_process_is_a_noop():
    from b200seg.module import allreduce_mean_
    t = torch.ones(8)
    allreduce_mean_(t)
    assert torch.equal(t, torch.ones(8))
