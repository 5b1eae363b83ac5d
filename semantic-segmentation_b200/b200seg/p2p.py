"""SyncBN mailboxes: peer-mapped device memory shared by the ranks of one node (NVLink / NVSwitch, CUDA IPC).

Replaces apex.parallel.SyncBatchNorm's 1264 per-layer all_gather / all_reduce calls per two-scale step (SURVEY.md §2b
collective C2, config.py:216-225) with a one-shot exchange fused into the BatchNorm finalisers (csrc/bn_kernels.cu
sync_exchange). This module only owns the plumbing: allocation, handle exchange over torch.distributed, offset tables.
"""
import ctypes

import torch

from ._lib import BnSync, check, lib


def exchange_layout(bn_channels, world, n_passes):
    """Offsets of every (pass, layer, direction) exchange: -> (mail_off, flag_off, doubles per parity half, flags).
    Exchange id order = pass-major, then layer, then direction (0 forward statistics, 1 backward sums); a mail region
    is [rank][2][C] doubles, a flag region [rank][ceil(C/32)] uint32. Pure function (unit-tested on the CPU)."""
    mail_off, flag_off = {}, {}
    moff = foff = 0
    for p in range(n_passes):
        for name, c in bn_channels.items():
            for d in (0, 1):
                mail_off[(p, name, d)] = moff
                flag_off[(p, name, d)] = foff
                moff += world * 2 * c
                foff += world * ((c + 31) // 32)
    return mail_off, flag_off, moff, foff


class SyncBNContext:
    def __init__(self, bn_channels, n_passes=2, group=None):
        """bn_channels: ordered {bn layer name: channels}; every rank must pass the same table."""
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.names = list(bn_channels)
        self.n_passes = n_passes
        dev = torch.device("cuda", torch.cuda.current_device())
        # the one-warp waiter kernels (1 KB of reserved shared memory each, up to 2 passes x 4 branches in flight, several
        # blocks per exchange) must fit next to any convolution CTA (csrc/bn_kernels.cu sync_wait_fold)
        check(lib().b200seg_set_smem_reserve(12288), "set_smem_reserve", 0)
        self.mail_off, self.flag_off, moff, foff = exchange_layout(bn_channels, self.world, n_passes)
        self.parity_stride = moff
        L = lib()
        self._own = []
        self._opened = []
        mail_ptrs = self._shared_alloc(L, 2 * moff * 8)
        flag_ptrs = self._shared_alloc(L, foff * 4)
        self.mail_table = torch.tensor(mail_ptrs, dtype=torch.int64, device=dev)
        self.flag_table = torch.tensor(flag_ptrs, dtype=torch.int64, device=dev)
        self.step = torch.zeros((1,), dtype=torch.int32, device=dev)     # uint32 on the device; +1 per training step
        # pinned host record the kernels write (UVA: same pointer on the device): readable after a trapped / hung run
        self.beacon = torch.zeros((8,), dtype=torch.int32).pin_memory()
        dist.barrier(group)

    def _shared_alloc(self, L, nbytes):
        import torch.distributed as dist
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        check(L.b200seg_p2p_alloc(nbytes, ctypes.byref(ptr), handle), "p2p_alloc", 0)
        self._own.append(ptr.value)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        ptrs = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                ptrs.append(ptr.value)
                continue
            peer = ctypes.c_void_p()
            buf = (ctypes.c_uint8 * 64).from_buffer_copy(hb)
            check(L.b200seg_p2p_open(buf, ctypes.byref(peer)), "p2p_open", 0)
            self._opened.append(peer.value)
            ptrs.append(peer.value)
        return ptrs

    def advance(self):
        """Once per training step, on the step's stream (captured into the CUDA graph)."""
        self.step.add_(1)

    def args(self, pass_id, name, direction):
        s = BnSync()
        s.mail_peers = self.mail_table.data_ptr()
        s.flag_peers = self.flag_table.data_ptr()
        s.step = self.step.data_ptr()
        s.mail_offset = self.mail_off[(pass_id, name, direction)]
        s.parity_stride = self.parity_stride
        s.flag_offset = self.flag_off[(pass_id, name, direction)]
        s.world, s.rank = self.world, self.rank
        s.beacon = self.beacon.data_ptr()
        return s

    def describe_beacon(self):
        """Human-readable post-mortem: the exchange this rank entered / completed last (see b200seg_bn_sync.beacon)."""
        inv = {v: k for k, v in self.flag_off.items()}
        b = self.beacon.tolist()
        return "rank %d: entered %s at step %d, completed %s at step %d" % (self.rank, inv.get(b[0]), b[1], inv.get(b[4]), b[5])

    def close(self):
        L = lib()
        torch.cuda.synchronize()
        for p in self._opened:
            L.b200seg_p2p_close(ctypes.c_void_p(p))
        for p in self._own:
            L.b200seg_p2p_free(ctypes.c_void_p(p))
        self._opened, self._own = [], []
