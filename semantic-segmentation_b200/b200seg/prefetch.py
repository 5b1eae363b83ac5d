"""Host -> device input staging for the training loop (the transfer half of the input pipeline;
reference: train.py:485-488 moves every batch with ``.cuda()`` on the compute stream right before ``net(inputs)``).

``DevicePrefetcher`` wraps any iterable of ``{'images': fp32 [N,3,H,W], 'gts': int64 [N,H,W]}`` batches held in (ideally
pinned) host memory and yields the same dicts on the device: the copy of batch i+1 runs on a dedicated copy stream into
the other half of a double buffer while step i executes, so the PCIe transfer (42 MB per 1024x2048 crop) leaves the
critical path. The buffers are reused: a yielded batch is valid until the batch after the next one is requested."""
import torch


class DevicePrefetcher:
    def __init__(self, batches, device=None, depth=2):
        self.it = iter(batches)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.bufs = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [None] * depth         # recorded on the compute stream when a slot's batch has been handed out
        self.slot = 0
        self.pending = None
        self._stage()

    def _stage(self):
        """Start the copy of the next host batch into the next slot (waits until the slot's previous user is done)."""
        try:
            host = next(self.it)
        except StopIteration:
            self.pending = None
            return
        k = self.slot
        if self.bufs[k] is None or any(self.bufs[k][n].shape != t.shape or self.bufs[k][n].dtype != t.dtype
                                       for n, t in host.items()):
            self.bufs[k] = {n: torch.empty(t.shape, dtype=t.dtype, device=self.device) for n, t in host.items()}
        if self.consumed[k] is not None:
            self.copy_stream.wait_event(self.consumed[k])
        with torch.cuda.stream(self.copy_stream):
            for n, t in host.items():
                self.bufs[k][n].copy_(t, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.pending = k
        self.slot = (k + 1) % self.depth

    def __iter__(self):
        return self

    def __next__(self):
        if self.pending is None:
            raise StopIteration
        k = self.pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[k])
        batch = self.bufs[k]
        # the slot two batches back is free once everything enqueued on the compute stream so far has run
        prev = (k - 1) % self.depth
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[prev] = ev
        self._stage()
        return batch
