"""Device input pipeline: what sits between the reference's DataLoader and the fused model.

The reference moves every batch with a pageable, synchronous `.to(device)` of normalised fp32 pixels
(utils_network.py:409-410: 77 MB per 128-image batch). `DevicePrefetcher` instead
  * keeps the batch in the form the loader can produce cheapest -- raw uint8 [B,C,H,W] bytes (19 MB) when the dataset's
    ToTensor + Normalize are moved to the GPU (PatchEmbed.set_input_normalization -> vitk_normalize_u8), or bf16 / fp32
    tensors unchanged;
  * stages it through pinned host buffers and copies it on a dedicated stream into one of two device slots, so that the
    H2D copy of batch i+1 overlaps the step on batch i; the consumer stream waits on the copy's event, and a slot is
    only overwritten after the work that read it has been enqueued behind an event.
PyTorch is used for memory, streams and events only."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Iterate (images, labels) device batches from an iterable of host batches.

        for x, y in DevicePrefetcher(loader, device):      # x: uint8 / bf16 / fp32 [B,C,H,W] on the device
            loss = trainer.step(x, y)

    `slots` device buffers per tensor rotate; host batches that are not pinned are copied into rotating pinned
    staging buffers first (DataLoader(pin_memory=True) batches are used as they are)."""

    def __init__(self, loader, device, slots: int = 2):
        self.loader, self.device, self.slots = loader, torch.device(device), max(2, int(slots))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._dev = [None] * self.slots      # (x, y) device buffers per slot
        self._pin = [None] * self.slots
        self._ready = [torch.cuda.Event() for _ in range(self.slots)]     # copy into the slot finished
        self._free = [None] * self.slots                                   # consumer of the slot enqueued
        self.bytes_per_batch = 0

    def _buffers(self, k, x, y):
        d = self._dev[k]
        if d is None or d[0].shape != x.shape or d[0].dtype != x.dtype or d[1].shape != y.shape or d[1].dtype != y.dtype:
            self._dev[k] = (torch.empty(x.shape, dtype=x.dtype, device=self.device),
                            torch.empty(y.shape, dtype=y.dtype, device=self.device))
            self._pin[k] = None
        if not (x.is_pinned() and y.is_pinned()):
            p = self._pin[k]
            if p is None:
                p = self._pin[k] = (torch.empty(x.shape, dtype=x.dtype).pin_memory(),
                                    torch.empty(y.shape, dtype=y.dtype).pin_memory())
            p[0].copy_(x)
            p[1].copy_(y)
            x, y = p
        return x, y

    def _issue(self, k, x, y):
        if self._free[k] is not None:
            self.copy_stream.wait_event(self._free[k])       # the previous reader of this slot has been enqueued
        x, y = self._buffers(k, x, y)
        with torch.cuda.stream(self.copy_stream):
            self._dev[k][0].copy_(x, non_blocking=True)
            self._dev[k][1].copy_(y, non_blocking=True)
            self._ready[k].record(self.copy_stream)
        self.bytes_per_batch = x.numel() * x.element_size() + y.numel() * y.element_size()

    def __iter__(self):
        it = iter(self.loader)
        k = 0
        try:
            x, y = next(it)
        except StopIteration:
            return
        self._issue(k, x, y)
        while True:
            cur = k
            nxt = None
            try:
                nxt = next(it)
            except StopIteration:
                pass
            if nxt is not None:
                k = (k + 1) % self.slots
                self._issue(k, *nxt)                          # overlaps the consumer's work on `cur`
            torch.cuda.current_stream(self.device).wait_event(self._ready[cur])
            yield self._dev[cur]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._free[cur] = ev
            if nxt is None:
                return

    def __len__(self):
        return len(self.loader)
