"""Host -> device input staging for the training loop (the `imgs.to(device)` of utils/engine.py:222-243).

`DevicePrefetcher` walks an iterable of pinned host batches and keeps ONE batch in flight on a copy
stream, so the host->device transfer of step i+1 overlaps the kernels of step i.  Every batch is still
copied exactly once; nothing is cached.  The device side is two fixed slots (no allocation per step):
the copy into a slot is ordered after the step that last read it."""
import torch


class DevicePrefetcher:
    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.slots = [None, None]
        self.k = 0
        self._next = None
        self._preload()

    def _preload(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        if self.stream is None:
            self._next = tuple(t.to(self.device) for t in host)
            return
        slot = self.slots[self.k]
        if slot is None or any(s.shape != h.shape or s.dtype != h.dtype for s, h in zip(slot, host)):
            # same strides as the host tensor (e.g. channels_last_3d clips stay channels_last_3d)
            slot = tuple(torch.empty_strided(h.shape, h.stride(), dtype=h.dtype, device=self.device) for h in host)
            self.slots[self.k] = slot
        # everything enqueued so far on the compute stream (incl. the step that read this slot) precedes the copy
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            for s, h in zip(slot, host):
                s.copy_(h, non_blocking=True)
        self._next = slot
        self.k ^= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        batch = self._next
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self._preload()                  # enqueue the next copy before the caller launches this step
        return batch
