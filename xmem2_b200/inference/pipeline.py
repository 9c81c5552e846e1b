"""Driver-side plumbing around `InferenceCore.step` for throughput: what the reference driver
does per frame in `_post_process` + `.cpu()` (inference/run_on_video.py:165-173: bilinear resize to the original size, argmax,
uint8, synchronous device -> host copy), done as one fused kernel (csrc/postproc.cu) and an OVERLAPPED copy: label maps land in
a ring of pinned host buffers, and the host only waits for a buffer when it is about to be reused.

    dl = MaskDownloader((H, W), device)
    for ti, rgb in enumerate(frames):
        prob = core.step(rgb)
        dl.submit(ti, prob)            # fused resize + argmax (+ label table) and an async D2H
        for tj, mask in dl.ready():    # numpy uint8 [H, W] of earlier frames, in order
            save(tj, mask)
    for tj, mask in dl.drain(): save(tj, mask)
"""
from __future__ import annotations

from collections import deque
from typing import Optional, Sequence

import torch

from .postprocess import post_process


class MaskDownloader:
    def __init__(self, out_shape: Sequence[int], device, depth: int = 3, label_table: Optional[torch.Tensor] = None):
        self.shape = (int(out_shape[0]), int(out_shape[1]))
        self.device = torch.device(device)
        self.table = label_table
        self.depth = depth
        self.host = [torch.empty(self.shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.dev = [torch.empty(self.shape, dtype=torch.uint8, device=self.device) for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.labelled = [torch.cuda.Event() for _ in range(depth)]
        self.stream = torch.cuda.Stream(device=self.device)      # device -> host copies run beside the next frame's kernels
        self.pending = deque()          # (frame index, slot)
        self.n = 0
        self.bytes_per_frame = self.shape[0] * self.shape[1]

    def submit(self, frame_index: int, prob: torch.Tensor):
        """prob [n_obj+1, h, w] fp32 on the device (what `InferenceCore.step` returns)."""
        out = []
        if len(self.pending) == self.depth:              # the ring is full: hand the oldest mask out first
            out.append(self._pop())
        slot = self.n % self.depth
        self.n += 1
        # slot's device buffer is free: its previous copy was waited for when the slot was popped
        post_process(prob, self.shape, self.table, out=self.dev[slot])
        self.labelled[slot].record()
        self.stream.wait_event(self.labelled[slot])
        with torch.cuda.stream(self.stream):
            self.host[slot].copy_(self.dev[slot], non_blocking=True)
            self.events[slot].record(self.stream)
        self.pending.append((frame_index, slot))
        return out

    def _pop(self):
        ti, slot = self.pending.popleft()
        self.events[slot].synchronize()
        return ti, self.host[slot].numpy().copy()

    def ready(self):
        """masks whose copy has already finished (never blocks)."""
        out = []
        while self.pending and self.events[self.pending[0][1]].query():
            out.append(self._pop())
        return out

    def drain(self):
        out = []
        while self.pending:
            out.append(self._pop())
        return out


class FrameUploader:
    """Host -> device copies of the NEXT frame on a side stream while the current frame is being segmented (the reference's
    DataLoader hands over host tensors, inference/run_on_video.py:88-104).  `depth` device buffers per shape are recycled.

        up = FrameUploader(device)
        up.prefetch(0, frames[0])
        for ti in range(n):
            if ti + 1 < n: up.prefetch(ti + 1, frames[ti + 1])     # pinned host tensor
            rgb = up.get(ti)                                        # device tensor, ordered after its copy
    """

    def __init__(self, device, depth: int = 3):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.slots = {}                 # shape/dtype -> list of device buffers
        self.inflight = {}              # key -> (buffer, event)
        self.count = 0

    def prefetch(self, key, host_tensor: torch.Tensor):
        sig = (tuple(host_tensor.shape), host_tensor.dtype)
        bufs = self.slots.setdefault(sig, [])
        if len(bufs) < self.depth:
            bufs.append(torch.empty(host_tensor.shape, dtype=host_tensor.dtype, device=self.device))
        buf = bufs[self.count % self.depth] if len(bufs) == self.depth else bufs[-1]
        self.count += 1
        # the buffer may still be read by kernels of the frame that used it `depth` frames ago
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            buf.copy_(host_tensor, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(self.stream)
        self.inflight[key] = (buf, ev)

    def get(self, key) -> torch.Tensor:
        buf, ev = self.inflight.pop(key)
        torch.cuda.current_stream(self.device).wait_event(ev)
        return buf
