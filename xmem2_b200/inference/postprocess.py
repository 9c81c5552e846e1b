"""Driver-side post-processing on the GPU.

`post_process(prob, shape)` is the fused equivalent of the reference driver's `_post_process`
(inference/run_on_video.py:165-173): bilinear resize of the probabilities to the original frame size
(`F.interpolate(..., mode='bilinear', align_corners=False)`), argmax over the objects, uint8 — one kernel
(csrc/postproc.cu), one byte per pixel written, no resized fp32 probabilities materialised; the optional `label_table`
also folds `MaskMapper.remap_index_mask` (inference/data/mask_mapper.py:56-64) into the same pass.
The result stays on the device; the caller decides when to copy it (`.cpu()`, or a pinned buffer with `non_blocking=True`
to overlap with the next frame).

The per-pixel arithmetic is verified against torch on the CPU (tests/test_postproc.py, host harness built from the same
header as the kernel); the kernel itself was written after the last GPU session of round 1 and has not run yet.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .. import lib


def label_table(remappings: dict) -> torch.Tensor:
    """256-entry uint8 table from `MaskMapper.remappings` (original label -> internal index): internal index -> label."""
    t = torch.arange(256, dtype=torch.uint8)
    for label, idx in remappings.items():
        t[int(idx)] = int(label)
    return t


def post_process(prob: torch.Tensor, shape: Optional[Sequence[int]] = None, label_table: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """prob [n_obj+1, H, W] fp32 (a strided view is fine as long as the last stride is 1) -> uint8 [shape] index mask on
    the same device.  `shape=None` keeps H x W (the `need_resize == False` branch of the reference).  `out`: optional uint8 [shape]
    device buffer to write into (callers that overlap the download with the next frame own a ring of them)."""
    lib.require_cuda(prob, 'prob')
    assert prob.dim() == 3 and prob.dtype == torch.float32 and prob.stride(2) == 1
    c, h, w = prob.shape
    oh, ow = (h, w) if shape is None else (int(shape[0]), int(shape[1]))
    if out is None:
        out = torch.empty((oh, ow), dtype=torch.uint8, device=prob.device)
    assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (oh, ow) and out.is_contiguous()
    lut = None
    if label_table is not None:
        lut = label_table.to(device=prob.device, dtype=torch.uint8).contiguous()
        assert lut.numel() == 256
    lib.check(lib.load().xm_resize_argmax(prob.data_ptr(), c, h, w, prob.stride(0), prob.stride(1), oh, ow,
                                          lut.data_ptr() if lut is not None else None, out.data_ptr(), lib.stream_ptr()),
              'xm_resize_argmax')
    return out
