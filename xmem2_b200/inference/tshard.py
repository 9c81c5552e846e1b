"""T-sharded memory read (BASELINE.json config 4): the memory banks of ONE long video are
distributed over the ranks of a process group by stored frame (rank r owns frames r, r+R, ...), every rank holds the
same query, and the exact global top-k softmax readout is assembled with three small NCCL collectives:

    stage_a (local slot maxima -> lower bound)          all_reduce(MAX)   tau_lo      [hw_pad]            fp32
    stage_b (local scores > pred(tau_lo) -> 32 largest) all_gather        top32       [R][hw_pad][32]     fp32
    merge   (exact tau, 1/den; identical on every rank)
    stage_c (local P.V with the global normalisers)     all_reduce(SUM)   readout_f32 [n_obj][hw_pad][512] fp32

No max/sum all-reduce of a softmax is needed: the top-k branch of the reference has no max subtraction
(model/memory_util.py:48-49), so the gathered candidate scores determine both the threshold and the denominator.
Usage statistics stay on the rank that owns the column.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from .. import lib

_ARGT = False


def _bind():
    global _ARGT
    if _ARGT:
        return lib.load()
    L = lib.load()
    vp, i32 = C.c_void_p, C.c_int32
    L.xm_affinity_tshard_stage_a.argtypes = [C.POINTER(lib.XmAffinityArgs), vp, vp]
    L.xm_affinity_tshard_stage_b.argtypes = [C.POINTER(lib.XmAffinityArgs), vp, vp, vp]
    L.xm_affinity_tshard_merge.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
    L.xm_affinity_tshard_stage_c.argtypes = [C.POINTER(lib.XmAffinityArgs), vp, vp, vp, vp]
    L.xm_affinity_tshard_cast.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    _ARGT = True
    return L


def frames_of_rank(n_frames: int, rank: int, world: int):
    """Stored memory frames owned by `rank` (round-robin by frame index)."""
    return list(range(rank, n_frames, world))


class ShardedReader:
    """Runs the staged read on THIS rank's shard and the collectives on `group`."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._bufs = {}

    def _buffers(self, hw_pad, n_obj, device):
        key = (hw_pad, n_obj, str(device))
        if key not in self._bufs:
            self._bufs[key] = dict(
                tau_lo=torch.empty(hw_pad, dtype=torch.float32, device=device),
                top32=torch.empty((hw_pad, 32), dtype=torch.float32, device=device),
                gathered=torch.empty((self.world, hw_pad, 32), dtype=torch.float32, device=device),
                tau=torch.empty(hw_pad, dtype=torch.float32, device=device),
                inv_den=torch.empty(hw_pad, dtype=torch.float32, device=device),
                acc=torch.empty((n_obj, hw_pad, lib.CV), dtype=torch.float32, device=device))
        return self._bufs[key]

    def read(self, args: "lib.XmAffinityArgs", out_hwc: torch.Tensor, stage_times: list = None):
        """`args`: banks/groups[0]/query/workspace of THIS rank (see MemoryManager._read_args); out_hwc [n_obj][hw][512] fp16.
        `stage_times` (diagnostics): receives the CUDA-event milliseconds of the 8 steps."""
        L = _bind()
        dev = out_hwc.device
        b = self._buffers(args.hw_pad, args.n_obj_total, dev)
        st = lib.stream_ptr()
        ev = []

        def mark():
            if stage_times is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); ev.append(e)

        mark()
        lib.check(L.xm_affinity_tshard_stage_a(C.byref(args), b['tau_lo'].data_ptr(), st), 'tshard_stage_a')
        mark()
        if self.world > 1:
            dist.all_reduce(b['tau_lo'], op=dist.ReduceOp.MAX, group=self.group)
        mark()
        lib.check(L.xm_affinity_tshard_stage_b(C.byref(args), b['tau_lo'].data_ptr(), b['top32'].data_ptr(), st), 'tshard_stage_b')
        mark()
        if self.world > 1:
            dist.all_gather_into_tensor(b['gathered'], b['top32'], group=self.group)
            gathered = b['gathered']
        else:
            gathered = b['top32']
        mark()
        lib.check(L.xm_affinity_tshard_merge(gathered.data_ptr(), self.world, args.hw, args.hw_pad, args.top_k, b['tau'].data_ptr(),
                                             b['inv_den'].data_ptr(), st), 'tshard_merge')
        lib.check(L.xm_affinity_tshard_stage_c(C.byref(args), b['tau'].data_ptr(), b['inv_den'].data_ptr(), b['acc'].data_ptr(), st),
                  'tshard_stage_c')
        mark()
        if self.world > 1:
            dist.all_reduce(b['acc'], op=dist.ReduceOp.SUM, group=self.group)
        mark()
        lib.check(L.xm_affinity_tshard_cast(b['acc'].data_ptr(), args.n_obj_total, args.hw, args.hw_pad, None, out_hwc.data_ptr(), st),
                  'tshard_cast')
        mark()
        if stage_times is not None:
            torch.cuda.synchronize()
            stage_times[:] = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]
        return out_hwc, b['tau']
