"""T-sharded memory manager: the WRITE side of SURVEY.md section 8e (the read side is `tshard.ShardedReader`).

The memory banks of ONE long video are distributed over the ranks of a process group by stored block: the b-th block
ever added to a bank lives on rank `b % world` (working and permanent frames are blocks of HW columns, long-term
prototype batches are blocks of `num_prototypes` columns).  Every rank runs the same `add_memory` calls with the same
arguments (SPMD: the encoders are replicated), stores only the blocks it owns, and mirrors the group bookkeeping of the
others with zero-width adds, so a local column range keeps the meaning "everything added since the group appeared".

Consolidation (reference inference/memory_manager.py:316-390) needs all candidate columns in their original order: the
ranks all-gather their local candidates (keys, shrinkage, selection, usage, values), interleave them back into global
frame order, run the reference computation redundantly (identical inputs -> identical prototypes on every rank), and
the rank that owns the new prototype block keeps it.  Least-used eviction of long-term memory
(reference inference/kv_memory_store.py:160-181) uses a usage threshold computed over the gathered usage of all shards.

Restrictions: one object group while long-term consolidation is active (the reference already restricts eviction to
one group, kv_memory_store.py:171-176); all working frames have the same HW.

Status: host logic verified on CPU with a world-size-2 gloo group against the single-process `MemoryManager`
(tests/test_sharded_memory_gloo.py: the union of the shards equals the single-process banks after every step).
`InferenceCore` uses this manager when its config has `t_shard=True` (every rank then runs the same video SPMD, CUDA
graphs off); that end-to-end combination has NOT run on GPUs yet.  The sharded READ (`match_memory` below) goes through
`ShardedReader.read`, whose kernels + collectives are covered by tests/test_gpu_tshard.py on hand-built shards.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from .. import lib
from .memory_manager import MemoryManager
from .tshard import ShardedReader


def _empty(t):
    return t[..., 0:0] if t is not None else None


class ShardedMemoryManager(MemoryManager):
    def __init__(self, config, group=None):
        super().__init__(config)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._blocks = {'temp': 0, 'perm': 0, 'long': 0}   # blocks ever added per bank (same on every rank)
        self._temp_frames: List[int] = []                  # block ids of ALL working frames still stored, oldest first
        self._perm_frames: List[int] = []                  # block ids of all permanent frames
        self._long_blocks: List[List[int]] = []            # per prototype block: [owner rank, surviving columns]
        self._reader = None

    # ------------------------------------------------------------------ global <-> local bookkeeping
    def _owner_of_next(self, bank: str) -> int:
        return self._blocks[bank] % self.world

    def _local(self, frames: List[int]) -> List[int]:
        return [f for f in frames if f % self.world == self.rank]

    @property
    def global_temp_size(self) -> int:
        return len(self._temp_frames) * (self.HW or 0)

    @property
    def global_long_size(self) -> int:
        return sum(c for _, c in self._long_blocks)

    def local_positions(self, bank: str) -> torch.Tensor:
        """global column position (in the single-process bank) of every local column of 'temp' | 'perm' | 'long'."""
        pos, cursor = [], 0
        if bank == 'long':
            for owner, cols in self._long_blocks:
                if owner == self.rank:
                    pos.extend(range(cursor, cursor + cols))
                cursor += cols
        else:
            for f in (self._temp_frames if bank == 'temp' else self._perm_frames):
                if f % self.world == self.rank:
                    pos.extend(range(cursor, cursor + self.HW))
                cursor += self.HW
        return torch.tensor(pos, dtype=torch.long)

    # ------------------------------------------------------------------ write (reference memory_manager.py:212-281)
    def add_memory(self, key, shrinkage, value, objects, selection=None, permanent=False, ignore=False, ti=None):
        if self.H is None or self.reset_config:
            self.reset_config = False
            self.H, self.W = key.shape[-2:]
            self.HW = self.H * self.W
            if self.enable_long_term:
                self.min_work_elements = self.min_mt_frames * self.HW
                self.max_work_elements = self.max_mt_frames * self.HW
                self.temporary_work_mem._reserve = -(-self.max_mt_frames // self.world) * self.HW + self.HW
        assert key.shape[-2] * key.shape[-1] == self.HW, 'T-sharded banks need frames of one size'
        key = key.flatten(start_dim=2)
        shrinkage = shrinkage.flatten(start_dim=2)
        value = value[0].flatten(start_dim=2)
        self.CK, self.CV = key.shape[1], value.shape[1]
        if selection is not None:
            selection = selection.flatten(start_dim=2)

        if not ignore:
            bank = 'perm' if permanent else 'temp'
            store = self.permanent_work_mem if permanent else self.temporary_work_mem
            if self._owner_of_next(bank) == self.rank:
                store.add(key, value, shrinkage, selection, objects)
            else:
                # not my block: mirror the object-group bookkeeping only
                store.add(_empty(key), _empty(value), _empty(shrinkage), _empty(selection), objects)
            if permanent:
                self._perm_frames.append(self._blocks['perm'])
                if ti is not None:
                    # the block position the single-process store reports (kv_memory_store.py:92, float floor included)
                    n_after = len(self._perm_frames) * self.HW
                    self.frame_id_to_permanent_mem_idx[ti] = int((n_after + 1e-9) // (self.HW + 1e-9)) - 1
            else:
                self._temp_frames.append(self._blocks['temp'])
            self._blocks[bank] += 1

        n_temp, n_perm = self.temporary_work_mem.num_groups, self.permanent_work_mem.num_groups
        if not self.temporary_work_mem.engaged() or n_temp != n_perm:
            target = self.temporary_work_mem if n_perm > n_temp else self.permanent_work_mem
            target.add(_empty(key), _empty(value), _empty(shrinkage), _empty(selection), objects)

        if self.enable_long_term and self.global_temp_size >= self.max_work_elements:
            if self.global_long_size >= (self.max_long_elements - self.num_prototypes):
                self.remove_obsolete_features(self.max_long_elements - self.num_prototypes)
            self.compress_features()

    # ------------------------------------------------------------------ permanent-memory editing (reference :192-210)
    def update_permanent_memory(self, frame_idx, key, shrinkage, value, selection=None):
        block = self.frame_id_to_permanent_mem_idx[frame_idx]          # position in the GLOBAL permanent bank
        if self._perm_frames[block] % self.world != self.rank:
            return
        local_pos = len(self._local(self._perm_frames[:block]))
        key = key.flatten(start_dim=2)
        shrinkage = shrinkage.flatten(start_dim=2)
        value = value[0].flatten(start_dim=2)
        if selection is not None:
            selection = selection.flatten(start_dim=2)
        self.permanent_work_mem.replace_at(local_pos, key, value, shrinkage, selection)

    def remove_from_permanent_memory(self, frame_idx):
        raise NotImplementedError('removing annotated frames from a T-sharded permanent bank')

    # ------------------------------------------------------------------ collectives
    def _all_gather_cols(self, t: torch.Tensor, max_cols: int) -> List[torch.Tensor]:
        """all-gather a [..., n_local] tensor, zero-padded to max_cols columns; one padded tensor per rank."""
        pad = torch.zeros(t.shape[:-1] + (max_cols,), dtype=t.dtype, device=t.device)
        pad[..., :t.shape[-1]] = t
        if self.world == 1:
            return [pad]
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(out, pad, group=self.group)
        return out

    # ------------------------------------------------------------------ consolidation (reference :316-390)
    def compress_features(self):
        temp = self.temporary_work_mem
        if temp.num_groups > 1:
            raise NotImplementedError('T-sharded consolidation supports a single object group')
        HW = self.HW
        n_cand = len(self._temp_frames) - self.min_mt_frames          # the oldest frames leave working memory
        cand_frames = self._temp_frames[:n_cand]
        n_loc = len(self._local(cand_frames)) * HW                    # they are the oldest columns of the local shard too
        max_cols = max(1, -(-n_cand // self.world)) * HW
        parts = [self._all_gather_cols(t[:, :, :n_loc], max_cols) if t is not None else None
                 for t in (temp.k, temp.s, temp.e, temp.get_usage(), temp.v[0])]

        def in_frame_order(per_rank):
            if per_rank is None:
                return None
            taken, cols = [0] * self.world, []
            for f in cand_frames:
                r = f % self.world
                cols.append(per_rank[r][..., taken[r] * HW:(taken[r] + 1) * HW])
                taken[r] += 1
            return torch.cat(cols, -1)

        ck, cs, ce, cu, cv = (in_frame_order(p) for p in parts)
        prototype_key, prototype_value, prototype_shrinkage = self.consolidation(ck, cs, ce, cu, [cv])

        if n_loc > 0:
            temp.sieve_by_range(0, n_loc, min_size=0)
        self._temp_frames = self._temp_frames[n_cand:]

        owner = self._owner_of_next('long')
        if owner != self.rank:
            prototype_key, prototype_shrinkage = _empty(prototype_key), _empty(prototype_shrinkage)
            prototype_value = [_empty(v) for v in prototype_value]
        n_proto = self.num_prototypes
        self.long_mem.add(prototype_key, prototype_value, prototype_shrinkage, selection=None, objects=None,
                          group_objects=temp.obj_groups)
        self._long_blocks.append([owner, n_proto])
        self._blocks['long'] += 1

    def remove_obsolete_features(self, max_size: int):
        """evict the globally least-used long-term columns (strict '>' on the threshold like the reference)."""
        long = self.long_mem
        per_rank = [sum(c for o, c in self._long_blocks if o == r) for r in range(self.world)]
        local_u = long.get_usage().flatten()
        assert local_u.numel() == per_rank[self.rank]
        parts = self._all_gather_cols(local_u.view(1, 1, -1), max(per_rank + [1]))
        everyone = torch.cat([parts[r][0, 0, :per_rank[r]] for r in range(self.world)])
        values, _ = torch.topk(everyone, k=self.global_long_size - max_size, largest=False, sorted=True)
        threshold = values[-1]
        if local_u.numel():
            long.keep_columns(local_u > threshold)
        # every rank tracks how many columns of every block survive
        taken = [0] * self.world
        for blk in self._long_blocks:
            owner, cols = blk
            u = parts[owner][0, 0, taken[owner]:taken[owner] + cols]
            taken[owner] += cols
            blk[1] = int((u > threshold).sum())

    # ------------------------------------------------------------------ read (reference memory_manager.py:61-190)
    def match_memory(self, query_key, selection, disable_usage_updates=False):
        """Exact global top-k softmax readout over all shards: the staged kernels on the local columns + three NCCL
        collectives (tshard.py).  Same signature/result as `MemoryManager.match_memory`; every rank gets the full readout.
        NOT yet run on GPUs in this form (the kernels + collectives are covered by tests/test_gpu_tshard.py on
        hand-built shards); a rank whose shard is still empty relies on the kernels' zero-tile path."""
        lib.require_cuda(query_key, 'query_key')
        if selection is None:
            raise NotImplementedError('the fused read kernel needs the selection term (enable_long_term or need_segment path)')
        if self._reader is None:
            self._reader = ShardedReader(self.group)
        h, w = query_key.shape[-2:]
        hw = h * w
        hw_pad = (hw + 127) // 128 * 128
        dev = query_key.device
        a, n_obj, use_long, _ = self.refresh_plan(dev, disable_usage_updates)
        if a.n_groups != 1:
            raise NotImplementedError('T-sharded read supports a single object group')
        wsb = self._ensure_ws(hw, n_obj, dev)
        krows = query_key[0].permute(1, 2, 0).reshape(hw, -1).to(torch.float16).contiguous()
        erows = selection[0].permute(1, 2, 0).reshape(hw, -1).to(torch.float16).contiguous()
        qp, bsq = lib.query_pack(krows, erows, hw_pad)
        out = torch.empty((n_obj, hw, lib.CV), dtype=torch.float16, device=dev)
        a.qp, a.bsq, a.hw, a.hw_pad = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), wsb
        a.plan_is_resident = 0                      # the staged entry points build and upload the plan themselves
        self._reader.read(a, out)
        if self.enable_long_term and not disable_usage_updates:
            self.temporary_work_mem.tick_life()
            if use_long and self.enable_long_term_usage:
                self.long_mem.tick_life()
        return out.view(n_obj, h, w, lib.CV).permute(0, 3, 1, 2)

    # ------------------------------------------------------------------ description of the local shard for the read kernels
    def read_args(self):
        """(XmAffinityArgs of THIS rank's shard, n_obj): fill in the query fields and hand to `ShardedReader.read`."""
        a, n_obj, _ = self._read_args()
        return a, n_obj
