"""Arena-backed key/value memory bank — drop-in for the reference `inference/kv_memory_store.py:4-239`.

The reference grows every tensor with `torch.cat` on each `add` (O(N) realloc + copy per memory frame,
kv_memory_store.py:49-56,70) and re-concatenates all banks on every frame (memory_manager.py:82-83,122-128).
Here a bank owns pre-allocated device arenas laid out for the fused read kernel (csrc/k1_affinity.cu):

    packed keys  kp  fp16 [cap][128]      row n = (k_n^2 , k_n)          (UMMA K-major operand rows)
    shrinkage    s   fp32 [cap]
    selection    e   fp16 [cap][64]
    values       v   fp16 [n_obj][512][cap]   one shared column index for all object groups; group g
                                               owns the suffix [group_begin[g], size) of the columns
    use / life   fp32 [cap]

`add` appends in place (amortised O(new columns)); `sieve_by_range` compacts in place.  The public surface
(`k, v, s, e, use_count, life_count, obj_groups, all_objects`, `size`, `num_groups`, `key/value/...`
properties, `get_v_size`, ...) returns reference-shaped VIEWS of the arenas.
"""
from __future__ import annotations

import os

from typing import List, Optional

import torch

from .. import lib

CK, CV = lib.CK, lib.CV


def _round_up(x, m):
    return (x + m - 1) // m * m


# Arenas of finished videos are recycled: a serving process handles many clips, and re-using the same device
# addresses lets InferenceCore re-use its recorded CUDA graphs (their kernels have the arena addresses baked in).
# The pool is bounded by entries per key AND by total bytes (XMEM_ARENA_POOL_MB, default 4096); `clear_arena_pool()` drops it (call
# it before torch.cuda.empty_cache() if the memory is wanted back; the recorded graphs of arenas that are gone are simply
# re-recorded).  The reference-shaped views handed out by `key` / `value` / `shrinkage` / `get_usage` alias the arena: they are
# valid while the store is alive (a recycled arena is overwritten by the next video) -- clone them to keep them.
_ARENA_POOL = {}
_ARENA_POOL_MAX = 8
_ARENA_POOL_BYTES = [0]
_ARENA_POOL_LIMIT = int(os.environ.get('XMEM_ARENA_POOL_MB', '4096')) * (1 << 20)


def _arena_bytes(arena):
    return sum(t.numel() * t.element_size() for t in arena)


def clear_arena_pool():
    """drop every pooled arena (their device memory goes back to the caching allocator)"""
    _ARENA_POOL.clear()
    _ARENA_POOL_BYTES[0] = 0


class KeyValueMemoryStore:
    def __init__(self, count_usage: bool, reserve: int = 0):
        self.count_usage = count_usage
        self.obj_groups: List[List[int]] = []
        self.all_objects: List[int] = []
        self._reserve = reserve
        self._n = 0
        self._cap = 0
        self._engaged = False
        self._group_begin: List[int] = []     # first column of each value group
        self._kp = self._s = self._e = self._v = self._use = self._life = None
        self._has_e = False
        self._dev = None

    # ------------------------------------------------------------------ arena management
    def _alloc(self, cap, n_obj, device):
        cap = _round_up(max(cap, 64), 64)
        pooled = _ARENA_POOL.get((str(device), cap, max(n_obj, 1)))
        if pooled:
            arena = pooled.pop()                          # stale contents are finite and masked by `size`
            _ARENA_POOL_BYTES[0] -= _arena_bytes(arena)
            kp, s, e, v, use, life = arena
        else:
            kp = torch.zeros((cap, 2 * CK), dtype=torch.float16, device=device)
            s = torch.ones((cap,), dtype=torch.float32, device=device)
            e = torch.zeros((cap, CK), dtype=torch.float16, device=device)
            v = torch.zeros((max(n_obj, 1), CV, cap), dtype=torch.float16, device=device)
            use = torch.zeros((cap,), dtype=torch.float32, device=device)
            life = torch.zeros((cap,), dtype=torch.float32, device=device)
        if self._kp is not None and self._n > 0:
            n = self._n
            kp[:n] = self._kp[:n]; s[:n] = self._s[:n]; e[:n] = self._e[:n]
            v[:self._v.shape[0], :, :n] = self._v[:, :, :n]
            use[:n] = self._use[:n]; life[:n] = self._life[:n]
        self._release()
        self._kp, self._s, self._e, self._v, self._use, self._life = kp, s, e, v, use, life
        self._cap, self._dev = cap, device

    def _release(self):
        if self._kp is not None:
            key = (str(self._dev), self._cap, self._v.shape[0])
            lst = _ARENA_POOL.setdefault(key, [])
            arena = (self._kp, self._s, self._e, self._v, self._use, self._life)
            nbytes = _arena_bytes(arena)
            if len(lst) < _ARENA_POOL_MAX and _ARENA_POOL_BYTES[0] + nbytes <= _ARENA_POOL_LIMIT:
                lst.append(arena)
                _ARENA_POOL_BYTES[0] += nbytes
            self._kp = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure(self, extra_cols, n_obj, device):
        need = self._n + extra_cols
        cur_obj = 0 if self._v is None else self._v.shape[0]
        if self._kp is None or need > self._cap or n_obj > cur_obj:
            cap = max(self._cap, self._reserve)
            while cap < need:
                cap = max(64, cap * 2)
            self._alloc(cap, max(n_obj, cur_obj), device)

    # ------------------------------------------------------------------ reference API: add
    def add(self, key, value, shrinkage, selection, objects: Optional[List[int]], group_objects=None):
        """key [1,CK,n] (or NHWC view thereof), value [n_obj,CV,n] tensor (objects given) or per-group list
        (objects None, long-term), shrinkage [1,1,n], selection [1,CK,n] or None.  Returns the frame position
        of the newly added block like the reference (kv_memory_store.py:92).  `group_objects` (extension) names
        the object indices of each group for the list form."""
        lib.require_cuda(key, 'memory key')
        n_new = key.shape[-1]
        dev = key.device
        cur_planes = 0 if self._v is None else self._v.shape[0]
        if objects is not None:
            assert isinstance(value, torch.Tensor)
            n_obj_needed = max(list(objects) + [0])
        else:
            assert isinstance(value, list)
            self._list_groups = self._plan_group_list(value, group_objects)
            n_obj_needed = max([cur_planes] + [g[-1] + 1 for g in self._list_groups if g is not None])
        self._ensure(n_new, n_obj_needed, dev)
        n0 = self._n
        if n_new > 0:
            rows = key[0].transpose(0, 1)                       # [n, CK]
            rows = rows if rows.is_contiguous() else rows.contiguous()
            lib.key_pack(rows.to(torch.float16), self._kp[n0:n0 + n_new])
            if shrinkage is not None:
                self._s[n0:n0 + n_new] = shrinkage.reshape(-1).float()
            if selection is not None:
                self._e[n0:n0 + n_new] = selection[0].transpose(0, 1)
                self._has_e = True
            self._use[n0:n0 + n_new] = 0
            self._life[n0:n0 + n_new] = 1e-7                    # kv_memory_store.py:38
        elif selection is not None:
            self._has_e = True
        self._engaged = True
        self._n = n0 + n_new

        if objects is not None:
            remaining = [o - 1 for o in objects]
            for grp in self.obj_groups:
                for o in grp:
                    remaining.remove(o)                         # ValueError on overlapping groups, as the reference
            if n_new > 0 and self.obj_groups:
                old = [o for grp in self.obj_groups for o in grp]
                self._write_values(value, old, n0, n_new)
            if remaining:
                new_group = list(remaining)
                if n_new > 0:
                    self._write_values(value, new_group, n0, n_new)
                self.obj_groups.append(new_group)
                self._group_begin.append(n0)
                self.all_objects.extend(new_group)
                assert sorted(self.all_objects) == self.all_objects, 'Objects MUST be inserted in sorted order '
        else:
            self._add_group_list(value, n0, n_new)
        return int((self._n + 1e-9) // (n_new + 1e-9)) - 1 if n_new > 0 else -1

    def _write_values(self, value, objs, col0, n_new):
        """value [n_obj, CV, n] (reference layout) or an NHWC-backed view of it -> arena planes `objs`."""
        if value.stride(-1) != 1 and value.stride(1) == 1 and objs == list(range(objs[0], objs[0] + len(objs))):
            src = value[objs[0]:objs[0] + len(objs)].transpose(1, 2)        # [n_obj, n, CV]
            if src.is_contiguous() and src.dtype == torch.float16:
                lib.check(lib.load().xm_value_append(src.data_ptr(), len(objs), n_new, self._v[objs[0]].data_ptr(), self._cap, col0,
                                                     lib.stream_ptr()), 'xm_value_append')
                return
        for o in objs:
            self._v[o, :, col0:col0 + n_new] = value[o]

    def _plan_group_list(self, value, group_objects):
        """object indices of every group in a per-group value list (None = group absent from this add)."""
        plan, cursor = [], 0
        for gi, gv in enumerate(value):
            if gi < self.num_groups:
                grp = self.obj_groups[gi]
            elif gv is None:
                grp = None
            elif group_objects is not None:
                grp = list(group_objects[gi])
            else:
                grp = list(range(cursor, cursor + gv.shape[0]))
            if grp is not None:
                cursor = grp[-1] + 1
            plan.append(grp)
        return plan

    def _add_group_list(self, value, n0, n_new):
        """long-term path (kv_memory_store.py:80-90): value[gi] is [n_g, CV, m_g] with m_g <= n_new, or None."""
        for gi, gv in enumerate(value):
            if gv is None:
                continue
            grp = self._list_groups[gi]
            m = gv.shape[-1]
            if gi >= self.num_groups:
                self.obj_groups.append(grp)
                self.all_objects.extend(grp)
                self._group_begin.append(self._n - m)
            else:
                old_size = n0 - self._group_begin[gi]
                if m != n_new and old_size > 0:
                    # the reference reads a group's long-term values against the LAST get_v_size(gi) columns
                    # (memory_manager.py:99-103): keep them right-aligned when fewer than n_new arrive
                    src = self._v[grp[0]:grp[-1] + 1, :, self._group_begin[gi]:n0].clone()
                    self._v[grp[0]:grp[-1] + 1, :, self._n - m - old_size:self._n - m] = src
                self._group_begin[gi] = self._n - m - old_size
            self._v[grp[0]:grp[-1] + 1, :, self._n - m:self._n] = gv.to(torch.float16)

    # ------------------------------------------------------------------ usage
    def update_usage(self, usage):
        """kv_memory_store.py:96-103 (the fused read kernel normally accumulates `use_count` itself)."""
        if not self.count_usage:
            return
        self._use[:self._n] += usage.reshape(-1)
        self._life[:self._n] += 1

    def tick_life(self):
        # whole arena (static shape: CUDA-graph friendly); slots beyond `size` are re-initialised by add()
        if self.count_usage and self._life is not None:
            self._life += 1

    def get_usage(self):
        if not self.count_usage:
            raise RuntimeError('I did not count usage!')
        return (self._use[:self._n] / self._life[:self._n]).view(1, 1, -1)

    # ------------------------------------------------------------------ editing
    def replace_at(self, start_pos: int, key, value, shrinkage=None, selection=None):
        n_new = key.shape[-1]
        a, b = start_pos * n_new, (start_pos + 1) * n_new
        rows = key[0].transpose(0, 1).contiguous().to(torch.float16)
        lib.key_pack(rows, self._kp[a:b])
        for gi, grp in enumerate(self.obj_groups):
            gv = value[gi] if isinstance(value, (list, tuple)) else value[grp]
            for j, o in enumerate(grp):
                self._v[o, :, a:b] = gv[j]
        if shrinkage is not None:
            self._s[a:b] = shrinkage.reshape(-1).float()
        if self._has_e and selection is not None:
            self._e[a:b] = selection[0].transpose(0, 1)

    def remove_at(self, start: int, elem_size: int):
        self.sieve_by_range(start, start + elem_size, min_size=0)

    def sieve_by_range(self, start: int, end: int, min_size: int):
        """keep the columns OUTSIDE [start, end) (negative `end` counts from the back; end == 0 means 'to the
        end'), reference kv_memory_store.py:125-158.  In-place compaction of every arena.  With the shared
        column index a value group simply follows its keys: a group that started inside the removed range now
        starts at `start`; `min_size` (groups too small to be consolidated keep all their columns) is implied
        because such a group lies entirely in the surviving tail."""
        n = self._n
        a = start if start >= 0 else n + start
        b = n if end == 0 else (end if end > 0 else n + end)
        a = min(max(a, 0), n); b = min(max(b, a), n)
        cut = b - a
        if cut == 0:
            return
        tail = n - b
        if tail > 0:
            self._compact(None, cut, a, a + tail)
        for gi, g0 in enumerate(self._group_begin):
            if n - g0 < min_size and g0 < b:
                raise NotImplementedError('a value group below min_size overlaps the sieved range')
            self._group_begin[gi] = g0 if g0 <= a else (a if g0 < b else g0 - cut)
        self._n = n - cut

    def _compact(self, keep_idx, shift: int, first: int, m: int):
        """columns [first, m) of every arena <- columns keep_idx[i] (or i + shift): csrc/consolidate.cu, xm_bank_compact."""
        if m > first:
            lib.bank_compact(self._kp, self._s, self._e, self._use, self._life, self._v, keep_idx, shift, first, m)

    def remove_obsolete_features(self, max_size: int):
        """evict the least-used columns (kv_memory_store.py:160-181); strict '>' so threshold ties go too.  The threshold
        (k-th smallest usage), the survivor list and the compaction run on the device; the host only reads the new size."""
        if not self.count_usage:
            raise RuntimeError('I did not count usage!')
        if self.num_groups > 1:
            raise NotImplementedError('The current data structure does not support feature removal with multiple object groups')
        n = self._n
        keep, m = lib.usage_evict_list(self._use, self._life, n, n - max_size)
        self._compact(keep, 0, 0, m)
        self._n = m
        self._group_begin = [0 for _ in self._group_begin]

    def keep_columns(self, survived):
        """in-place compaction to the columns where the boolean mask `survived` [size] is set (single value group)."""
        idx = torch.nonzero(survived).flatten().to(torch.int32)
        m = idx.numel()
        self._compact(idx, 0, 0, m)
        self._n = m
        self._group_begin = [0 for _ in self._group_begin]

    def get_all_sliced(self, start: int, end: int):
        n = self._n
        b = n if end == 0 else (end if end > 0 else n + end)
        k = self.k[:, :, start:b]
        sk = self.s[:, :, start:b]
        ek = self.e[:, :, start:b] if self._has_e else None
        usage = self.get_usage()[:, :, start:b]
        return k, sk, ek, usage

    def get_v_size(self, ni: int):
        return self._n - self._group_begin[ni]

    def group_begin(self, gi: int) -> int:
        return self._group_begin[gi]

    def engaged(self):
        return self._engaged

    # ------------------------------------------------------------------ reference-shaped views
    @property
    def size(self):
        return self._n

    @property
    def num_groups(self):
        return len(self.obj_groups)

    @property
    def k(self):
        return None if not self._engaged else self._kp[:self._n, CK:].transpose(0, 1).unsqueeze(0)

    @property
    def s(self):
        return None if not self._engaged else self._s[:self._n].view(1, 1, -1)

    @property
    def e(self):
        return None if (not self._engaged or not self._has_e) else self._e[:self._n].transpose(0, 1).unsqueeze(0)

    @property
    def v(self):
        return [self._v[grp[0]:grp[-1] + 1, :, self._group_begin[gi]:self._n] for gi, grp in enumerate(self.obj_groups)]

    @property
    def use_count(self):
        return self._use[:self._n].view(1, 1, -1) if self.count_usage and self._engaged else None

    @property
    def life_count(self):
        return self._life[:self._n].view(1, 1, -1) if self.count_usage and self._engaged else None

    key = k
    value = v
    shrinkage = s
    selection = e

    # ------------------------------------------------------------------ kernel-side description
    def bank_struct(self, bank, with_usage: bool):
        """fill an XmBank (include/xmem2_b200.h) for the fused read kernel."""
        if self._kp is None:
            bank.size = 0
            bank.keys = None
            return
        bank.keys = self._kp.data_ptr(); bank.shrinkage = self._s.data_ptr(); bank.values = self._v.data_ptr()
        bank.usage = self._use.data_ptr() if (with_usage and self.count_usage) else None
        bank.cap = self._cap; bank.n_obj_cap = self._v.shape[0]; bank.size = self._n
