"""Three-bank memory manager on the fused B200 read kernel — drop-in for the reference
`inference/memory_manager.py:8-425` (same constructor, methods, attributes).

`match_memory` is ONE call into libxmem2_b200.so (xm_affinity_readout): similarity over
long-term | working | permanent banks, per-group top-k softmax, value readout and usage accumulation,
with no bank concatenation and no N x HW intermediate.  The banks are arena-backed
(`KeyValueMemoryStore`), so `add_memory` is an in-place append.
"""
from __future__ import annotations

import ctypes as C
import warnings

import torch

from .. import lib
from .kv_memory_store import KeyValueMemoryStore


class MemoryManager:
    def __init__(self, config):
        self.config = config
        self.hidden_dim = config['hidden_dim']
        self.top_k = config['top_k']
        self.enable_long_term = config['enable_long_term']
        self.enable_long_term_usage = config['enable_long_term_count_usage']
        if self.enable_long_term:
            self.max_mt_frames = config['max_mid_term_frames']
            self.min_mt_frames = config['min_mid_term_frames']
            self.num_prototypes = config['num_prototypes']
            self.max_long_elements = config['max_long_term_elements']
        self.CK = self.CV = None
        self.H = self.W = None
        self.hidden = None                       # [1, n_obj, CH, H, W] fp32 (NHWC-backed view)
        self.temporary_work_mem = KeyValueMemoryStore(count_usage=self.enable_long_term)
        self.permanent_work_mem = KeyValueMemoryStore(count_usage=False)
        self.frame_id_to_permanent_mem_idx = dict()
        if self.enable_long_term:
            self.long_mem = KeyValueMemoryStore(count_usage=self.enable_long_term_usage,
                                                reserve=self.max_long_elements + self.num_prototypes)
        self.reset_config = True
        self._ws = None
        self._plan_key = None
        self._plan_host = None
        self.HW = None

    def update_config(self, config):
        self.reset_config = True
        self.hidden_dim = config['hidden_dim']
        self.top_k = config['top_k']
        assert self.enable_long_term == config['enable_long_term'], 'cannot update this'
        assert self.enable_long_term_usage == config['enable_long_term_count_usage'], 'cannot update this'
        if self.enable_long_term:
            self.max_mt_frames = config['max_mid_term_frames']
            self.min_mt_frames = config['min_mid_term_frames']
            self.num_prototypes = config['num_prototypes']
            self.max_long_elements = config['max_long_term_elements']

    # ------------------------------------------------------------------ memory read (reference :61-190)
    def _read_args(self, n_obj_hint=None):
        """XmAffinityArgs describing the three banks and the object groups (no query yet)."""
        temp, perm = self.temporary_work_mem, self.permanent_work_mem
        use_long = self.enable_long_term and self.long_mem.engaged() and self.long_mem.size > 0
        num_groups = max(temp.num_groups, perm.num_groups)
        if num_groups == 0:
            raise RuntimeError('match_memory called with an empty memory')
        n_obj = len(perm.all_objects) if perm.num_groups >= temp.num_groups else len(temp.all_objects)
        a = lib.XmAffinityArgs()
        if use_long:
            self.long_mem.bank_struct(a.banks[0], with_usage=self.enable_long_term_usage)
        else:
            a.banks[0].size = 0
        temp.bank_struct(a.banks[1], with_usage=self.enable_long_term)
        perm.bank_struct(a.banks[2], with_usage=False)
        a.n_groups = num_groups
        for gi in range(num_groups):
            ref = perm if gi < perm.num_groups else temp
            grp = ref.obj_groups[gi]
            assert grp == list(range(grp[0], grp[0] + len(grp))), 'object groups must be contiguous index ranges'
            g = a.groups[gi]
            g.obj_begin, g.n_obj = grp[0], len(grp)
            g.begin[0] = self.long_mem.group_begin(gi) if (use_long and gi < self.long_mem.num_groups) else (self.long_mem.size if use_long else 0)
            g.begin[1] = temp.group_begin(gi) if gi < temp.num_groups else temp.size
            g.begin[2] = perm.group_begin(gi) if gi < perm.num_groups else perm.size
        a.top_k, a.n_obj_total = self.top_k, n_obj
        return a, n_obj, use_long

    def layout_signature(self):
        """Everything a recorded CUDA graph of the read depends on: arena addresses/capacities and group structure
        (NOT the bank sizes — those live in the device-side plan)."""
        a, n_obj, use_long = self._read_args()
        banks = tuple((a.banks[i].keys, a.banks[i].values, a.banks[i].cap, a.banks[i].n_obj_cap, bool(a.banks[i].usage))
                      for i in range(3))
        groups = tuple((a.groups[g].obj_begin, a.groups[g].n_obj) for g in range(a.n_groups))
        return banks, groups, n_obj, self.top_k

    def refresh_plan(self, device, disable_usage_updates=False):
        """Bank/group description for the kernel call plus the key that tells whether the device-side plan (column
        ranges, usage routing) is stale."""
        a, n_obj, use_long = self._read_args()
        if disable_usage_updates or not self.enable_long_term:
            for i in range(3):
                a.banks[i].usage = None
        key = (tuple((a.banks[i].keys, a.banks[i].shrinkage, a.banks[i].values, a.banks[i].usage, a.banks[i].cap, a.banks[i].size)
                     for i in range(3)),
               tuple((a.groups[g].obj_begin, a.groups[g].n_obj, tuple(a.groups[g].begin)) for g in range(a.n_groups)), self.top_k)
        return a, n_obj, use_long, key

    def _ensure_ws(self, hw, n_obj, device):
        wsb = lib.load().xm_affinity_workspace_bytes(hw, n_obj)
        if self._ws is None or self._ws.numel() < wsb or self._ws.device != device:
            self._ws = lib.affinity_workspace(hw, n_obj, device)
            self._plan_key = None
        return wsb

    def adopt_workspace(self, ws):
        """Use `ws` (the workspace a recorded CUDA graph points into) from now on; the device-side plan is re-uploaded."""
        self._ws = ws
        self._plan_key = None

    def upload_plan(self, hw, device, disable_usage_updates=False):
        a, n_obj, use_long, key = self.refresh_plan(device, disable_usage_updates)
        wsb = self._ensure_ws(hw, n_obj, device)
        if key != self._plan_key:
            host = torch.empty(4096, dtype=torch.uint8).pin_memory()
            lib.check(lib.load().xm_affinity_plan(C.byref(a), host.data_ptr(), 4096), 'xm_affinity_plan')
            self._ws[:4096].copy_(host, non_blocking=True)
            self._plan_host = host                     # keep alive until the copy has certainly run (next change)
            self._plan_key = key
        return a, n_obj, use_long, wsb

    def match_memory(self, query_key, selection, disable_usage_updates=False):
        """query_key, selection: [1, CK, h, w] -> readout [n_obj, CV, h, w] (fp16, NHWC-backed view)."""
        lib.require_cuda(query_key, 'query_key')
        if selection is None:
            raise NotImplementedError('the fused read kernel needs the selection term (enable_long_term or need_segment path)')
        h, w = query_key.shape[-2:]
        hw = h * w
        hw_pad = (hw + 127) // 128 * 128
        dev = query_key.device
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            # plan already uploaded by the caller (InferenceCore graph path); only describe the banks
            a, n_obj, use_long, _ = self.refresh_plan(dev, disable_usage_updates)
            wsb = lib.load().xm_affinity_workspace_bytes(hw, n_obj)
            assert self._ws is not None and self._ws.numel() >= wsb
        else:
            a, n_obj, use_long, wsb = self.upload_plan(hw, dev, disable_usage_updates)
        krows = query_key[0].permute(1, 2, 0).reshape(hw, -1)
        erows = selection[0].permute(1, 2, 0).reshape(hw, -1)
        krows = krows.to(torch.float16).contiguous(); erows = erows.to(torch.float16).contiguous()
        qp, bsq = lib.query_pack(krows, erows, hw_pad)
        out = torch.empty((n_obj, hw, lib.CV), dtype=torch.float16, device=dev)
        a.qp, a.bsq, a.hw, a.hw_pad = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad
        a.readout_chw, a.readout_hwc = None, out.data_ptr()
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), wsb
        a.plan_is_resident = 1
        lib.check(lib.load().xm_affinity_readout(C.byref(a), lib.stream_ptr()), 'xm_affinity_readout')
        if self.enable_long_term and not disable_usage_updates:
            # use_count was accumulated by the kernel; life_count += 1 (kv_memory_store.py:103)
            self.temporary_work_mem.tick_life()
            if use_long and self.enable_long_term_usage:
                self.long_mem.tick_life()
        return out.view(n_obj, h, w, lib.CV).permute(0, 3, 1, 2)

    # ------------------------------------------------------------------ permanent-memory editing (reference :192-210)
    def update_permanent_memory(self, frame_idx, key, shrinkage, value, selection=None):
        saved_pos = self.frame_id_to_permanent_mem_idx[frame_idx]
        key = key.flatten(start_dim=2)
        shrinkage = shrinkage.flatten(start_dim=2)
        value = value[0].flatten(start_dim=2)
        if selection is not None:
            selection = selection.flatten(start_dim=2)
        self.permanent_work_mem.replace_at(saved_pos, key, value, shrinkage, selection)

    def remove_from_permanent_memory(self, frame_idx):
        # NOTE reference quirk: the frame *position* is used as an element offset.
        elem_size = self.HW
        saved_pos = self.frame_id_to_permanent_mem_idx[frame_idx]
        self.permanent_work_mem.remove_at(saved_pos, elem_size)
        del self.frame_id_to_permanent_mem_idx[frame_idx]

    # ------------------------------------------------------------------ memory write (reference :212-281)
    def add_memory(self, key, shrinkage, value, objects, selection=None, permanent=False, ignore=False, ti=None):
        """key [1,CK,h,w], shrinkage [1,1,h,w], value [1,n_obj,CV,h,w], selection [1,CK,h,w]."""
        if self.H is None or self.reset_config:
            self.reset_config = False
            self.H, self.W = key.shape[-2:]
            self.HW = self.H * self.W
            if self.enable_long_term:
                self.min_work_elements = self.min_mt_frames * self.HW
                self.max_work_elements = self.max_mt_frames * self.HW
                self.temporary_work_mem._reserve = self.max_work_elements + self.HW
        key = key.flatten(start_dim=2)
        shrinkage = shrinkage.flatten(start_dim=2)
        value = value[0].flatten(start_dim=2)
        self.CK = key.shape[1]
        self.CV = value.shape[1]
        if selection is not None:
            if not self.enable_long_term:
                warnings.warn('the selection factor is only needed in long-term mode', UserWarning)
            selection = selection.flatten(start_dim=2)

        if ignore:
            pass
        elif permanent:
            pos = self.permanent_work_mem.add(key, value, shrinkage, selection, objects)
            if ti is not None:
                self.frame_id_to_permanent_mem_idx[ti] = pos
        else:
            self.temporary_work_mem.add(key, value, shrinkage, selection, objects)

        n_temp, n_perm = self.temporary_work_mem.num_groups, self.permanent_work_mem.num_groups
        if not self.temporary_work_mem.engaged() or n_temp != n_perm:
            # engage the other bank with a zero-width block so both banks know every object group (:250-267)
            z = lambda t: t[..., 0:0] if t is not None else None
            target = self.temporary_work_mem if n_perm > n_temp else self.permanent_work_mem
            target.add(z(key), z(value), z(shrinkage), z(selection), objects)

        if self.enable_long_term and self.temporary_work_mem.size >= self.max_work_elements:
            if self.long_mem.size >= (self.max_long_elements - self.num_prototypes):
                self.long_mem.remove_obsolete_features(self.max_long_elements - self.num_prototypes)
            self.compress_features()

    # ------------------------------------------------------------------ hidden state (reference :283-300)
    def create_hidden_state(self, n, sample_key):
        h, w = sample_key.shape[-2:]
        dev = sample_key.device
        if self.hidden is None:
            self.hidden = torch.zeros((1, n, h, w, self.hidden_dim), device=dev).permute(0, 1, 4, 2, 3)
        elif self.hidden.shape[1] != n:
            old = self.hidden.permute(0, 1, 3, 4, 2)
            grown = torch.zeros((1, n, h, w, self.hidden_dim), device=dev)
            grown[:, :old.shape[1]] = old
            self.hidden = grown.permute(0, 1, 4, 2, 3)
        assert self.hidden.shape[1] == n

    def set_hidden(self, hidden):
        self.hidden = hidden

    def get_hidden(self):
        return self.hidden

    def frame_already_saved(self, ti):
        return ti in self.frame_id_to_permanent_mem_idx

    # ------------------------------------------------------------------ consolidation (reference :316-390)
    def compress_features(self):
        """Move the oldest working-memory columns into `num_prototypes` long-term prototypes (reference :316-347): prototype
        selection, similarity, softmax and read-outs in csrc/consolidate.cu, the sieve as an in-arena compaction kernel."""
        HW = self.HW
        temp = self.temporary_work_mem
        total = temp.size
        m = self.min_work_elements
        candidate_value = []
        for gv in temp.value:
            ng = gv.shape[-1]
            if ng == total:
                candidate_value.append(gv[:, :, :ng - m])
            else:
                assert HW <= ng < total
                candidate_value.append(gv[:, :, :ng - m] if ng > m else None)
        k, sk, ek, usage = temp.get_all_sliced(0, -m)
        prototype_key, prototype_value, prototype_shrinkage = self.consolidation(k, sk, ek, usage, candidate_value)
        temp.sieve_by_range(0, -m, min_size=m + HW)
        self.long_mem.add(prototype_key, prototype_value, prototype_shrinkage, selection=None, objects=None,
                          group_objects=temp.obj_groups)

    def consolidation(self, candidate_key, candidate_shrinkage, candidate_selection, usage, candidate_value):
        """Reference signature (:349-390).  candidate_key [1,CK,N], candidate_shrinkage [1,1,N], candidate_selection [1,CK,N] or
        None, usage [1,1,N], candidate_value: per group [n_g,CV,N_g] (the group's LAST N_g candidates) or None.
        Prototype selection, similarity + full softmax, value and shrinkage read-out run as CUDA kernels (csrc/consolidate.cu);
        returns (prototype_key [1,CK,P], [prototype_value_g [n_g,CV,P_g] or None], prototype_shrinkage [1,1,P])."""
        dev = candidate_key.device
        lib.require_cuda(candidate_key, 'candidate_key')
        N = candidate_key.shape[-1]
        P = self.num_prototypes
        u = usage.reshape(-1).float().contiguous()
        proto = lib.usage_topk(u, torch.ones_like(u), P)
        # candidates in the kernel's layout: packed key rows, fp32 shrinkage, fp16 selection rows
        krows = candidate_key[0].transpose(0, 1).to(torch.float16).contiguous()
        kp = torch.empty((N, 2 * lib.CK), dtype=torch.float16, device=dev)
        lib.key_pack(krows, kp)
        s = candidate_shrinkage.reshape(-1).float().contiguous()
        e = candidate_selection[0].transpose(0, 1).to(torch.float16).contiguous() if candidate_selection is not None else None
        n_pad = (N + 7) // 8 * 8
        aff = torch.empty((P, n_pad), dtype=torch.float32, device=dev)
        shr_out = torch.empty(P, dtype=torch.float32, device=dev)
        prototype_value = []
        first = True
        for gv in candidate_value:
            if gv is None:
                prototype_value.append(None)
                continue
            ng = gv.shape[2]
            col_begin = N - ng
            lib.consolidate_affinity(kp, s, e, proto, col_begin, aff, shr_out if first else None)
            first = False
            if col_begin == 0:
                valid, n_valid = None, P
            else:       # prototypes outside this group's columns do not exist for it (:357-359); the host needs the count
                valid = torch.nonzero(proto >= col_begin).flatten().to(torch.int32)
                n_valid = int(valid.numel())
            if n_valid == 0:
                prototype_value.append(None)
                continue
            if not (gv.stride(2) == 1 and gv.stride(0) == gv.shape[1] * gv.stride(1) and gv.dtype == torch.float16):
                gv = gv.to(torch.float16).contiguous()
            prototype_value.append(lib.consolidate_values(gv, aff, col_begin, valid, n_valid))
        prototype_key = candidate_key[:, :, proto.long()]
        prototype_shrinkage = shr_out.view(1, 1, P) if candidate_shrinkage is not None else None
        return prototype_key, prototype_value, prototype_shrinkage

    # ------------------------------------------------------------------ GUI helper (reference :392-425)
    def copy_perm_mem_only(self):
        new_mem = MemoryManager(config=self.config)
        perm = self.permanent_work_mem
        if perm.key is None or perm.key.size(-1) == 0:
            return new_mem
        new_mem.permanent_work_mem = perm
        new_mem.frame_id_to_permanent_mem_idx = self.frame_id_to_permanent_mem_idx
        z = lambda t: t[..., 0:0] if t is not None else None
        objects = [o + 1 for o in perm.all_objects]
        value0 = torch.zeros((len(perm.all_objects), lib.CV, 0), dtype=torch.float16, device=perm.key.device)
        new_mem.temporary_work_mem.add(z(perm.key), value0, z(perm.shrinkage), z(perm.selection), objects)
        new_mem.temporary_work_mem.obj_groups = [list(g) for g in perm.obj_groups]
        new_mem.temporary_work_mem._group_begin = [0 for _ in perm.obj_groups]
        new_mem.CK, new_mem.CV = self.CK, self.CV
        new_mem.H, new_mem.W, new_mem.HW = self.H, self.W, self.HW
        sample_key = perm.key[..., 0:self.HW].reshape(1, -1, self.H, self.W)
        new_mem.create_hidden_state(len(perm.all_objects), sample_key)
        new_mem.reset_config = True
        return new_mem
