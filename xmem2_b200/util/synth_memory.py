"""Synthetic memory banks for benchmarks and tests of the fused read kernel: random keys / shrinkage / values in the kernel's
arena layout (include/xmem2_b200.h, xm_bank_t) plus a random query, and the XmAffinityArgs that describe them."""
from __future__ import annotations

import ctypes as C

import torch

from .. import lib

CK, CV = lib.CK, lib.CV


def make_case(hw, sizes, n_obj, group_begins, seed=0, device='cuda', key_scale=0.6):
    """sizes = (long, work, perm) columns; group_begins = list of (obj_begin, n_obj, [b_long, b_work, b_perm])."""
    g = torch.Generator().manual_seed(seed)
    banks = []
    for n in sizes:
        cap = max(8, (n + 64 + 7) // 8 * 8)
        if n == 0:
            banks.append(None); continue
        key = (torch.randn(n, CK, generator=g) * key_scale).half()
        shr = (torch.rand(n, generator=g) * 2 + 1).float()
        val = torch.zeros(n_obj, CV, cap).half()
        val[:, :, :n] = torch.randn(n_obj, CV, n, generator=g).half()
        banks.append(dict(key=key, shr=shr, val=val, cap=cap, n=n))
    qk = (torch.randn(hw, CK, generator=g) * key_scale).half()
    qe = torch.rand(hw, CK, generator=g).half()
    return dict(hw=hw, banks=banks, n_obj=n_obj, groups=group_begins, qk=qk, qe=qe, device=device)


def device_args(case, top_k=30, columns=None):
    """Upload `case` and describe it for the C ABI.  columns: optional {bank index: list of column indices} to keep only a
    subset of a bank (T-sharding).  Returns (XmAffinityArgs, dict of the device tensors that must stay alive)."""
    dev = case['device']
    hw = case['hw']; hw_pad = (hw + 127) // 128 * 128
    a = lib.XmAffinityArgs()
    keep = {'usage': []}
    for bi, b in enumerate(case['banks']):
        if b is None:
            a.banks[bi].size = 0; keep['usage'].append(None); continue
        if columns is not None and bi in columns:
            idx = torch.tensor(columns[bi], dtype=torch.long)
            n = len(columns[bi]); cap = max(64, (n + 64 + 7) // 8 * 8)
            key = b['key'][idx] if n else b['key'][:0]
            shr_h = b['shr'][idx] if n else b['shr'][:0]
            val_h = b['val'][:, :, idx] if n else b['val'][:, :, :0]
        else:
            n, cap, key, shr_h, val_h = b['n'], b['cap'], b['key'], b['shr'], b['val'][:, :, :b['n']]
        rows = torch.zeros(cap, 2 * CK, dtype=torch.float16, device=dev)
        shr = torch.ones(cap, dtype=torch.float32, device=dev)
        val = torch.zeros(case['n_obj'], CV, cap, dtype=torch.float16, device=dev)
        if n:
            lib.key_pack(key.to(dev).contiguous(), rows[:n])
            shr[:n] = shr_h.to(dev)
            val[:, :, :n] = val_h.to(dev)
        usage = torch.zeros(cap, dtype=torch.float32, device=dev)
        keep[f'bank{bi}'] = (rows, shr, val); keep['usage'].append(usage)
        bk = a.banks[bi]
        bk.keys, bk.shrinkage, bk.values, bk.usage = rows.data_ptr(), shr.data_ptr(), val.data_ptr(), usage.data_ptr()
        bk.cap, bk.n_obj_cap, bk.size = cap, case['n_obj'], n
    a.n_groups = len(case['groups'])
    for gi, (ob, no, begins) in enumerate(case['groups']):
        a.groups[gi].obj_begin, a.groups[gi].n_obj = ob, no
        for bi in range(3):
            a.groups[gi].begin[bi] = begins[bi]
    qp, bsq = lib.query_pack(case['qk'].to(dev).contiguous(), case['qe'].to(dev).contiguous(), hw_pad)
    ws = lib.affinity_workspace(hw, case['n_obj'], dev)
    keep['query'] = (qp, bsq); keep['ws'] = ws
    a.qp, a.bsq, a.hw, a.hw_pad, a.top_k, a.n_obj_total = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad, top_k, case['n_obj']
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    return a, keep


def time_readout(a, iters=10, warmup=3, flush_l2=True):
    """Mean seconds per xm_affinity_readout call, CUDA events on the current stream; an L2-sized buffer is rewritten between calls."""
    L = lib.load()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=torch.cuda.current_device()) if flush_l2 else None
    times = []
    for it in range(warmup + iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(L.xm_affinity_readout(C.byref(a), lib.stream_ptr()), 'xm_affinity_readout')
        e1.record(); torch.cuda.synchronize()
        if it >= warmup:
            times.append(e0.elapsed_time(e1) * 1e-3)
    return sum(times) / len(times)
