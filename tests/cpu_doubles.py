"""CPU doubles of the CUDA entry points the host-logic tests cannot run (no GPU in the build container): each one restates
what the kernel computes with torch ops, so that MemoryManager / KeyValueMemoryStore bookkeeping can be compared with the
oracle on CPU.  The kernels themselves are compared with the oracle on the GPU (tests/test_gpu_consolidate.py, clips)."""
import math

import torch

CK = 64


def key_pack(key_rows, dst_rows):
    k = key_rows.float()
    dst_rows[:, :CK] = (k * k).half()
    dst_rows[:, CK:] = key_rows


def usage_topk(use, life, k):
    r = use / life
    n = r.numel()
    # descending ratio, ties by ascending index
    order = sorted(range(n), key=lambda i: (-float(r[i]), i))
    return torch.tensor(order[:k], dtype=torch.int32)


def usage_evict_list(use, life, n, n_remove):
    r = (use[:n] / life[:n])
    thr = torch.topk(r, k=n_remove, largest=False, sorted=True).values[-1]
    idx = torch.nonzero(r > thr).flatten().to(torch.int32)
    keep = torch.zeros(n, dtype=torch.int32)
    keep[:idx.numel()] = idx
    return keep, int(idx.numel())


def consolidate_affinity(kp, s, e, proto, col_begin, aff, shr_out):
    k = kp[:, CK:].float()                                   # [N, CK]
    n = k.shape[0]
    for q in range(proto.numel()):
        pi = int(proto[q])
        if pi < col_begin:
            continue
        kq = k[pi]
        if e is not None:
            eq = e[pi].float()
            sim = -(((k - kq) ** 2) * eq).sum(1)
        else:
            sim = -(k * k).sum(1) + 2 * (k @ kq)
        sim = sim * s / math.sqrt(CK)
        a = torch.softmax(sim[col_begin:], dim=0)
        aff[q, :col_begin] = 0
        aff[q, col_begin:n] = a
        if shr_out is not None:
            shr_out[q] = (s[col_begin:] * a).sum()


def consolidate_values(gv, aff, col_begin, valid, n_valid):
    ng = gv.shape[2]
    rows = aff[:, col_begin:col_begin + ng] if valid is None else aff[valid.long(), col_begin:col_begin + ng]
    return (gv.float() @ rows[:n_valid].t()).half()


def bank_compact(kp, s, e, use, life, v, keep_idx, shift, first, m):
    src = keep_idx[first:m].long() if keep_idx is not None else torch.arange(first, m) + shift
    for t in (kp, s, e, use, life):
        t[first:m] = t[src].clone()
    v[:, :, first:m] = v[:, :, src].clone()


def install(monkeypatch, lib):
    monkeypatch.setattr(lib, 'require_cuda', lambda t, name: None)
    for name in ('key_pack', 'usage_topk', 'usage_evict_list', 'consolidate_affinity', 'consolidate_values', 'bank_compact'):
        monkeypatch.setattr(lib, name, globals()[name])
