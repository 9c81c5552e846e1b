"""Library-op versions of the attention math for the LOW-FREQUENCY callers only (consolidation every
~50 frames, reference inference/memory_manager.py:349-390, and the annotation-candidate selector).
Same signatures as the reference `model/memory_util.py:7-65`.  The per-frame read does NOT use these: it is
the fused kernel behind `MemoryManager.match_memory` (csrc/k1_affinity.cu)."""
import math
from typing import Optional

import torch


def get_similarity(mk, ms, qk, qe):
    """mk [B,CK,N], ms [B,1,N]|None, qk [B,CK,Q], qe [B,CK,Q]|None -> [B,N,Q] (memory_util.py:7-39)."""
    ck = mk.shape[1]
    mk = mk.flatten(start_dim=2)
    qk = qk.flatten(start_dim=2)
    mkt = mk.transpose(1, 2)
    if qe is not None:
        qe = qe.flatten(start_dim=2)
        sim = -(mkt.pow(2) @ qe) + 2 * (mkt @ (qk * qe)) - (qe * qk.pow(2)).sum(1, keepdim=True)
    else:
        sim = -mk.pow(2).sum(1).unsqueeze(2) + 2 * (mkt @ qk)
    if ms is not None:
        sim = sim * ms.flatten(start_dim=1).unsqueeze(2)
    return sim / math.sqrt(ck)


def do_softmax(similarity, top_k: Optional[int] = None, inplace=False, return_usage=False):
    """memory_util.py:41-65 (top-k branch without max subtraction; full branch with)."""
    if top_k is not None:
        values, indices = torch.topk(similarity, k=top_k, dim=1)
        x_exp = values.exp()
        x_exp = x_exp / x_exp.sum(dim=1, keepdim=True)
        target = similarity.zero_() if inplace else torch.zeros_like(similarity)
        affinity = target.scatter_(1, indices, x_exp)
    else:
        maxes = similarity.max(dim=1, keepdim=True)[0]
        x_exp = (similarity - maxes).exp()
        affinity = x_exp / x_exp.sum(dim=1, keepdim=True)
    if return_usage:
        return affinity, affinity.sum(dim=2)
    return affinity


def get_affinity(mk, ms, qk, qe):
    return do_softmax(get_similarity(mk, ms, qk, qe))


def readout(affinity, mv):
    B, CV, T, H, W = mv.shape
    return torch.bmm(mv.view(B, CV, T * H * W), affinity).view(B, CV, H, W)
