"""Annotation-candidate selector — drop-in for the reference `inference/frame_selection/frame_selection.py:99-244`
(the second consumer of the anisotropic-L2 similarity, no softmax).

`select_next_candidates` greedily adds the frame whose SMALLEST cycle dissimilarity to the already chosen frames is the
largest; the dissimilarity of an ordered pair (A chosen, B candidate) is  mean over [HW, HW] of relu(S_ab - S_ba)
(reference :213-221).  Differences to the reference, none of them visible in the result:
  * the similarity matrices come from the tcgen05 scan kernel of the memory read (csrc/k1_affinity.cu, the
    `debug_scores` output of `xm_affinity_readout`: frame A's packed keys are the 'memory bank', frame B the query) instead
    of `get_similarity`'s five library ops; keys/selections are packed once per frame, not once per pair;
  * the reference recomputes every (chosen, candidate) pair in every round — O(rounds^2 * N) pairs; here each candidate
    keeps its running minimum, so a round costs N pairs;
  * runs on the device of `keys` (the reference hard-codes `cuda:0`).
Round-1 status: host logic verified on CPU against the live reference (tests/test_selector.py, with the pair score
replaced by the oracle); the kernel it calls is the validated read kernel, but this composition has not run on a GPU yet.
A fused pair kernel (both score tiles in TMEM, relu-difference reduced in the epilogue) is the planned replacement.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

from ... import lib

CK, CV = lib.CK, lib.CV


def first_frame_only(*args, **kwargs) -> List[int]:
    return [0]


def uniformly_selected_frames(dataloader, *args, how_many_frames=10, **kwargs) -> List[int]:
    return np.linspace(0, len(dataloader) - 1, how_many_frames).astype(int).tolist()


def _composite_keys(keys, masks, previously_chosen, alpha, min_mask_presence_percent, epsilon):
    """validity of every frame and its mask-weighted key (reference :156-189)."""
    n = len(keys)
    h, w = keys[0].shape[-2:]
    valid, comp = [], []
    for i in range(n):
        m = masks[i] if masks[i].ndim == 3 else masks[i].unsqueeze(0)
        m_bin = m.max(dim=0).values
        ratio = (m_bin > epsilon).sum() / m_bin.numel() * 100
        if ratio < min_mask_presence_percent and i not in previously_chosen:
            valid.append(False); comp.append(None)
            continue
        small = F.interpolate(m.unsqueeze(0).float(), size=(h, w), mode='nearest')[0].to(keys.device)
        any_object = small.max(dim=0, keepdim=True).values
        ck = keys[i] * any_object
        ck = ck * alpha + keys[i] * (1 - alpha)
        valid.append(True); comp.append(ck.to(keys[i].dtype))
    return valid, comp


class _PackedFrames:
    """Per frame: the packed memory-side rows (k^2, k), shrinkage, and the packed query-side rows (-e, 2ke) + b_sq."""

    def __init__(self, comp, shrinkages, selections, valid):
        dev = shrinkages.device
        self.h, self.w = shrinkages.shape[-2:]
        self.hw = self.h * self.w
        self.hw_pad = (self.hw + 127) // 128 * 128
        self.cap = (self.hw + 7) // 8 * 8 + 64
        self.frames = {}
        self.device = dev
        for i, ok in enumerate(valid):
            if not ok:
                continue
            krows = comp[i].permute(1, 2, 0).reshape(self.hw, CK).to(torch.float16).contiguous()
            erows = selections[i].permute(1, 2, 0).reshape(self.hw, CK).to(torch.float16).contiguous()
            kp = torch.zeros((self.cap, 2 * CK), dtype=torch.float16, device=dev)
            lib.key_pack(krows, kp[:self.hw])
            ms = torch.ones(self.cap, dtype=torch.float32, device=dev)
            ms[:self.hw] = shrinkages[i].reshape(-1).float()
            qp, bsq = lib.query_pack(krows, erows, self.hw_pad)
            self.frames[i] = (kp, ms, qp, bsq)


def _pair_scores(packed: _PackedFrames, chosen: int, candidates: List[int]) -> torch.Tensor:
    """cycle dissimilarity of (A = `chosen`, B = j) for every j in `candidates` (reference :213-221) -> fp32 [len]:
    csrc/pair_dissim.cu, one launch for all candidates; both 128x128 score tiles of a pair stay in TMEM and only the
    relu-difference sums leave the kernel."""
    L = lib.load()
    dev = packed.device
    if not hasattr(packed, 'stacked'):
        ids = sorted(packed.frames)
        slot = {f: i for i, f in enumerate(ids)}
        hw, hwp = packed.hw, packed.hw_pad
        kp = torch.zeros((len(ids), hwp, 2 * CK), dtype=torch.float16, device=dev)
        qp = torch.zeros((len(ids), hwp, 2 * CK), dtype=torch.float16, device=dev)
        bsq = torch.zeros((len(ids), hwp), dtype=torch.float32, device=dev)
        ms = torch.zeros((len(ids), hwp), dtype=torch.float32, device=dev)
        for f, i in slot.items():
            fkp, fms, fqp, fbsq = packed.frames[f]
            kp[i, :hw] = fkp[:hw]; ms[i, :hw] = fms[:hw]; qp[i] = fqp; bsq[i] = fbsq
            bsq[i, hw:] = 0
        packed.stacked = (slot, kp, qp, bsq, ms)
    slot, kp, qp, bsq, ms = packed.stacked
    tiles = packed.hw_pad // 128
    out = torch.empty(len(candidates), dtype=torch.float32, device=dev)
    for p0 in range(0, len(candidates), 65535):
        part = candidates[p0:p0 + 65535]
        pa = torch.full((len(part),), slot[chosen], dtype=torch.int32, device=dev)
        pb = torch.tensor([slot[j] for j in part], dtype=torch.int32, device=dev)
        partial = torch.empty((len(part), tiles * tiles), dtype=torch.float32, device=dev)
        rc = L.xm_pair_dissimilarity(kp.data_ptr(), qp.data_ptr(), bsq.data_ptr(), ms.data_ptr(), kp.shape[0], packed.hw, packed.hw_pad,
                                     pa.data_ptr(), pb.data_ptr(), len(part), partial.data_ptr(), out[p0:p0 + len(part)].data_ptr(),
                                     lib.stream_ptr())
        if rc != 0:
            raise RuntimeError(f'xm_pair_dissimilarity failed ({rc}): {L.xm_last_error().decode()}')
    return out


def select_next_candidates(keys: torch.Tensor, shrinkages, selections, masks: List[torch.Tensor], num_next_candidates: int,
                           previously_chosen_candidates: List[int] = (0,), print_progress=False, alpha=0.5,
                           min_mask_presence_percent=0.25, device=None, progress_callback=None, only_new_candidates=True,
                           epsilon=0.5) -> List[int]:
    """Same arguments as the reference (:99).  keys [N,CK,h,w], shrinkages [N,1,h,w], selections [N,CK,h,w] (what
    `extract_keys(..., flatten=False)` + `torch.cat` produce), masks: N tensors [C,H,W] or [H,W].  `device` defaults to
    the device of `keys`."""
    assert len(keys) == len(masks)
    assert len(keys) > 0
    assert num_next_candidates > 0
    assert len(previously_chosen_candidates) > 0
    assert 0.0 <= alpha <= 1.0
    assert min_mask_presence_percent >= 0
    assert len(previously_chosen_candidates) < len(keys)
    with torch.no_grad():
        if len(keys) > 1:
            keys = keys.squeeze()
        dev = torch.device(device) if device is not None else keys.device
        keys, shrinkages, selections = keys.to(dev), shrinkages.to(dev), selections.to(dev)
        lib.require_cuda(keys, 'keys')
        n = len(keys)
        chosen = list(previously_chosen_candidates)
        valid, comp = _composite_keys(keys, masks, chosen, alpha, min_mask_presence_percent, epsilon)
        print(f"Frames with invalid (empty or too small) masks: {valid.count(False)} / {len(masks)}")
        packed = _PackedFrames(comp, shrinkages, selections, valid)
        candidates = [j for j in range(n) if valid[j]]
        cand_t = torch.tensor(candidates, dtype=torch.long, device=dev)
        # running minimum of the dissimilarity to the chosen set; frames with an unusable mask stay at 0 (:200-202)
        dis_min = torch.zeros(n, dtype=torch.float32, device=dev)
        dis_min[cand_t] = float('inf')
        for c in chosen:
            dis_min[cand_t] = torch.minimum(dis_min[cand_t], _pair_scores(packed, c, candidates))
        for i in range(num_next_candidates):
            new = int(torch.argmax(dis_min))
            chosen.append(new)
            if i + 1 < num_next_candidates:
                dis_min[cand_t] = torch.minimum(dis_min[cand_t], _pair_scores(packed, new, candidates))
            if progress_callback is not None:
                progress_callback.emit(i + 1)
        if only_new_candidates:
            chosen = chosen[len(previously_chosen_candidates):]
        return chosen
