"""Annotation-candidate selector on the GPU: `select_next_candidates` with its pair scores computed by
the fused pair kernel (csrc/pair_dissim.cu), against the picks and pair scores of the LIVE reference
(tests/golden/selector.npz) and against the round-1 composition of two similarity dumps per pair."""
import os

import numpy as np
import pytest
import torch

from tests.selector_case import selector_inputs, SELECTOR_CASES
from xmem2_b200.inference.frame_selection import frame_selection as fs

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden', 'selector.npz')


def test_pair_scores_match_the_reference():
    d = np.load(G)
    keys, shr, sel, masks = selector_inputs()
    dev = 'cuda'
    keys, shr, sel = keys.to(dev), shr.to(dev), sel.to(dev)
    valid, comp = fs._composite_keys(keys, masks, [0], 0.5, 0.25, 0.5)
    assert valid == d['valid'].tolist()
    packed = fs._PackedFrames(comp, shr, sel, valid)
    cands = [j for j in range(len(keys)) if valid[j]]
    worst = 0.0
    for a in cands:
        got = fs._pair_scores(packed, a, cands).cpu()
        want = torch.from_numpy(d['scores'][a, cands])
        worst = max(worst, ((got - want).abs() / (want.abs() + 1e-2)).max().item())
        assert abs(got[cands.index(a)].item()) < 1e-6       # D(A -> A) = 0 (two MMA streams: not bit-identical tiles)
    # fp16 operands (keys and selections are fp16-exact in the fixture; k^2 and 2ke are rounded to fp16) vs the fp32 reference
    assert worst < 2e-2, worst


def test_picks_match_the_reference():
    d = np.load(G)
    keys, shr, sel, masks = selector_inputs()
    keys, shr, sel = keys.cuda(), shr.cuda(), sel.cuda()
    for c, want in zip(SELECTOR_CASES, d['picks']):
        want = [int(x) for x in want if x >= 0]
        got = fs.select_next_candidates(keys, shr, sel, masks, c['k'], previously_chosen_candidates=list(c['prev']), alpha=c['alpha'])
        assert got == want, (c, got, want)


def test_fused_pair_kernel_equals_the_score_dump_composition():
    from tests.selector_case import pair_scores_via_score_dump
    keys, shr, sel, masks = selector_inputs()
    keys, shr, sel = keys.cuda(), shr.cuda(), sel.cuda()
    valid, comp = fs._composite_keys(keys, masks, [0], 0.5, 0.25, 0.5)
    packed = fs._PackedFrames(comp, shr, sel, valid)
    cands = [j for j in range(len(keys)) if valid[j]]
    for a in cands[:4]:
        plain = pair_scores_via_score_dump(packed, a, cands)
        fused = fs._pair_scores(packed, a, cands)
        torch.cuda.synchronize()
        assert torch.allclose(plain, fused, rtol=1e-4, atol=1e-6), (a, plain, fused)     # same operands, same fp32 expression
        assert abs(fused[cands.index(a)].item()) < 1e-6        # the two score tiles come from different MMA streams
