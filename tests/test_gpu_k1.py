"""GPU parity of the fused similarity -> top-k softmax -> readout(+usage) kernel against the oracle math
(reference: model/memory_util.py:7-65, inference/memory_manager.py:61-190).  Calls go through the C ABI."""
import pytest
import torch

from tests import k1_ref

pytestmark = pytest.mark.gpu

# tolerance: P is rounded to fp16 (as the reference does under autocast, memory_manager.py:59) and the
# readout is accumulated in fp32 then stored fp16 -> 2e-2 absolute on values ~N(0,1) sums.
TOL_OUT = 2e-2
TOL_SCORE = 2e-4
TOL_USAGE = 2e-3


def _check(case, top_k=30, max_ambiguous=None):
    r = k1_ref.compare(case, top_k)
    assert r['score_err'] < TOL_SCORE, r
    assert r['layout_err'] == 0.0, r
    assert r['out_err_clear'] < TOL_OUT, r
    assert r['ambiguous'] <= (max_ambiguous if max_ambiguous is not None else max(3, case['hw'] // 50)), r
    assert r['usage_err'] < TOL_USAGE, r
    assert abs(r['usage_sum'] - case['hw']) < 0.05 * case['hw'] + 1, r    # affinity columns sum to 1
    return r


def test_single_bank_small():
    _check(k1_ref.make_case(hw=96, sizes=(0, 0, 700), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=1))


def test_three_banks_ragged_sizes():
    # sizes not multiples of the 64-column tile; hw not a multiple of 128
    _check(k1_ref.make_case(hw=200, sizes=(130, 333, 517), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=2))


def test_two_groups_suffix_ranges():
    # group 1 (object 1) only sees a suffix of working/permanent memory and no long-term memory
    case = k1_ref.make_case(hw=150, sizes=(128, 300, 405), n_obj=2,
                            group_begins=[(0, 1, [0, 0, 0]), (1, 1, [128, 150, 135])], seed=3)
    _check(case)


def test_minimum_columns_equals_topk():
    _check(k1_ref.make_case(hw=64, sizes=(0, 0, 30), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=4))


def test_fewer_columns_than_topk_raises():
    case = k1_ref.make_case(hw=64, sizes=(0, 0, 20), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=5)
    with pytest.raises(RuntimeError, match='top_k'):
        k1_ref.run_kernel(case)


def test_two_objects_one_group():
    _check(k1_ref.make_case(hw=128, sizes=(0, 256, 256), n_obj=2, group_begins=[(0, 2, [0, 0, 0])], seed=6))


def test_config2_shape_480p():
    # BASELINE.json config 2 at its largest memory: HW=1620, 5 permanent + 9 working frames, 1 object
    case = k1_ref.make_case(hw=1620, sizes=(0, 9 * 1620, 5 * 1620), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=7)
    _check(case)


def test_planted_cluster_overflows_candidate_lists():
    # 50 memory columns nearly identical to every query sit in two consecutive tiles of ONE column slice, so a single
    # per-slice candidate list receives more than 32 entries and the in-kernel compaction path must stay exact.
    import torch
    case = k1_ref.make_case(hw=1620, sizes=(0, 4096, 0), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=8)
    g = torch.Generator().manual_seed(80)
    q0 = torch.randn(64, generator=g) * 0.6
    case['qk'] = (q0[None, :] + 0.02 * torch.randn(1620, 64, generator=g)).half()
    planted = list(range(128, 160)) + list(range(192, 210))
    case['banks'][1]['key'][planted] = (q0[None, :] + 0.05 * torch.randn(len(planted), 64, generator=g)).half()
    _check(case, max_ambiguous=400)     # near-identical planted columns produce near-ties at rank 30 by construction


def test_full_size_1080p_partition_of_unity():
    # BASELINE.json config-4 scale on one GPU (HW = 8160 queries, 5 working frames + 1 permanent = 48 960 columns):
    # size-independent properties instead of an oracle run — with all values == 1 the readout must be exactly the sum of
    # the affinity weights (== 1 for every query and channel), and the usage column sums must add up to the number of
    # queries (every affinity column of do_softmax sums to 1, model/memory_util.py:49,63).
    import torch
    hw = 8160
    case = k1_ref.make_case(hw=hw, sizes=(0, 5 * hw, hw), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=9)
    for b in case['banks']:
        if b is not None:
            b['val'][:, :, :b['n']] = 1.0
    out_chw, out_hwc, usage_bufs, _ = k1_ref.run_kernel(case)
    assert torch.isfinite(out_chw).all()
    assert (out_chw.float() - 1.0).abs().max().item() < 4e-3          # 30 fp16-rounded weights summed in fp32
    assert torch.equal(out_hwc.transpose(1, 2), out_chw)
    total = sum(float(u.sum()) for u in usage_bufs if u is not None)
    assert abs(total - hw) < 1e-2 * hw


def test_reference_golden_vectors_through_the_c_abi():
    # tests/golden/attention.npz: inputs AND outputs of the live reference's get_similarity / do_softmax(top_k=30) / readout.
    # (a) tight: against the oracle functions on the fp16-rounded operands; (b) against the reference's own fp32 outputs,
    # with the bound that fp16 rounding of keys / selection / values implies (asserted for the oracle on CPU too,
    # tests/test_k1_ref_vs_oracle.py).
    import os
    import numpy as np
    case = k1_ref.case_from_attention_golden()
    r = _check(case)
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'attention.npz'))
    out_chw, out_hwc, usage_bufs, dbg = k1_ref.run_kernel(case, 30, want_debug=True)
    hw = case['hw']
    assert (dbg[:, :hw].float().cpu() - torch.from_numpy(d['sim'])[0]).abs().max().item() < 6e-2
    _, _, amb, _ = k1_ref.expected(case, 30)
    got = out_chw.float().cpu()
    assert (got[:, :32][:, :, ~amb] - torch.from_numpy(d['readout'])[:, :, ~amb]).abs().max().item() < 8e-2
    assert got[:, 32:].abs().max().item() == 0.0
    u = usage_bufs[2][:case['banks'][2]['n']].float().cpu()
    du = (u - torch.from_numpy(d['usage'])[0]).abs()
    assert du.mean().item() < 2e-3 and abs(u.sum().item() - hw) < 1e-2 * hw, (du.mean().item(), u.sum().item())
