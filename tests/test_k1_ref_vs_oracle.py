"""CPU pin of the K1 test helper (tests/k1_ref.py) to the oracle: the expected values the GPU parity tests compare
against are the oracle's own functions (oracle.similarity / oracle.softmax_topk, themselves pinned to the reference by
tests/test_oracle_golden.py) evaluated on the operands the kernel consumes."""
import os

import numpy as np
import torch

from oracle import xmem_oracle as O
from tests import k1_ref

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_kernel_operand_similarity_is_the_oracle_similarity_up_to_fp16_operand_rounding():
    case = k1_ref.make_case(hw=96, sizes=(40, 333, 200), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=21, device='cpu')
    a, b = k1_ref.kernel_operand_similarity(case), k1_ref.oracle_similarity(case)
    for sa, sb in zip(a, b):
        # k^2 and 2ke rounded to fp16 (rel. 2^-11 each) inside 64-term sums of O(1) magnitudes, times shrinkage <= 3, / 8
        assert (sa - sb).abs().max().item() < 5e-3
        assert (sa - sb).abs().mean().item() < 5e-4


def test_expected_uses_oracle_softmax_on_the_golden_vectors():
    # the reference-generated attention fixture, fed through k1_ref.expected in the kernel's layout
    case = k1_ref.case_from_attention_golden(device='cpu')
    out, usage, amb, scores = k1_ref.expected(case, 30)
    d = np.load(os.path.join(G, 'attention.npz'))
    # fp16 rounding of keys / selection / values against the fp32 reference outputs
    assert (scores.float() - torch.from_numpy(d['sim'])[0]).abs().max().item() < 6e-2
    clear = ~amb
    ref_ro = torch.from_numpy(d['readout']).double()
    assert (out[:, :32][:, :, clear] - ref_ro[:, :, clear]).abs().max().item() < 6e-2
    assert out[:, 32:].abs().max().item() == 0.0
