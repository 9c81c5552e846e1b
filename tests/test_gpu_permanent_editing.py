"""Permanent-memory editing API on the GPU (reference inference/memory_manager.py:192-210,392-425,
inference/inference_core.py:40-48,154-186): update / remove / clear_memory(keep_permanent=True).  The CPU test
(tests/test_permanent_memory_editing.py) pins the bookkeeping to the live reference; here the same calls run on the device
arenas and kernels and must be CONSISTENT: an edited memory segments exactly like a memory that was built that way."""
import pytest
import torch

from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict, synth_frame, synth_mask

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
CFG = dict(mem_every=3, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True, hidden_dim=64,
           key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
           max_long_term_elements=10000)
H, W = 96, 128
dev = 'cuda'


@pytest.fixture(scope='module')
def net():
    n = XMem(dict(CFG), None).to(dev).eval()
    n.load_weights(synth_state_dict(0))
    return n


def _f(ti):
    return synth_frame(ti, H, W, structured=True).to(dev)


def _m(ti):
    return synth_mask(ti, H, W, 1).to(dev)


def _probe(core, start=20, n=3):
    # disable_memory_updates: the probe frames do not change the state they probe
    return [core.step(_f(start + i), disable_memory_updates=True).clone() for i in range(n)]


def test_update_permanent_memory_equals_building_it_that_way(net):
    a = InferenceCore(net, dict(CFG)); a.set_all_labels([1])
    assert a.put_to_permanent_memory(_f(0), _m(0), ti=0) is False
    assert a.put_to_permanent_memory(_f(5), _m(5), ti=5) is False
    assert a.put_to_permanent_memory(_f(0), _m(7), ti=0) is True          # frame 0 re-annotated with another mask
    b = InferenceCore(net, dict(CFG)); b.set_all_labels([1])
    b.put_to_permanent_memory(_f(0), _m(7), ti=0)
    b.put_to_permanent_memory(_f(5), _m(5), ti=5)
    assert a.permanent_memory_frames == [0, 5] and a.memory.permanent_work_mem.size == b.memory.permanent_work_mem.size
    assert torch.equal(a.memory.permanent_work_mem.key, b.memory.permanent_work_mem.key)
    assert torch.equal(a.memory.permanent_work_mem.value[0], b.memory.permanent_work_mem.value[0])
    for x, y in zip(_probe(a), _probe(b)):
        assert torch.equal(x, y)


def test_clear_memory_keep_permanent_restarts_from_the_permanent_frames(net):
    a = InferenceCore(net, dict(CFG)); a.set_all_labels([1])
    a.put_to_permanent_memory(_f(0), _m(0), ti=0)
    a.step(_f(0), _m(0), [1], do_not_add_mask_to_memory=True)
    for ti in range(1, 9):
        a.step(_f(ti))
    assert a.memory.temporary_work_mem.size > 0
    a.clear_memory(keep_permanent=True)
    assert a.memory.temporary_work_mem.size == 0 and a.memory.permanent_work_mem.size == (H // 16) * (W // 16)
    assert a.curr_ti == -1 and a.permanent_memory_frames == [0]
    b = InferenceCore(net, dict(CFG)); b.set_all_labels([1])
    b.put_to_permanent_memory(_f(0), _m(0), ti=0)
    # same frames from the same starting state -> same probabilities (hidden state starts from zeros in both)
    for ti in range(4):
        m = _m(0) if ti == 0 else None
        pa = a.step(_f(ti), m, [1] if m is not None else None, do_not_add_mask_to_memory=m is not None)
        pb = b.step(_f(ti), m, [1] if m is not None else None, do_not_add_mask_to_memory=m is not None)
        assert torch.equal(pa, pb), ti


def test_remove_from_permanent_memory_follows_the_reference_indexing(net):
    # reference quirk: remove_at(pos, HW) removes elements [pos, pos + HW) with the FRAME position used as
    # an element offset -> removing frame position 0 is exact; positions > 0 cut across frames.  Reproduced, not fixed.
    hw = (H // 16) * (W // 16)
    a = InferenceCore(net, dict(CFG)); a.set_all_labels([1])
    a.put_to_permanent_memory(_f(0), _m(0), ti=0)
    a.put_to_permanent_memory(_f(5), _m(5), ti=5)
    keys_before = a.memory.permanent_work_mem.key.clone()
    a.remove_from_permanent_memory(0)
    assert a.permanent_memory_frames == [5] and a.memory.permanent_work_mem.size == hw
    assert torch.equal(a.memory.permanent_work_mem.key, keys_before[:, :, hw:])
    b = InferenceCore(net, dict(CFG)); b.set_all_labels([1])
    b.put_to_permanent_memory(_f(5), _m(5), ti=5)
    for x, y in zip(_probe(a), _probe(b)):
        assert torch.equal(x, y)
