"""Permanent-memory editing API (GUI-facing) checked against the LIVE reference MemoryManager on CPU
(build container only): add with a frame id, update in place, remove (including the reference's position-as-offset quirk,
memory_manager.py:204-210), and `copy_perm_mem_only` used by `InferenceCore.clear_memory(keep_permanent=True)`.
The arena store's CUDA entry points are replaced by CPU doubles; only host logic is exercised."""
import os
import sys
import types

import pytest
import torch

from xmem2_b200 import lib
from xmem2_b200.inference import kv_memory_store as kv
from xmem2_b200.inference.memory_manager import MemoryManager

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present on this box')
CK, CV, H, W = 64, 512, 3, 4
HW = H * W


@pytest.fixture
def ref_manager_cls(monkeypatch):
    from tests import cpu_doubles
    cpu_doubles.install(monkeypatch, lib)
    monkeypatch.setattr(kv, '_ARENA_POOL', {})
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('inference', 'model', 'util')}
    for k in saved:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(REF)
    import inference.memory_manager as ref_mm
    yield ref_mm.MemoryManager
    for k in [k for k in sys.modules if k.split('.')[0] in ('inference', 'model', 'util')]:
        sys.modules.pop(k, None)
    sys.modules.update(saved)


def _cfg():
    return dict(hidden_dim=64, top_k=30, enable_long_term=True, enable_long_term_count_usage=True, max_mid_term_frames=10,
                min_mid_term_frames=5, num_prototypes=128, max_long_term_elements=10000, key_dim=64, value_dim=512)


def _frame(g):
    key = (torch.randn(1, CK, H, W, generator=g) * 0.5).half()
    shr = torch.rand(1, 1, H, W, generator=g) + 1
    sel = torch.rand(1, CK, H, W, generator=g).half()
    val = torch.randn(1, 1, CV, H, W, generator=g).half()
    return key, shr, val, sel


def _same_perm(mm, rm):
    a, b = mm.permanent_work_mem, rm.permanent_work_mem
    assert a.size == b.size
    if a.size:
        assert torch.equal(a.key.float(), b.key.float()) and torch.allclose(a.shrinkage, b.shrinkage)
        assert torch.equal(a.value[0].float(), b.value[0].float())
    assert mm.frame_id_to_permanent_mem_idx == rm.frame_id_to_permanent_mem_idx


def test_add_update_remove_follow_the_reference(ref_manager_cls):
    g = torch.Generator().manual_seed(0)
    mm, rm = MemoryManager(_cfg()), ref_manager_cls(_cfg())
    frames = {}
    for ti in (0, 7, 15):
        key, shr, val, sel = _frame(g)
        frames[ti] = (key, shr, val, sel)
        mm.add_memory(key, shr, val, [1], selection=sel, permanent=True, ti=ti)
        rm.add_memory(key.float(), shr, val.float(), [1], selection=sel.float(), permanent=True, ti=ti)
    assert mm.frame_already_saved(7) and not mm.frame_already_saved(8)
    _same_perm(mm, rm)
    # in-place update of frame 7 (put_to_permanent_memory on an already saved frame, inference_core.py:170-172)
    key, shr, val, sel = _frame(g)
    mm.update_permanent_memory(7, key, shr, val, selection=sel)
    rm.update_permanent_memory(7, key.float(), shr, val.float(), selection=sel.float())
    _same_perm(mm, rm)
    # removal: the reference passes the frame POSITION as an element offset; same here
    mm.remove_from_permanent_memory(7)
    rm.remove_from_permanent_memory(7)
    _same_perm(mm, rm)
    assert mm.permanent_work_mem.size == 2 * HW


def test_copy_perm_mem_only_keeps_permanent_and_resets_the_rest(ref_manager_cls):
    g = torch.Generator().manual_seed(1)
    mm, rm = MemoryManager(_cfg()), ref_manager_cls(_cfg())
    for ti in (0, 3):
        key, shr, val, sel = _frame(g)
        mm.add_memory(key, shr, val, [1], selection=sel, permanent=True, ti=ti)
        rm.add_memory(key.float(), shr, val.float(), [1], selection=sel.float(), permanent=True, ti=ti)
    key, shr, val, sel = _frame(g)
    mm.add_memory(key, shr, val, [1], selection=sel)
    rm.add_memory(key.float(), shr, val.float(), [1], selection=sel.float())
    mm.create_hidden_state(1, key); rm.create_hidden_state(1, key.float())
    nm, nr = mm.copy_perm_mem_only(), rm.copy_perm_mem_only()
    _same_perm(nm, nr)
    assert nm.temporary_work_mem.size == nr.temporary_work_mem.size == 0
    assert nm.temporary_work_mem.num_groups == nr.temporary_work_mem.num_groups == 1
    assert nm.get_hidden().shape == nr.get_hidden().shape and float(nm.get_hidden().abs().sum()) == 0.0
    assert (nm.H, nm.W, nm.HW, nm.CK, nm.CV) == (nr.H, nr.W, nr.HW, nr.CK, nr.CV)
    # the copy keeps working as a memory: a new working frame can be added
    key, shr, val, sel = _frame(g)
    nm.add_memory(key, shr, val, [1], selection=sel)
    nr.add_memory(key.float(), shr, val.float(), [1], selection=sel.float())
    assert nm.temporary_work_mem.size == nr.temporary_work_mem.size == HW
