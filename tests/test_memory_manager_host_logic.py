"""CPU test of MemoryManager's write path and consolidation (reference inference/memory_manager.py:212-390) on the
arena-backed stores, against the oracle's restatement: permanent preload that introduces a second object group, working
memory growth, usage-driven prototype selection, per-group full-softmax potentiation, least-used eviction of long-term
memory.  The store's two CUDA entry points are replaced by CPU doubles (see test_store_host_logic.py); consolidation
itself is plain torch ops and runs unchanged on CPU."""
import pytest
import torch

from oracle import xmem_oracle as O
from xmem2_b200 import lib
from xmem2_b200.inference import kv_memory_store as kv
from xmem2_b200.inference.memory_manager import MemoryManager

CK, CV = 64, 512
H, W = 3, 4
HW = H * W


@pytest.fixture(autouse=True)
def cpu_doubles(monkeypatch):
    from tests import cpu_doubles
    cpu_doubles.install(monkeypatch, lib)
    monkeypatch.setattr(kv, '_ARENA_POOL', {})
    yield


def _cfg(**over):
    cfg = dict(hidden_dim=64, top_k=30, enable_long_term=True, enable_long_term_count_usage=True, max_mid_term_frames=6,
               min_mid_term_frames=3, num_prototypes=8, max_long_term_elements=28, key_dim=64, value_dim=512)
    cfg.update(over)
    return cfg


def _frame(g, n_obj):
    key = (torch.randn(1, CK, H, W, generator=g) * 0.5).half()
    shr = torch.rand(1, 1, H, W, generator=g) + 1
    sel = torch.rand(1, CK, H, W, generator=g).half()
    val = torch.randn(1, n_obj, CV, H, W, generator=g).half()
    return key, shr, val, sel


def _check(mm, om, atol=2e-2):
    for mine, theirs in ((mm.temporary_work_mem, om.temp), (mm.permanent_work_mem, om.perm), (mm.long_mem, om.long)):
        assert mine.size == theirs.size and mine.num_groups == theirs.num_groups
        if mine.size:
            assert torch.allclose(mine.k.float(), theirs.k.float(), atol=2e-3)
            assert torch.allclose(mine.s, theirs.s, atol=2e-3)
        for gi in range(mine.num_groups):
            assert mine.get_v_size(gi) == theirs.v[gi].shape[-1], gi
            assert torch.allclose(mine.v[gi].float(), theirs.v[gi].float(), atol=atol), gi


def test_write_path_consolidation_and_eviction_follow_the_reference():
    g = torch.Generator().manual_seed(0)
    cfg = _cfg(max_long_term_elements=400)      # eviction with two groups raises in the reference (kv_memory_store.py:171-176)
    mm, om = MemoryManager(dict(cfg)), O.OracleMemory(dict(cfg))
    # permanent preload: first annotated frame has one object, the second introduces object 2 -> two value groups
    for n_obj in (1, 2):
        key, shr, val, sel = _frame(g, n_obj)
        objs = list(range(1, n_obj + 1))
        mm.add_memory(key, shr, val, objs, selection=sel, permanent=True)
        om.add(key.float(), shr, val.float(), objs, selection=sel.float(), permanent=True)
    assert mm.permanent_work_mem.obj_groups == [[0], [1]] and mm.temporary_work_mem.num_groups == 2
    _check(mm, om)
    consolidations = 0
    for step in range(20):
        key, shr, val, sel = _frame(g, 2)
        before = mm.long_mem.size
        # identical usage statistics on both sides (the fused kernel normally accumulates them)
        if mm.temporary_work_mem.size:
            u = torch.rand(mm.temporary_work_mem.size, generator=g)
            mm.temporary_work_mem.update_usage(u); om.temp.update_usage(u)
        if mm.long_mem.size:
            u = torch.rand(mm.long_mem.size, generator=g)
            mm.long_mem.update_usage(u); om.long.update_usage(u)
        mm.add_memory(key, shr, val, [1, 2], selection=sel)
        om.add(key.float(), shr, val.float(), [1, 2], selection=sel.float())
        consolidations += mm.long_mem.size != before
        _check(mm, om)
    assert consolidations >= 4                      # 20 frames, compress every (6-3) memory frames once full
    assert mm.long_mem.size <= cfg['max_long_term_elements']
    assert mm.temporary_work_mem.size == om.temp.size <= cfg['max_mid_term_frames'] * HW
    # the kernel-side description of the banks agrees with the reference's suffix slicing
    a, n_obj, use_long = mm._read_args()
    assert n_obj == 2 and use_long and a.n_groups == 2
    for gi in range(2):
        assert a.groups[gi].begin[1] == mm.temporary_work_mem.size - om.temp.v[gi].shape[-1]
        assert a.groups[gi].begin[2] == mm.permanent_work_mem.size - om.perm.v[gi].shape[-1]


def test_eviction_with_two_groups_raises_like_the_reference():
    g = torch.Generator().manual_seed(2)
    cfg = _cfg(max_long_term_elements=12)
    mm = MemoryManager(dict(cfg))
    for n_obj in (1, 2):
        key, shr, val, sel = _frame(g, n_obj)
        mm.add_memory(key, shr, val, list(range(1, n_obj + 1)), selection=sel, permanent=True)
    with pytest.raises(NotImplementedError):
        for step in range(16):
            key, shr, val, sel = _frame(g, 2)
            if mm.temporary_work_mem.size:
                mm.temporary_work_mem.update_usage(torch.rand(mm.temporary_work_mem.size, generator=g))
            if mm.long_mem.size:
                mm.long_mem.update_usage(torch.rand(mm.long_mem.size, generator=g))
            mm.add_memory(key, shr, val, [1, 2], selection=sel)


def test_single_group_eviction_keeps_most_used_prototypes():
    g = torch.Generator().manual_seed(1)
    cfg = _cfg(max_long_term_elements=20)
    mm, om = MemoryManager(dict(cfg)), O.OracleMemory(dict(cfg))
    key, shr, val, sel = _frame(g, 1)
    mm.add_memory(key, shr, val, [1], selection=sel, permanent=True)
    om.add(key.float(), shr, val.float(), [1], selection=sel.float(), permanent=True)
    for step in range(16):
        key, shr, val, sel = _frame(g, 1)
        if mm.temporary_work_mem.size:
            u = torch.rand(mm.temporary_work_mem.size, generator=g)
            mm.temporary_work_mem.update_usage(u); om.temp.update_usage(u)
        if mm.long_mem.size:
            u = torch.rand(mm.long_mem.size, generator=g)
            mm.long_mem.update_usage(u); om.long.update_usage(u)
        mm.add_memory(key, shr, val, [1], selection=sel)
        om.add(key.float(), shr, val.float(), [1], selection=sel.float())
        _check(mm, om)
    assert 0 < mm.long_mem.size <= 20
