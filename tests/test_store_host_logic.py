"""CPU test of the arena bookkeeping of `KeyValueMemoryStore` (group suffix ranges, in-place sieve, least-used
eviction, right-aligned long-term groups) against the oracle's `torch.cat`-based store, which restates the reference
(inference/kv_memory_store.py:36-206).  The two CUDA entry points the store calls are replaced by CPU stand-ins with
the same contract (test doubles, not a product fallback), so only host logic is exercised."""
import pytest
import torch

from oracle import xmem_oracle as O
from xmem2_b200 import lib
from xmem2_b200.inference import kv_memory_store as kv

CK, CV = 64, 512


@pytest.fixture(autouse=True)
def cpu_doubles(monkeypatch):
    from tests import cpu_doubles
    cpu_doubles.install(monkeypatch, lib)
    monkeypatch.setattr(kv, '_ARENA_POOL', {})
    yield


def _frame(g, n, n_obj):
    key = (torch.randn(1, CK, n, generator=g) * 0.5).half()
    shr = torch.rand(1, 1, n, generator=g) + 1
    sel = torch.rand(1, CK, n, generator=g).half()
    val = torch.randn(n_obj, CV, n, generator=g).half()
    return key, val, shr, sel


def _same(store, ostore):
    assert store.size == ostore.size and store.num_groups == ostore.num_groups
    if store.size:
        assert torch.equal(store.k.float(), ostore.k.float())
        assert torch.allclose(store.s, ostore.s)
    for gi in range(store.num_groups):
        assert store.get_v_size(gi) == ostore.v[gi].shape[-1]
        assert torch.equal(store.v[gi].float(), ostore.v[gi].float()), gi


def test_append_groups_and_sieve_match_the_reference_layout():
    g = torch.Generator().manual_seed(0)
    store, ostore = kv.KeyValueMemoryStore(True), O.OracleStore(True)
    HW = 24
    # frames 0,1: object 1 only; frames 2..9: objects 1 and 2 -> second group owns a suffix of the columns
    for f in range(10):
        n_obj = 1 if f < 2 else 2
        key, val, shr, sel = _frame(g, HW, n_obj)
        objs = list(range(1, n_obj + 1))
        store.add(key, val, shr, sel, objs)
        ostore.add(key.float(), val.float(), shr, sel.float(), objs)
        _same(store, ostore)
    assert store.obj_groups == [[0], [1]] and store.group_begin(1) == 2 * HW
    # usage bookkeeping (kv_memory_store.py:96-103)
    u = torch.rand(store.size, generator=g)
    store.update_usage(u); ostore.update_usage(u)
    assert torch.allclose(store.get_usage(), ostore.usage())
    # compress_features' sieve: keep the last m columns; a group with fewer than m + HW columns keeps all of them
    m = 6 * HW
    store.sieve_by_range(0, -m, min_size=m + HW)
    ostore.k = ostore.k[:, :, -m:]; ostore.s = ostore.s[:, :, -m:]; ostore.e = ostore.e[:, :, -m:]
    ostore.use = ostore.use[:, :, -m:]; ostore.life = ostore.life[:, :, -m:]
    ostore.v = [v[:, :, -m:] if v.shape[-1] >= m + HW else v for v in ostore.v]
    _same(store, ostore)
    assert torch.allclose(store.get_usage(), ostore.usage())
    # appending after the compaction keeps working
    key, val, shr, sel = _frame(g, HW, 2)
    store.add(key, val, shr, sel, [1, 2]); ostore.add(key.float(), val.float(), shr, sel.float(), [1, 2])
    _same(store, ostore)


def test_zero_width_engage_and_remove_at():
    g = torch.Generator().manual_seed(1)
    store, ostore = kv.KeyValueMemoryStore(False), O.OracleStore(False)
    key, val, shr, sel = _frame(g, 16, 1)
    z = lambda t: t[..., 0:0]
    store.add(z(key), z(val), z(shr), z(sel), [1]); ostore.add(z(key).float(), z(val).float(), z(shr), z(sel).float(), [1])
    assert store.engaged() and store.size == 0 and store.num_groups == 1
    for _ in range(3):
        key, val, shr, sel = _frame(g, 16, 1)
        store.add(key, val, shr, sel, [1]); ostore.add(key.float(), val.float(), shr, sel.float(), [1])
    store.remove_at(16, 16)                       # drop the middle frame
    keep = torch.cat([torch.arange(0, 16), torch.arange(32, 48)])
    assert store.size == 32 and torch.equal(store.k.float(), ostore.k[:, :, keep].float())
    assert torch.equal(store.v[0].float(), ostore.v[0][:, :, keep].float())


def test_least_used_eviction_matches_reference():
    g = torch.Generator().manual_seed(2)
    store, ostore = kv.KeyValueMemoryStore(True), O.OracleStore(True)
    for _ in range(4):
        key, val, shr, _ = _frame(g, 32, 1)
        store.add(key, [val], shr, None, None); ostore.add(key.float(), [val.float()], shr, None, None)
    u = torch.rand(store.size, generator=g)
    store.update_usage(u); ostore.update_usage(u)
    store.remove_obsolete_features(100); ostore.remove_obsolete(100)
    _same(store, ostore)
    assert store.size <= 100


def test_long_term_group_with_fewer_prototypes_stays_right_aligned():
    # reference: a later group's long-term values are read against the LAST get_v_size(gi) keys (memory_manager.py:99-103)
    g = torch.Generator().manual_seed(3)
    store, ostore = kv.KeyValueMemoryStore(False), O.OracleStore(False)
    for n_valid in (16, 10, 16):
        key, v0, shr, _ = _frame(g, 16, 1)
        v1 = torch.randn(1, CV, n_valid, generator=g).half()
        store.add(key, [v0, v1], shr, None, None, group_objects=[[0], [1]])
        ostore.add(key.float(), [v0.float(), v1.float()], shr, None, None)
        _same(store, ostore)
    assert store.get_v_size(1) == 42 and store.group_begin(1) == 48 - 42


def test_arena_pool_is_bounded_by_bytes_and_can_be_cleared(monkeypatch):
    # round-1 advisor finding: the pool of recycled arenas must not grow without bound and must be droppable
    import torch
    from xmem2_b200.inference import kv_memory_store as K
    K.clear_arena_pool()
    dev = torch.device('cpu')
    s1 = K.KeyValueMemoryStore(count_usage=True)
    s1._alloc(256, 1, dev)
    one = K._arena_bytes((s1._kp, s1._s, s1._e, s1._v, s1._use, s1._life))
    monkeypatch.setattr(K, '_ARENA_POOL_LIMIT', int(one * 1.5))          # room for exactly one pooled arena of this size
    s2 = K.KeyValueMemoryStore(count_usage=True)
    s2._alloc(256, 1, dev)
    s1._release(); s2._release()
    assert K._ARENA_POOL_BYTES[0] == one and sum(len(v) for v in K._ARENA_POOL.values()) == 1
    s3 = K.KeyValueMemoryStore(count_usage=True)
    s3._alloc(256, 1, dev)                                                # takes the pooled arena back
    assert K._ARENA_POOL_BYTES[0] == 0
    s3._release()
    K.clear_arena_pool()
    assert K._ARENA_POOL_BYTES[0] == 0 and not K._ARENA_POOL
