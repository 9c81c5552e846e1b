"""world_size-2 gloo test (CPU) of the T-sharded WRITE path: frame ownership in `add_memory`,
the candidate gather of the consolidation and the globally agreed long-term eviction.  Both ranks feed identical frames
to a `ShardedMemoryManager`; rank 0 also drives the single-process `MemoryManager`.  After every step the union of the
two shards, put back into global column order, must equal the single-process banks (working, permanent, long-term:
keys, shrinkage, selection, values, usage).  The stores' CUDA entry points are replaced by the CPU doubles of
test_store_host_logic.py."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

CK, CV = 64, 512
H, W = 3, 4
HW = H * W


def _install_cpu_doubles():
    from xmem2_b200 import lib
    from xmem2_b200.inference import kv_memory_store as kv

    from tests import cpu_doubles
    lib.require_cuda = lambda t, name: None
    for name in ('key_pack', 'usage_topk', 'usage_evict_list', 'consolidate_affinity', 'consolidate_values', 'bank_compact'):
        setattr(lib, name, getattr(cpu_doubles, name))
    kv._ARENA_POOL.clear()


def _cfg():
    return dict(hidden_dim=64, top_k=30, enable_long_term=True, enable_long_term_count_usage=True, max_mid_term_frames=6,
                min_mid_term_frames=3, num_prototypes=8, max_long_term_elements=28, key_dim=64, value_dim=512)


def _frame(g, n_obj):
    key = (torch.randn(1, CK, H, W, generator=g) * 0.5).half()
    shr = torch.rand(1, 1, H, W, generator=g) + 1
    sel = torch.rand(1, CK, H, W, generator=g).half()
    val = torch.randn(1, n_obj, CV, H, W, generator=g).half()
    return key, shr, val, sel


def _bank_tensors(store):
    if store.size == 0:
        return None
    return dict(k=store.k.float().clone(), s=store.s.clone(), e=store.e.float().clone() if store.e is not None else None,
                v=store.v[0].float().clone(), use=store.use_count.clone() if store.use_count is not None else None)


def _worker(rank, world, port, out):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        _install_cpu_doubles()
        from xmem2_b200.inference.memory_manager import MemoryManager
        from xmem2_b200.inference.sharded_memory import ShardedMemoryManager
        g = torch.Generator().manual_seed(0)
        sm = ShardedMemoryManager(_cfg())
        ref = MemoryManager(_cfg())          # every rank keeps the single-process manager: it also supplies the global sizes
        stats = dict(consolidations=0, evictions=0, checks=0)

        def compare(step):
            for name, mine, theirs in (('temp', sm.temporary_work_mem, ref.temporary_work_mem),
                                       ('perm', sm.permanent_work_mem, ref.permanent_work_mem), ('long', sm.long_mem, ref.long_mem)):
                pos = sm.local_positions(name)
                assert pos.numel() == mine.size, (step, name, pos.numel(), mine.size)
                sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
                dist.all_gather(sizes, torch.tensor([mine.size]))
                assert sum(int(s) for s in sizes) == theirs.size, (step, name, sizes, theirs.size)
                assert mine.num_groups == theirs.num_groups, (step, name)
                if mine.size == 0:
                    continue
                # this rank's columns must be exactly the single-process columns at `pos` (bit-exact: same inputs, same ops)
                t, r = _bank_tensors(mine), _bank_tensors(theirs)
                for field in ('k', 's', 'e', 'v', 'use'):
                    if t[field] is None:
                        assert r[field] is None
                        continue
                    want = r[field][..., pos]
                    assert torch.equal(t[field], want), (step, name, field, (t[field] - want).abs().max().item())
                stats['checks'] += 1

        # permanent preload: three annotated frames of one object
        for ti in range(3):
            key, shr, val, sel = _frame(g, 1)
            sm.add_memory(key, shr, val, [1], selection=sel, permanent=True, ti=ti)
            ref.add_memory(key, shr, val, [1], selection=sel, permanent=True, ti=ti)
            compare(('perm', ti))
        assert all(sm.frame_already_saved(ti) for ti in range(3))
        assert sm.frame_id_to_permanent_mem_idx == ref.frame_id_to_permanent_mem_idx

        # an edited annotation replaces the same global block as in the single-process bank, on its owner only
        key, shr, val, sel = _frame(g, 1)
        sm.update_permanent_memory(1, key, shr, val, selection=sel)
        ref.update_permanent_memory(1, key, shr, val, selection=sel)
        compare(('edit', 1))

        for step in range(40):
            key, shr, val, sel = _frame(g, 1)
            # identical usage statistics everywhere (the read kernel accumulates them per owned column)
            for name, mine, theirs in (('temp', sm.temporary_work_mem, ref.temporary_work_mem), ('long', sm.long_mem, ref.long_mem)):
                if theirs.size:
                    u = torch.rand(theirs.size, generator=g)
                    theirs.update_usage(u)
                    mine.update_usage(u[sm.local_positions(name)])
            long_before, blocks_before = ref.long_mem.size, sm._blocks['long']
            sm.add_memory(key, shr, val, [1], selection=sel)
            ref.add_memory(key, shr, val, [1], selection=sel)
            if sm._blocks['long'] > blocks_before:
                stats['consolidations'] += 1
                if ref.long_mem.size < long_before + 8:
                    stats['evictions'] += 1
            assert sm.global_temp_size == ref.temporary_work_mem.size and sm.global_long_size == ref.long_mem.size
            compare(step)
        assert stats['consolidations'] >= 8 and stats['evictions'] >= 3, stats
        out.put((rank, 'ok', stats, sm.long_mem.size, sm.temporary_work_mem.size))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as exc:                                   # surface the failure in the parent
        import traceback
        out.put((rank, 'error', traceback.format_exc(), 0, 0))
        raise exc


@pytest.mark.parametrize('world', [2, 3])
def test_shards_hold_exactly_the_single_process_memory(world):
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank, status, info, n_long, n_temp in res:
        assert status == 'ok', info
    # every shard took part: prototype blocks rotate over the ranks, working frames too
    assert all(r[3] > 0 and r[4] > 0 for r in res), res


def test_single_process_sharded_manager_equals_the_plain_one():
    _install_cpu_doubles()
    from xmem2_b200.inference.memory_manager import MemoryManager
    from xmem2_b200.inference.sharded_memory import ShardedMemoryManager
    g = torch.Generator().manual_seed(1)
    sm, ref = ShardedMemoryManager(_cfg()), MemoryManager(_cfg())
    assert sm.world == 1
    for step in range(15):
        key, shr, val, sel = _frame(g, 1)
        if ref.temporary_work_mem.size:
            u = torch.rand(ref.temporary_work_mem.size, generator=g)
            ref.temporary_work_mem.update_usage(u); sm.temporary_work_mem.update_usage(u)
        sm.add_memory(key, shr, val, [1], selection=sel)
        ref.add_memory(key, shr, val, [1], selection=sel)
        for mine, theirs in ((sm.temporary_work_mem, ref.temporary_work_mem), (sm.long_mem, ref.long_mem)):
            assert mine.size == theirs.size
            if mine.size:
                assert torch.equal(mine.k, theirs.k) and torch.equal(mine.v[0], theirs.v[0]) and torch.equal(mine.s, theirs.s)
    assert ref.long_mem.size > 0
