"""GPU parity of the long-term memory maintenance kernels (csrc/consolidate.cu) against the oracle's
restatement of inference/memory_manager.py:349-390 and inference/kv_memory_store.py:125-181, at the BASELINE shapes:
HW = 1620 (480p: 5 candidate frames = 8100 columns) and HW = 8160 (1080p: 40 800 columns)."""
import pytest
import torch

from oracle import xmem_oracle as O
from xmem2_b200 import lib
from xmem2_b200.inference.kv_memory_store import KeyValueMemoryStore
from xmem2_b200.inference.memory_manager import MemoryManager

pytestmark = pytest.mark.gpu
CK, CV = 64, 512
torch.set_grad_enabled(False)


def test_usage_topk_matches_torch_topk():
    g = torch.Generator().manual_seed(1)
    n = 8100
    use = torch.rand(n, generator=g) * 3
    use[torch.randperm(n, generator=g)[:500]] = 0.0            # exact ties at zero
    life = torch.rand(n, generator=g) * 20 + 1
    idx = lib.usage_topk(use.cuda(), life.cuda(), 128).cpu().long()
    r = use / life
    ref_vals = torch.topk(r, 128, sorted=True).values
    assert torch.equal(r[idx], ref_vals)                        # same values in the same (descending) order
    assert len(set(idx.tolist())) == 128
    ties = r[idx][1:] == r[idx][:-1]
    assert bool((idx[1:][ties] > idx[:-1][ties]).all())         # ties by ascending index


def test_eviction_list_matches_reference_rule():
    g = torch.Generator().manual_seed(2)
    n, max_size = 10128, 9872
    use = torch.rand(n, generator=g); life = torch.rand(n, generator=g) * 50 + 1
    use[:300] = 0.0
    keep, m = lib.usage_evict_list(use.cuda(), life.cuda(), n, n - max_size)
    r = use / life
    thr = torch.topk(r, k=n - max_size, largest=False, sorted=True).values[-1]
    want = torch.nonzero(r > thr).flatten()
    assert m == want.numel()
    assert torch.equal(keep[:m].cpu().long(), want)


def test_bank_compaction_shift_and_gather():
    g = torch.Generator().manual_seed(3)
    cap, n_obj, n = 4096, 2, 3500
    dev = 'cuda'
    kp = torch.randn(cap, 128, generator=g).half().to(dev); s = torch.rand(cap, generator=g).to(dev)
    e = torch.rand(cap, 64, generator=g).half().to(dev); use = torch.rand(cap, generator=g).to(dev); life = torch.rand(cap, generator=g).to(dev)
    v = torch.randn(n_obj, CV, cap, generator=g).half().to(dev)
    ref = [t.clone() for t in (kp, s, e, use, life, v)]
    # overlapping shift: drop columns [100, 300) of 3500 -> tail of 3200 moves down by 200
    lib.bank_compact(kp, s, e, use, life, v, None, 200, 100, 100 + 3200)
    for t, r in zip((kp, s, e, use, life), ref[:5]):
        assert torch.equal(t[100:3300], r[300:3500]) and torch.equal(t[:100], r[:100])
    assert torch.equal(v[:, :, 100:3300], ref[5][:, :, 300:3500]) and torch.equal(v[:, :, :100], ref[5][:, :, :100])
    # gather by a survivor list
    ref = [t.clone() for t in (kp, s, e, use, life, v)]
    keep = torch.nonzero(torch.rand(3300, generator=g) > 0.3).flatten().to(torch.int32).to(dev)
    m = keep.numel()
    lib.bank_compact(kp, s, e, use, life, v, keep, 0, 0, m)
    for t, r in zip((kp, s, e, use, life), ref[:5]):
        assert torch.equal(t[:m], r[keep.long()])
    assert torch.equal(v[:, :, :m], ref[5][:, :, keep.long()])


def _cfg(**over):
    cfg = dict(hidden_dim=64, top_k=30, enable_long_term=True, enable_long_term_count_usage=True, max_mid_term_frames=10,
               min_mid_term_frames=5, num_prototypes=128, max_long_term_elements=10000, key_dim=64, value_dim=512)
    cfg.update(over)
    return cfg


@pytest.mark.parametrize('hw,groups', [(1620, 1), (1620, 2), (8160, 1)])
def test_consolidation_matches_oracle(hw, groups):
    """MemoryManager.consolidation (kernels) vs OracleMemory.consolidate (torch fp32) on the same fp16-valued candidates."""
    g = torch.Generator().manual_seed(10 + hw + groups)
    n = 5 * hw
    dev = 'cuda'
    key = (torch.randn(1, CK, n, generator=g) * 0.5).half().float()
    shr = torch.rand(1, 1, n, generator=g) * 2 + 1
    sel = torch.rand(1, CK, n, generator=g).half().float()
    usage = torch.rand(1, 1, n, generator=g)
    vals = [torch.randn(1, CV, n, generator=g).half().float()]
    if groups == 2:
        vals.append(torch.randn(1, CV, 3 * hw, generator=g).half().float())        # second group: the last 3 frames only
    cfg = _cfg()
    om = O.OracleMemory(dict(cfg))
    pk, pv, ps = om.consolidate(key, shr, sel, usage, vals)
    mm = MemoryManager(dict(cfg))
    # values as arena-like strided views (channel pitch > columns) to exercise the pitch argument
    gvals = []
    for v in vals:
        arena = torch.zeros(v.shape[0], CV, v.shape[2] + 64, dtype=torch.float16, device=dev)
        arena[:, :, :v.shape[2]] = v.half().to(dev)
        gvals.append(arena[:, :, :v.shape[2]])
    gk, gv, gs = mm.consolidation(key.half().to(dev), shr.to(dev), sel.half().to(dev), usage.to(dev), gvals)
    torch.cuda.synchronize()
    assert torch.equal(gk.float().cpu(), pk)                      # same prototypes (usage top-128, no ties in this data)
    assert torch.allclose(gs.cpu(), ps, rtol=2e-3, atol=2e-3), (gs.cpu() - ps).abs().max()
    for a, b in zip(gv, pv):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape
            err = (a.float().cpu() - b).abs().max().item()
            assert err < 2e-2, err                                  # fp16 output of an fp32 accumulation of O(1) values


def test_compress_features_and_eviction_on_the_arena_follow_the_oracle():
    """Whole write path on the GPU arenas (add -> compress -> evict) against the oracle memory, small frames."""
    H, W = 6, 9
    hw = H * W
    cfg = _cfg(max_mid_term_frames=6, min_mid_term_frames=3, num_prototypes=16, max_long_term_elements=40)
    mm, om = MemoryManager(dict(cfg)), O.OracleMemory(dict(cfg))
    g = torch.Generator().manual_seed(5)
    dev = 'cuda'
    for t in range(20):
        key = (torch.randn(1, CK, H, W, generator=g) * 0.5).half()
        shr = torch.rand(1, 1, H, W, generator=g) + 1
        sel = torch.rand(1, CK, H, W, generator=g).half()
        val = torch.randn(1, 1, CV, H, W, generator=g).half()
        # fake usage so that the prototype choice is deterministic and identical on both sides
        mm.add_memory(key.to(dev), shr.to(dev), val.to(dev), [1], selection=sel.to(dev))
        om.add(key.float(), shr, val.float(), [1], selection=sel.float())
        for mine, theirs in ((mm.temporary_work_mem, om.temp), (mm.long_mem, om.long)):
            if mine.size:
                u = torch.rand(mine.size, generator=g)
                mine._use[:mine.size] = u.to(dev); mine._life[:mine.size] = 1.0
                theirs.use = u.view(1, 1, -1).clone(); theirs.life = torch.ones(1, 1, mine.size)
        assert [mm.temporary_work_mem.size, mm.long_mem.size] == [om.temp.size, om.long.size], t
    torch.cuda.synchronize()
    assert om.long.size > 0 and mm.long_mem.size == om.long.size
    assert torch.allclose(mm.long_mem.k.float().cpu(), om.long.k, atol=2e-3)
    assert torch.allclose(mm.long_mem.s.cpu(), om.long.s, rtol=5e-3, atol=5e-3)
    assert (mm.long_mem.v[0].float().cpu() - om.long.v[0]).abs().max().item() < 3e-2
    assert torch.allclose(mm.temporary_work_mem.k.float().cpu(), om.temp.k, atol=2e-3)
