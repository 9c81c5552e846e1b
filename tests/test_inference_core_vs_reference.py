"""Row J on CPU: this package's InferenceCore against the LIVE reference InferenceCore (build container only), both driven
with the same fake network and with `match_memory` stubbed, so that only the per-frame state machine is compared:
memory-frame / deep-update scheduling, hidden-state hand-over, user-mask merging (full and partial labels),
`manually_curated_masks`, `disable_memory_updates`, `end`, and what gets written to which memory bank
(reference inference/inference_core.py:62-179)."""
import os
import sys

import pytest
import torch

from xmem2_b200 import lib
from xmem2_b200.inference import kv_memory_store as kv
from xmem2_b200.inference.inference_core import InferenceCore as MyCore

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present on this box')
H, W = 32, 48
h, w = H // 16, W // 16


class FakeNet(torch.nn.Module):
    """Deterministic stand-in with the XMem call surface; logs every call with its flags."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.log = []

    def encode_key(self, image, need_sk=True, need_ek=True):
        self.log.append(('encode_key', bool(need_sk), bool(need_ek)))
        m = image.mean(dim=(1, 2, 3)).view(1, 1, 1, 1)
        base = torch.arange(64 * h * w, dtype=torch.float32).view(1, 64, h, w) / (64 * h * w)
        key = (base + m).half().float()
        shr = torch.ones(1, 1, h, w) + 0.5
        sel = (torch.full((1, 64, h, w), 0.5)).half().float()
        f16 = torch.zeros(1, 8, h, w) + m
        return key, shr if need_sk else None, sel if need_ek else None, f16, f16, f16

    def segment(self, feats, readout, hidden, selector=None, h_out=True, strip_bg=True):
        n = readout.shape[1]
        self.log.append(('segment', bool(h_out), bool(strip_bg), n))
        m = feats[0].mean()
        yy = torch.linspace(-2, 2, H).view(1, 1, H, 1); xx = torch.linspace(-2, 2, W).view(1, 1, 1, W)
        logits = torch.cat([(yy + xx * (o + 1) + m) for o in range(n)], 1)
        prob = torch.sigmoid(logits)
        bg = torch.prod(1 - prob, dim=1, keepdim=True)
        p = torch.cat([bg, prob], 1).clamp(1e-7, 1 - 1e-7)
        p = torch.softmax(torch.log(p / (1 - p)), dim=1)
        new_h = hidden + 1.0 if h_out else None
        return new_h, None, (p[:, 1:] if strip_bg else p)

    def encode_value(self, frame, f16, h16, masks, is_deep_update=True):
        n = masks.shape[1]
        self.log.append(('encode_value', bool(is_deep_update), n))
        val = (masks.mean(dim=(2, 3)).view(1, n, 1, 1, 1) + torch.zeros(1, n, 512, h, w) + f16.mean()).half().float()
        return val, (h16 * 0.5 + 10.0 if is_deep_update else h16)


@pytest.fixture
def ref_core_cls(monkeypatch):
    def key_pack(key_rows, dst_rows):
        k = key_rows.float()
        dst_rows[:, :64] = (k * k).half()
        dst_rows[:, 64:] = key_rows
    monkeypatch.setattr(lib, 'require_cuda', lambda t, name: None)
    monkeypatch.setattr(lib, 'key_pack', key_pack)
    monkeypatch.setattr(kv, '_ARENA_POOL', {})
    saved = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('inference', 'model', 'util')}
    for k in saved:
        monkeypatch.delitem(sys.modules, k)
    monkeypatch.syspath_prepend(REF)
    import inference.inference_core as ref_ic
    real_zeros = torch.zeros

    def cpu_zeros(*a, **k):
        k.pop('device', None)
        return real_zeros(*a, **k)

    def build(net, cfg):
        torch.zeros = cpu_zeros                  # warm-up bypass (inference_core.py:26 hard-codes cuda:0)
        try:
            return ref_ic.InferenceCore(net, config=cfg)
        finally:
            torch.zeros = real_zeros
    yield build
    for k in [k for k in sys.modules if k.split('.')[0] in ('inference', 'model', 'util')]:
        sys.modules.pop(k, None)
    sys.modules.update(saved)


def _cfg(**over):
    cfg = dict(mem_every=3, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True, hidden_dim=64,
               key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
               max_long_term_elements=10000, use_cuda_graph=False)
    cfg.update(over)
    return cfg


def _pair(ref_build, cfg):
    na, nb = FakeNet(), FakeNet()
    mine, ref = MyCore(na, dict(cfg)), ref_build(nb, dict(cfg))
    stub = lambda core: (lambda key, sel, disable_usage_updates=False: torch.zeros(len(core.all_labels), 512, h, w))
    na.log.clear(); nb.log.clear()
    return mine, ref, na, nb, stub


def _masks(ti, n):
    yy = torch.arange(H).view(H, 1); xx = torch.arange(W).view(1, W)
    out = torch.zeros(n, H, W)
    for o in range(n):
        out[o] = ((yy - 8 - 2 * o - ti) ** 2 + (xx - 12 - 9 * o) ** 2 <= 36).float()
    return out


def _compare(mine, ref, na, nb):
    assert na.log == nb.log
    assert (mine.curr_ti, mine.last_mem_ti) == (ref.curr_ti, ref.last_mem_ti)
    assert mine.memory.temporary_work_mem.size == ref.memory.temporary_work_mem.size
    assert mine.memory.permanent_work_mem.size == ref.memory.permanent_work_mem.size
    hm, hr = mine.memory.get_hidden(), ref.memory.get_hidden()
    assert (hm is None) == (hr is None)
    if hm is not None:
        assert torch.allclose(hm.float(), hr.float())


@pytest.mark.parametrize('deep_every', [-1, 2])
def test_driver_style_clip(ref_core_cls, deep_every):
    mine, ref, na, nb, stub = _pair(ref_core_cls, _cfg(deep_update_every=deep_every))
    for core in (mine, ref):
        core.set_all_labels([1])
        core.put_to_permanent_memory(torch.ones(3, H, W) * 0.1, _masks(0, 1))
        core.memory.match_memory = stub(core)
    _compare(mine, ref, na, nb)
    for ti in range(11):
        img = torch.ones(3, H, W) * (0.1 + 0.05 * ti)
        msk = _masks(ti, 1) if ti in (0, 5) else None
        kw = dict(end=(ti == 10), do_not_add_mask_to_memory=msk is not None)
        pm = mine.step(img, msk.clone() if msk is not None else None, [1] if msk is not None else None, **kw)
        pr = ref.step(img, msk.clone() if msk is not None else None, [1] if msk is not None else None, **kw)
        assert torch.allclose(pm.float(), pr.float(), atol=1e-6), ti
        _compare(mine, ref, na, nb)


def test_partial_labels_merge_prediction_and_user_mask(ref_core_cls):
    mine, ref, na, nb, stub = _pair(ref_core_cls, _cfg())
    for core in (mine, ref):
        core.set_all_labels([1, 2])
        core.put_to_permanent_memory(torch.ones(3, H, W) * 0.2, _masks(0, 2))
        core.memory.match_memory = stub(core)
    for ti in range(5):
        img = torch.ones(3, H, W) * (0.2 + 0.03 * ti)
        msk = _masks(ti, 2) if ti == 2 else None
        if msk is not None:
            msk[1] = 0                                  # only label 1 is annotated on this frame
        vl = [1] if msk is not None else None
        pm = mine.step(img, msk.clone() if msk is not None else None, vl)
        pr = ref.step(img, msk.clone() if msk is not None else None, vl)
        assert torch.allclose(pm.float(), pr.float(), atol=1e-6), ti
        _compare(mine, ref, na, nb)


def test_curated_masks_and_disabled_updates(ref_core_cls):
    mine, ref, na, nb, stub = _pair(ref_core_cls, _cfg(mem_every=2))
    for core in (mine, ref):
        core.set_all_labels([1])
        core.put_to_permanent_memory(torch.ones(3, H, W) * 0.3, _masks(0, 1))
        core.memory.match_memory = stub(core)
    plan = [dict(), dict(manually_curated_masks=True), dict(disable_memory_updates=True), dict(manually_curated_masks=True),
            dict(), dict(disable_memory_updates=True), dict(end=True)]
    for ti, kw in enumerate(plan):
        img = torch.ones(3, H, W) * (0.3 + 0.02 * ti)
        msk = _masks(ti, 1) if ti == 3 else None
        pm = mine.step(img, msk.clone() if msk is not None else None, [1] if msk is not None else None, **kw)
        pr = ref.step(img, msk.clone() if msk is not None else None, [1] if msk is not None else None, **kw)
        assert torch.allclose(pm.float(), pr.float(), atol=1e-6), ti
        _compare(mine, ref, na, nb)
