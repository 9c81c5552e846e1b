"""Generate golden fixtures from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference (read-only), loads the
hash-seeded synthetic parameters from `xmem2_b200.util.synth`, runs the reference's own
`InferenceCore` / `MemoryManager` / `get_similarity` / `do_softmax` on seeded synthetic
inputs (CPU, fp32) and stores small fixtures next to this script.  It also asserts that the
oracle restatement (`oracle/xmem_oracle.py`) reproduces the reference on every fixture.

Shims needed to import the reference here: a stub `progressbar` module and a
bypass of the hard-coded `cuda:0` warm-up in `InferenceCore.__init__` (inference_core.py:26).
"""
import os, sys, types, io, contextlib
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REF); sys.path.insert(0, REPO)

pb = types.ModuleType('progressbar')
class _PB:
    def __init__(self, *a, **k): pass
    def update(self, *a, **k): pass
    def finish(self, *a, **k): pass
pb.ProgressBar = _PB; pb.progressbar = lambda x, *a, **k: x
sys.modules['progressbar'] = pb

from xmem2_b200.util.synth import synth_state_dict, synth_frame, synth_mask   # noqa: E402
from oracle import xmem_oracle as O                                          # noqa: E402

torch.set_grad_enabled(False)
torch.manual_seed(0)


def ref_modules():
    from model.network import XMem
    from inference.inference_core import InferenceCore
    from inference.memory_manager import MemoryManager
    import model.memory_util as mu
    return XMem, InferenceCore, MemoryManager, mu


def make_ref_core(XMem, InferenceCore, cfg, state):
    with contextlib.redirect_stdout(io.StringIO()):
        net = XMem(cfg, None, pretrained_key_encoder=False, pretrained_value_encoder=False).eval()
    net.load_state_dict(state)
    real_zeros = torch.zeros
    def cpu_zeros(*a, **k):
        k.pop('device', None); return real_zeros(*a, **k)
    torch.zeros = cpu_zeros                      # warm-up bypass (inference_core.py:26)
    try:
        core = InferenceCore(net, config=cfg)
    finally:
        torch.zeros = real_zeros
    return net, core


def base_cfg(**over):
    cfg = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True,
               hidden_dim=64, key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5,
               num_prototypes=128, max_long_term_elements=10000)
    cfg.update(over); return cfg


def golden_attention(mu):
    """Rows B-D: get_similarity / do_softmax / readout on random tensors (memory_util.py:7-65)."""
    g = torch.Generator().manual_seed(7)
    CK, N, Q, CV = 64, 700, 96, 32
    mk = torch.randn(1, CK, N, generator=g) * 0.6
    ms = torch.rand(1, 1, N, generator=g) * 2 + 1
    qk = torch.randn(1, CK, Q, generator=g) * 0.6
    qe = torch.rand(1, CK, Q, generator=g)
    v = torch.randn(2, CV, N, generator=g)
    sim = mu.get_similarity(mk, ms, qk, qe)
    aff, usage = mu.do_softmax(sim.clone(), top_k=30, inplace=False, return_usage=True)
    full = mu.do_softmax(sim.clone(), top_k=None)
    ro = v @ aff
    o_sim = O.similarity(mk, ms, qk, qe)
    o_aff, o_usage = O.softmax_topk(o_sim, 30, want_usage=True)
    assert torch.equal(sim, o_sim) and torch.equal(aff, o_aff) and torch.equal(usage, o_usage)
    assert torch.equal(full, O.softmax_topk(o_sim, None))
    np.savez_compressed(os.path.join(HERE, 'attention.npz'), mk=mk.numpy(), ms=ms.numpy(), qk=qk.numpy(), qe=qe.numpy(),
                        v=v.numpy(), sim=sim.numpy(), usage=usage.numpy(), readout=ro.numpy(),
                        topk_vals=torch.topk(sim, 30, dim=1)[0].numpy(), full_readout=(v @ full).numpy())
    print('attention golden ok')


def golden_network(XMem, InferenceCore):
    """Rows A,F,G: encode_key / encode_value / segment on one 64x96 frame with 2 objects."""
    state = synth_state_dict(0)
    cfg = base_cfg()
    net, _ = make_ref_core(XMem, InferenceCore, cfg, state)
    H, W = 64, 96
    img = synth_frame(0, H, W, structured=True)[None]
    masks = synth_mask(0, H, W, 2)[None]
    key, shr, sel, f16, f8, f4 = net.encode_key(img)
    hid = torch.randn(1, 2, 64, H // 16, W // 16, generator=torch.Generator().manual_seed(3)) * 0.5
    val, hid2 = net.encode_value(img, f16, hid, masks, is_deep_update=True)
    ro = torch.randn(1, 2, 512, H // 16, W // 16, generator=torch.Generator().manual_seed(4))
    nh, logits, prob = net.segment((f16, f8, f4), ro, hid, h_out=True, strip_bg=False)
    on = O.OracleNet(state)
    ok, os_, oe, of16, of8, of4 = on.encode_key(img)
    ov, oh2 = on.encode_value(img, of16, hid, masks, True)
    onh, ologits, oprob = on.segment((of16, of8, of4), ro, hid, True, False)
    for a, b, n in ((key, ok, 'key'), (shr, os_, 'shr'), (sel, oe, 'sel'), (f16, of16, 'f16'), (f8, of8, 'f8'), (f4, of4, 'f4'),
                    (val, ov, 'val'), (hid2, oh2, 'hid2'), (nh, onh, 'nh'), (logits, ologits, 'logits'), (prob, oprob, 'prob')):
        err = (a - b).abs().max().item()
        assert err <= 1e-4 * max(1.0, a.abs().max().item()), (n, err)
    np.savez_compressed(os.path.join(HERE, 'network.npz'), hid=hid.numpy(), ro=ro.numpy(),
                        key=key.numpy(), shr=shr.numpy(), sel=sel.numpy(), f16=f16[:, ::16].numpy(), f8=f8[:, ::32, ::2, ::2].numpy(),
                        f4=f4[:, ::32, ::4, ::4].numpy(), val=val[:, :, ::8].numpy(), hid2=hid2.numpy(), nh=nh.numpy(),
                        logits=logits[:, :, ::2, ::2].numpy(), prob=prob[:, :, ::2, ::2].numpy())
    print('network golden ok')


def run_clip(XMem, InferenceCore, name, H, W, n_frames, n_obj, annotated, first_frame_of, cfg_over, save_every=1):
    """Rows E,H,I,J: full InferenceCore trace on a synthetic clip, driver-style (run_on_video.py:65-108)."""
    state = synth_state_dict(0)
    cfg = base_cfg(**cfg_over)
    net, core = make_ref_core(XMem, InferenceCore, dict(cfg), state)
    ocore = O.OracleCore(O.OracleNet(state), dict(cfg))
    order = list(set(annotated))                       # CPython set order, run_on_video.py:45,65-66,205
    n_seen = 0                                          # MaskMapper(exhaustive=True) semantics, mask_mapper.py:26-52
    for j in order:
        present = n_obj if first_frame_of is None else sum(1 for f in first_frame_of if f <= j)
        n_seen = max(n_seen, present)
        labels = list(range(1, n_seen + 1))
        core.set_all_labels(labels); ocore.set_all_labels(labels)
        img = synth_frame(j, H, W, structured=True); msk = synth_mask(j, H, W, n_obj, first_frame_of)[:n_seen]
        core.put_to_permanent_memory(img, msk); ocore.put_to_permanent_memory(img, msk.clone())
    labels = list(range(1, n_seen + 1))
    probs, sizes, maxerr, meanerr = [], [], 0.0, 0.0
    for ti in range(n_frames):
        img = synth_frame(ti, H, W, structured=True)
        msk = synth_mask(ti, H, W, n_obj, first_frame_of) if ti in annotated else None
        kw = dict(end=(ti == n_frames - 1), do_not_add_mask_to_memory=msk is not None)
        vl = labels if msk is not None else None
        p = core.step(img, msk, vl, **kw)
        po = ocore.step(img, msk.clone() if msk is not None else None, vl, **kw)
        maxerr = max(maxerr, (p - po).abs().max().item()); meanerr = max(meanerr, (p - po).abs().mean().item())
        m = core.memory
        sizes.append([m.temporary_work_mem.size, m.permanent_work_mem.size, m.long_mem.size if m.enable_long_term else 0,
                      m.temporary_work_mem.num_groups] + [m.permanent_work_mem.get_v_size(g) for g in range(m.permanent_work_mem.num_groups)][:2]
                     + [0] * (2 - min(2, m.permanent_work_mem.num_groups)))
        om = ocore.mem
        assert sizes[-1][:3] == [om.temp.size, om.perm.size, om.long.size if om.long is not None else 0], (ti, sizes[-1])
        if ti % save_every == 0:
            probs.append(p.numpy().astype(np.float16))
    # fp32 re-association noise is amplified by the decoder (logits reach +-16) and by rare top-k
    # membership flips at near-ties; the mean error is the meaningful check.
    assert maxerr < 5e-2 and meanerr < 2e-4, (maxerr, meanerr)
    hid = core.memory.get_hidden()
    assert (hid - ocore.mem.hidden).abs().max().item() < 2e-3
    np.savez_compressed(os.path.join(HERE, f'clip_{name}.npz'), probs=np.stack(probs), sizes=np.array(sizes, dtype=np.int32),
                        order=np.array(order), annotated=np.array(sorted(annotated)), hw=np.array([H, W, n_frames, n_obj, save_every]),
                        first_frame_of=np.array(first_frame_of if first_frame_of else [0] * n_obj),
                        final_hidden=hid.numpy().astype(np.float16),
                        temp_usage=(core.memory.temporary_work_mem.get_usage().numpy()
                                    if (core.memory.temporary_work_mem.size and core.memory.enable_long_term) else np.zeros(0)),
                        cfg_keys=np.array(list(cfg_over.keys())), cfg_vals=np.array(list(cfg_over.values())))
    print(f'clip {name}: oracle-vs-reference max prob err {maxerr:.2e} mean {meanerr:.2e}; final sizes {sizes[-1]}')


def golden_chair(XMem, InferenceCore):
    """BASELINE.json config 1: the reference's own example clip (example_videos/chair) through the reference's own reader
    (inference/data/video_reader.py) at size=160, frame 0 annotated and preloaded, 8 frames, CPU.  Inputs are stored as the
    fp16-rounded normalised tensors the reader produced, so no image decoding is needed to replay them."""
    from inference.data.video_reader import VideoReader
    from inference.data.mask_mapper import MaskMapper
    root = os.path.join(REF, 'example_videos', 'chair')
    reader = VideoReader('', os.path.join(root, 'JPEGImages'), os.path.join(root, 'Annotations'), size=160, use_all_masks=True)
    mapper = MaskMapper()
    n_frames = 8
    rgbs = [reader[i].rgb.half() for i in range(n_frames)]
    s0 = reader[0]
    msk, labels = mapper.convert_mask(s0.mask, exhaustive=True)
    msk = torch.Tensor(msk)
    if s0.need_resize:
        msk = reader.resize_mask(msk.unsqueeze(0))[0]
    state = synth_state_dict(0)
    cfg = base_cfg(mem_every=3)
    net, core = make_ref_core(XMem, InferenceCore, dict(cfg), state)
    ocore = O.OracleCore(O.OracleNet(state), dict(cfg))
    all_labels = list(mapper.remappings.values())
    for c in (core, ocore):
        c.set_all_labels(all_labels)
        c.put_to_permanent_memory(rgbs[0].float(), msk.clone())
    probs, err = [], 0.0
    for ti in range(n_frames):
        m = msk.clone() if ti == 0 else None
        kw = dict(end=(ti == n_frames - 1), do_not_add_mask_to_memory=m is not None)
        p = core.step(rgbs[ti].float(), m, list(labels) if m is not None else None, **kw)
        po = ocore.step(rgbs[ti].float(), m.clone() if m is not None else None, list(labels) if m is not None else None, **kw)
        err = max(err, (p - po).abs().mean().item())
        probs.append(p.numpy().astype(np.float16))
    assert err < 2e-4, err
    np.savez_compressed(os.path.join(HERE, 'clip_chair.npz'), rgb=torch.stack(rgbs).numpy(), mask0=msk.numpy().astype(np.uint8),
                        probs=np.stack(probs), labels=np.array(all_labels), mem_every=np.array(3),
                        temp_size=np.array(core.memory.temporary_work_mem.size), perm_size=np.array(core.memory.permanent_work_mem.size))
    print(f'chair golden ok: {tuple(rgbs[0].shape)} frames={n_frames} oracle-vs-reference mean err {err:.2e}; '
          f'temp={core.memory.temporary_work_mem.size} perm={core.memory.permanent_work_mem.size}')


from tests.selector_case import selector_inputs, SELECTOR_CASES          # noqa: E402


def golden_selector():
    """select_next_candidates of the LIVE reference (frame_selection.py:99-244) on CPU, plus the full
    matrix of pair scores it is built from (computed with the reference's get_similarity)."""
    from inference.frame_selection.frame_selection import select_next_candidates as ref_select
    from model.memory_util import get_similarity
    import torch.nn.functional as F
    keys, shr, sel, masks = selector_inputs()
    picks = []
    for c in SELECTOR_CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            r = ref_select(keys, shr, sel, masks, c['k'], previously_chosen_candidates=list(c['prev']), alpha=c['alpha'], device='cpu')
        o = O.select_next_candidates(keys, shr, sel, masks, c['k'], previously_chosen_candidates=list(c['prev']), alpha=c['alpha'])
        assert r == o, (c, r, o)
        picks.append(r + [-1] * (8 - len(r)))
    # pair scores at alpha = 0.5 for every ordered pair of valid frames (invalid ones: -1)
    valid, comp = O.composite_keys_for_selection(keys, masks, [0], 0.5, 0.25, 0.5)
    n = len(keys)
    scores = -np.ones((n, n), dtype=np.float32)
    for a in range(n):
        for b in range(n):
            if valid[a] and valid[b]:
                s_ab = get_similarity(comp[a].unsqueeze(0), shr[a].unsqueeze(0), comp[b].unsqueeze(0), sel[b].unsqueeze(0))
                s_ba = get_similarity(comp[b].unsqueeze(0), shr[b].unsqueeze(0), comp[a].unsqueeze(0), sel[a].unsqueeze(0))
                d = (s_ab - s_ba).float()
                scores[a, b] = (F.relu(d).sum() / d.numel()).item()
                assert abs(scores[a, b] - O.cycle_dissimilarity(comp[a], shr[a], sel[a], comp[b], shr[b], sel[b]).item()) < 1e-6
    np.savez_compressed(os.path.join(HERE, 'selector.npz'), picks=np.array(picks), valid=np.array(valid), scores=scores)
    print('selector golden ok:', picks, 'valid', valid)


if __name__ == '__main__':
    XMem, InferenceCore, MemoryManager, mu = ref_modules()
    if len(sys.argv) > 1 and sys.argv[1] == 'chair':
        golden_chair(XMem, InferenceCore)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'plain':
        # free-running deep updates (deep_update_every > 0: not synchronised with the memory frames), two annotated frames.
        # (enable_long_term=False cannot be pinned: the reference itself raises in add_memory, memory_manager.py:261,
        # `selection[..., 0:0]` with selection None, as soon as permanent memory is in use.)
        run_clip(XMem, InferenceCore, 'plain', 64, 96, 24, 1, [0, 12], None,
                 dict(mem_every=2, deep_update_every=3, max_mid_term_frames=6, min_mid_term_frames=3, num_prototypes=16), save_every=2)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'selector':
        golden_selector()
        sys.exit(0)
    golden_attention(mu)
    golden_network(XMem, InferenceCore)
    # one object, 5 annotated frames, long-term consolidation reached (HW=24: small so CPU is quick)
    run_clip(XMem, InferenceCore, 'one_obj', 64, 96, 60, 1, [0, 8, 16, 24, 32], None,
             dict(mem_every=2, max_mid_term_frames=6, min_mid_term_frames=3, num_prototypes=16, max_long_term_elements=72), save_every=3)
    # two objects, the second appears at the 2nd annotated frame -> two value groups with suffix ranges
    run_clip(XMem, InferenceCore, 'two_obj', 96, 64, 30, 2, [0, 6, 12, 18, 24], [0, 5],
             dict(mem_every=3, max_mid_term_frames=6, min_mid_term_frames=3, num_prototypes=16, max_long_term_elements=400), save_every=3)
