"""BASELINE.json config 1 on the GPU: the reference's own example clip (example_videos/chair, read by the reference's
VideoReader at size=160 when the fixture was generated) through this package's InferenceCore, against the outputs of the
LIVE reference on CPU (tests/golden/clip_chair.npz).  Added after the last GPU session of round 1, hence the generous
bounds (the synthetic clips measure mean |dprob| ~3.6e-3 and >= 97 % argmax agreement with the same pipeline) and the
file name that sorts it after every other GPU test."""
import os
import numpy as np
import pytest
import torch

from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
torch.set_grad_enabled(False)


def test_chair_clip_matches_reference_trace():
    dev = 'cuda'
    d = np.load(os.path.join(G, 'clip_chair.npz'))
    cfg = dict(mem_every=int(d['mem_every']), deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True,
               hidden_dim=64, key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5,
               num_prototypes=128, max_long_term_elements=10000)
    net = XMem(dict(cfg), None).to(dev).eval()
    net.load_weights(synth_state_dict(0))
    core = InferenceCore(net, cfg)
    labels = [int(x) for x in d['labels']]
    rgb = torch.from_numpy(d['rgb']).float().to(dev)
    msk = torch.from_numpy(d['mask0']).float().to(dev)
    core.set_all_labels(labels)
    core.put_to_permanent_memory(rgb[0], msk.clone())
    n = rgb.shape[0]
    worst_mean, worst_agree = 0.0, 1.0
    for ti in range(n):
        m = msk.clone() if ti == 0 else None
        p = core.step(rgb[ti], m, labels if m is not None else None, end=(ti == n - 1), do_not_add_mask_to_memory=m is not None)
        ref = torch.from_numpy(d['probs'][ti]).float()
        assert p.shape == ref.shape and torch.isfinite(p).all()
        worst_mean = max(worst_mean, (p.float().cpu() - ref).abs().mean().item())
        worst_agree = min(worst_agree, (p.argmax(0).cpu() == ref.argmax(0)).float().mean().item())
    assert core.memory.temporary_work_mem.size == int(d['temp_size'])
    assert core.memory.permanent_work_mem.size == int(d['perm_size'])
    assert worst_mean < 2e-2 and worst_agree > 0.9, (worst_mean, worst_agree)


def test_returned_probabilities_survive_later_frames():
    """step() hands out fresh tensors like the reference (inference_core.py:152): frames replayed from a recorded graph
    must not overwrite what an earlier call returned."""
    from xmem2_b200.util.synth import synth_frame, synth_mask
    dev = 'cuda'
    cfg = dict(mem_every=3, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True, hidden_dim=64,
               key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
               max_long_term_elements=10000)
    net = XMem(dict(cfg), None).to(dev).eval()
    net.load_weights(synth_state_dict(0))
    core = InferenceCore(net, cfg)
    core.set_all_labels([1])
    H, W = 96, 128
    kept = []
    for ti in range(9):
        m = synth_mask(ti, H, W, 1, [0]).to(dev) if ti == 0 else None
        p = core.step(synth_frame(ti, H, W, structured=True).to(dev), m, [1] if m is not None else None)
        kept.append((p, p.clone()))
    torch.cuda.synchronize()
    for ti, (p, snapshot) in enumerate(kept):
        assert torch.equal(p, snapshot), ti
    assert len({p.data_ptr() for p, _ in kept}) == len(kept)


def test_free_running_deep_updates_clip_matches_reference():
    """deep_update_every = 3 (reference inference_core.py:84-87: deep updates not synchronised with the memory frames; every
    frame runs eagerly except the ordinary ones) against tests/golden/clip_plain.npz, same bounds as the other clips."""
    from tests.test_gpu_clip import _run
    net = XMem({}, None).to('cuda').eval()
    net.load_weights(synth_state_dict(0))
    _run(net, 'plain')
