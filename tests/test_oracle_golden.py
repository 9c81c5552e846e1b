"""Pin the CPU oracle (oracle/xmem_oracle.py) against fixtures produced by the live reference
(tests/golden/make_golden.py).  Runs without a GPU."""
import os
import numpy as np
import torch

from oracle import xmem_oracle as O
from xmem2_b200.util.synth import synth_state_dict, synth_frame, synth_mask, xmem_param_spec

G = os.path.join(os.path.dirname(__file__), 'golden')
torch.set_grad_enabled(False)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_param_spec_matches_upstream_checkpoint_layout():
    spec = xmem_param_spec()
    assert len(spec) == 412                       # reference XMem().state_dict() size, measured
    assert spec['value_encoder.conv1.weight'][0] == (64, 5, 7, 7)
    assert spec['decoder.hidden_update.g4_conv.weight'][0] == (256, 257, 1, 1)
    assert spec['decoder.fuser.block1.conv1.weight'][0] == (512, 1600, 3, 3)


def test_attention_math_matches_reference():
    d = np.load(os.path.join(G, 'attention.npz'))
    sim = O.similarity(_t(d['mk']), _t(d['ms']), _t(d['qk']), _t(d['qe']))
    assert torch.equal(sim, _t(d['sim']))
    aff, usage = O.softmax_topk(sim, 30, want_usage=True)
    assert torch.allclose(usage, _t(d['usage']), atol=1e-6)
    assert torch.allclose(_t(d['v']) @ aff, _t(d['readout']), atol=1e-5)
    assert torch.allclose(_t(d['v']) @ O.softmax_topk(sim, None), _t(d['full_readout']), atol=1e-5)


def test_network_passes_match_reference():
    d = np.load(os.path.join(G, 'network.npz'))
    net = O.OracleNet(synth_state_dict(0))
    H, W = 64, 96
    img = synth_frame(0, H, W, structured=True)[None]
    masks = synth_mask(0, H, W, 2)[None]
    key, shr, sel, f16, f8, f4 = net.encode_key(img)
    tol = dict(atol=2e-4, rtol=1e-4)
    assert torch.allclose(key, _t(d['key']), **tol) and torch.allclose(shr, _t(d['shr']), **tol)
    assert torch.allclose(sel, _t(d['sel']), **tol) and torch.allclose(f16[:, ::16], _t(d['f16']), **tol)
    assert torch.allclose(f8[:, ::32, ::2, ::2], _t(d['f8']), **tol) and torch.allclose(f4[:, ::32, ::4, ::4], _t(d['f4']), **tol)
    hid = _t(d['hid'])
    val, hid2 = net.encode_value(img, f16, hid, masks, True)
    assert torch.allclose(val[:, :, ::8], _t(d['val']), **tol) and torch.allclose(hid2, _t(d['hid2']), **tol)
    nh, logits, prob = net.segment((f16, f8, f4), _t(d['ro']), hid, True, False)
    assert torch.allclose(nh, _t(d['nh']), **tol)
    assert torch.allclose(logits[:, :, ::2, ::2], _t(d['logits']), atol=2e-3, rtol=1e-3)
    assert torch.allclose(prob[:, :, ::2, ::2], _t(d['prob']), atol=1e-3)


def _run_clip(name):
    d = np.load(os.path.join(G, f'clip_{name}.npz'))
    H, W, n_frames, n_obj, save_every = [int(x) for x in d['hw']]
    cfg = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True,
               hidden_dim=64, key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5,
               num_prototypes=128, max_long_term_elements=10000)
    cfg.update({str(k): int(v) for k, v in zip(d['cfg_keys'], d['cfg_vals'])})
    ffo = [int(x) for x in d['first_frame_of']]
    annotated = [int(x) for x in d['annotated']]
    core = O.OracleCore(O.OracleNet(synth_state_dict(0)), cfg)
    n_seen = 0
    for j in [int(x) for x in d['order']]:
        n_seen = max(n_seen, sum(1 for f in ffo if f <= j))
        core.set_all_labels(list(range(1, n_seen + 1)))
        core.put_to_permanent_memory(synth_frame(j, H, W, structured=True), synth_mask(j, H, W, n_obj, ffo)[:n_seen])
    labels = list(range(1, n_seen + 1))
    mean_err, k = 0.0, 0
    for ti in range(n_frames):
        msk = synth_mask(ti, H, W, n_obj, ffo) if ti in annotated else None
        p = core.step(synth_frame(ti, H, W, structured=True), msk, labels if msk is not None else None,
                      end=(ti == n_frames - 1), do_not_add_mask_to_memory=msk is not None)
        long_size = core.mem.long.size if core.mem.long is not None else 0
        assert [core.mem.temp.size, core.mem.perm.size, long_size] == d['sizes'][ti][:3].tolist(), ti
        if ti % save_every == 0:
            ref = _t(d['probs'][k]).float(); k += 1
            err = (p - ref).abs()
            assert err.max().item() < 6e-2, (ti, err.max().item())
            mean_err = max(mean_err, err.mean().item())
    assert mean_err < 5e-4                         # fixtures are stored as fp16
    assert core.mem.perm.num_groups == int(d['sizes'][-1][3])


def test_clip_one_object_trace_matches_reference():
    _run_clip('one_obj')


def test_clip_two_objects_two_groups_trace_matches_reference():
    _run_clip('two_obj')


def test_clip_free_running_deep_updates_trace_matches_reference():
    _run_clip('plain')         # deep_update_every = 3 (not synchronised with the memory frames), two annotated frames


def test_chair_example_clip_through_the_reference_reader():
    # BASELINE.json config 1: example_videos/chair via the reference's own VideoReader (size=160), frame 0 annotated and
    # preloaded, 8 frames; fixture = reference outputs on CPU (tests/golden/make_golden.py chair)
    d = np.load(os.path.join(G, 'clip_chair.npz'))
    cfg = dict(mem_every=int(d['mem_every']), deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True,
               hidden_dim=64, key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5,
               num_prototypes=128, max_long_term_elements=10000)
    core = O.OracleCore(O.OracleNet(synth_state_dict(0)), cfg)
    labels = [int(x) for x in d['labels']]
    rgb = _t(d['rgb']).float()
    msk = _t(d['mask0']).float()
    core.set_all_labels(labels)
    core.put_to_permanent_memory(rgb[0], msk.clone())
    n = rgb.shape[0]
    for ti in range(n):
        m = msk.clone() if ti == 0 else None
        p = core.step(rgb[ti], m, labels if m is not None else None, end=(ti == n - 1), do_not_add_mask_to_memory=m is not None)
        ref = _t(d['probs'][ti]).float()
        assert (p - ref).abs().mean().item() < 5e-4 and (p.argmax(0) == ref.argmax(0)).float().mean().item() > 0.999
    assert core.mem.temp.size == int(d['temp_size']) and core.mem.perm.size == int(d['perm_size'])
