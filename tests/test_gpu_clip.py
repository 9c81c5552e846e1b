"""End-to-end GPU parity: InferenceCore.step / put_to_permanent_memory traces on the synthetic clips whose
expected outputs were produced by the LIVE reference (tests/golden/make_golden.py): memory-bank sizes per frame
(bit-exact), object-group layout, probabilities, argmax maps, hidden state, usage statistics.
Covers working-memory growth, consolidation into long-term prototypes, least-used eviction, and a second object
group with suffix value ranges (reference inference/inference_core.py:62-179, memory_manager.py:61-390)."""
import os
import numpy as np
import pytest
import torch

from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict, synth_frame, synth_mask

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
torch.set_grad_enabled(False)

BASE = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True, hidden_dim=64,
            key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
            max_long_term_elements=10000)


@pytest.fixture(scope='module')
def net():
    n = XMem({}, None).to('cuda').eval()
    n.load_weights(synth_state_dict(0))
    return n


def _run(net, name):
    dev = 'cuda'
    d = np.load(os.path.join(G, f'clip_{name}.npz'))
    H, W, n_frames, n_obj, save_every = [int(x) for x in d['hw']]
    cfg = dict(BASE); cfg.update({str(k): int(v) for k, v in zip(d['cfg_keys'], d['cfg_vals'])})
    ffo = [int(x) for x in d['first_frame_of']]; annotated = [int(x) for x in d['annotated']]
    core = InferenceCore(net, cfg)
    n_seen = 0
    for j in [int(x) for x in d['order']]:
        n_seen = max(n_seen, sum(1 for f in ffo if f <= j))
        core.set_all_labels(list(range(1, n_seen + 1)))
        core.put_to_permanent_memory(synth_frame(j, H, W, structured=True).to(dev), synth_mask(j, H, W, n_obj, ffo)[:n_seen].to(dev))
    labels = list(range(1, n_seen + 1))
    k, worst_mean, worst_agree = 0, 0.0, 1.0
    for ti in range(n_frames):
        msk = synth_mask(ti, H, W, n_obj, ffo).to(dev) if ti in annotated else None
        p = core.step(synth_frame(ti, H, W, structured=True).to(dev), msk, labels if msk is not None else None,
                      end=(ti == n_frames - 1), do_not_add_mask_to_memory=msk is not None)
        assert p.shape == (n_obj + 1, H, W)
        m = core.memory
        assert [m.temporary_work_mem.size, m.permanent_work_mem.size, m.long_mem.size] == d['sizes'][ti][:3].tolist(), ti
        if ti % save_every == 0:
            ref = torch.from_numpy(d['probs'][k]).float(); k += 1
            e = (p.float().cpu() - ref).abs()
            worst_mean = max(worst_mean, e.mean().item())
            worst_agree = min(worst_agree, (p.argmax(0).cpu() == ref.argmax(0)).float().mean().item())
    m = core.memory
    assert m.permanent_work_mem.num_groups == int(d['sizes'][-1][3])
    for g in range(m.permanent_work_mem.num_groups):
        assert m.permanent_work_mem.get_v_size(g) == int(d['sizes'][-1][4 + g])
    # fp16 tensor-core pipeline vs the fp32 reference trace (see tests/test_gpu_network.py for the calibration)
    assert worst_mean < 1e-2, worst_mean
    assert worst_agree > 0.95, worst_agree
    hid = m.get_hidden().float().cpu()
    assert (hid - torch.from_numpy(d['final_hidden']).float()).abs().mean().item() < 5e-3
    if len(d['temp_usage']):
        # usage = sum of top-k weights / life: one top-k membership flip at a near-tie (fp16 keys vs the fp32
        # reference trace) moves an entry by ~1/30, so bound the mean tightly and the max loosely
        u = m.temporary_work_mem.get_usage().float().cpu()
        du = (u - torch.from_numpy(d['temp_usage'])).abs()
        assert du.mean().item() < 5e-3 and du.max().item() < 0.1, (du.mean().item(), du.max().item())
    return core


def test_one_object_long_term_consolidation_and_eviction(net):
    core = _run(net, 'one_obj')
    assert core.memory.long_mem.size == 72          # 56 survivors + 16 new prototypes (golden: measured on the reference)


def test_two_objects_two_groups(net):
    core = _run(net, 'two_obj')
    assert core.memory.permanent_work_mem.obj_groups == [[0], [1]]


def test_disable_memory_updates_does_not_advance_state(net):
    dev = 'cuda'
    H, W = 96, 128                                 # 48 memory columns per frame >= top_k
    core = InferenceCore(net, dict(BASE))
    core.set_all_labels([1])
    core.put_to_permanent_memory(synth_frame(0, H, W, structured=True).to(dev), synth_mask(0, H, W, 1).to(dev))
    core.step(synth_frame(0, H, W, structured=True).to(dev), synth_mask(0, H, W, 1).to(dev), [1], do_not_add_mask_to_memory=True)
    ti, use = core.curr_ti, core.memory.permanent_work_mem.size
    hidden = core.memory.get_hidden().clone()
    p = core.step(synth_frame(1, H, W, structured=True).to(dev), disable_memory_updates=True)
    assert core.curr_ti == ti and core.memory.permanent_work_mem.size == use
    assert torch.equal(core.memory.get_hidden(), hidden)
    assert torch.isfinite(p).all() and abs(p.sum(0).mean().item() - 1.0) < 1e-3


def test_two_interleaved_cores_share_one_recorded_graph_without_crosstalk(net):
    # two videos processed alternately on one network (the recorded frame graph and its static buffers are shared):
    # each must produce exactly what it produces when run alone
    dev = 'cuda'
    H, W = 96, 128

    def run(seed_shift, other=None):
        core = InferenceCore(net, dict(BASE))
        core.set_all_labels([1])
        f = lambda ti: synth_frame(ti + seed_shift, H, W, structured=True).to(dev)
        core.put_to_permanent_memory(f(0), synth_mask(0, H, W, 1).to(dev))
        core.step(f(0), synth_mask(0, H, W, 1).to(dev), [1], do_not_add_mask_to_memory=True)
        outs = []
        for ti in range(1, 8):
            outs.append(core.step(f(ti)).clone())
            if other is not None:
                other()
        return outs

    alone = run(0)
    # a second core that advances one frame every time the first one does
    state = {}

    def make_other():
        core = InferenceCore(net, dict(BASE))
        core.set_all_labels([1])
        g = lambda ti: synth_frame(ti + 50, H, W, structured=True).to(dev)
        core.put_to_permanent_memory(g(0), synth_mask(3, H, W, 1).to(dev))
        core.step(g(0), synth_mask(3, H, W, 1).to(dev), [1], do_not_add_mask_to_memory=True)
        state['ti'] = 1
        def advance():
            core.step(g(state['ti'])); state['ti'] += 1
        return advance

    mixed = run(0, other=make_other())
    for a, b in zip(alone, mixed):
        assert torch.equal(a, b)


def test_sequential_cores_reuse_the_graph_after_empty_cache(net):
    # a finished video's arenas are recycled, so the next video of the same shape hits the recorded graph; the graph's read
    # kernels point into the FIRST core's K1 workspace, which the graph entry keeps alive and the new core adopts.  With the
    # allocator emptied in between, a dangling workspace pointer would give garbage or fault (round-1 advisor finding).
    import gc
    dev = 'cuda'
    H, W = 96, 128

    def run(use_graph):
        cfg = dict(BASE); cfg['use_cuda_graph'] = use_graph
        core = InferenceCore(net, cfg)
        core.set_all_labels([1])
        f = lambda ti: synth_frame(ti, H, W, structured=True).to(dev)
        core.put_to_permanent_memory(f(0), synth_mask(0, H, W, 1).to(dev))
        core.step(f(0), synth_mask(0, H, W, 1).to(dev), [1], do_not_add_mask_to_memory=True)
        outs = [core.step(f(ti)).clone() for ti in range(1, 9)]
        sizes = (core.memory.temporary_work_mem.size, core.memory.permanent_work_mem.size)
        del core
        return outs, sizes

    eager, sizes_e = run(False)
    first, sizes_1 = run(True)
    gc.collect(); torch.cuda.synchronize(); torch.cuda.empty_cache()
    junk = [torch.full((1 << 22,), float('nan'), device=dev) for _ in range(8)]      # occupy whatever was just freed
    second, sizes_2 = run(True)
    del junk
    assert sizes_e == sizes_1 == sizes_2
    for a, b, c in zip(eager, first, second):
        assert torch.isfinite(c).all()
        assert (a - b).abs().max().item() < 2e-3 and (a - c).abs().max().item() < 2e-3
