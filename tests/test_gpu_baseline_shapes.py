"""End-to-end GPU parity at the BASELINE.json shapes (not toy sizes): config 2 (the clip bench.py times: 3x480x854 noise
frames, 1 object, 5 permanent-memory masks) and config 3's shape (portrait 853x480, 2 objects in two value groups,
through the first long-term consolidation), this package's InferenceCore against the oracle in fp32 on the same GPU,
frame by frame (tests/parity_clip.py).  Bars live in tests/golden/parity_bars.json (1.2 x measured on a B200) and are
printed next to the figures of the oracle under fp16 autocast (how the reference runs on a GPU) and the north star's 1e-3.
Reference: inference/inference_core.py:62-152, inference/memory_manager.py:61-390."""
import json
import os

import pytest
import torch

from tests.parity_clip import run_lockstep
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
HERE = os.path.dirname(os.path.abspath(__file__))
BARS = json.load(open(os.path.join(HERE, 'golden', 'parity_bars.json')))
CFG = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=False, hidden_dim=64,
           key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
           max_long_term_elements=10000)


@pytest.fixture(scope='module')
def setup():
    state = synth_state_dict(0)
    net = XMem(dict(CFG), None).to('cuda').eval()
    net.load_weights(dict(state))
    return state, net


def _record(name, ours, auto):
    out = os.path.join(os.path.dirname(HERE), 'gpurun_out')
    line = {'case': name, 'ours_vs_fp32_oracle': ours.as_dict(), 'autocast_oracle_vs_fp32_oracle': auto.as_dict() if auto else None,
            'north_star_logit_tol': 1e-3}
    print('\nPARITY', json.dumps(line))
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'parity_measured.jsonl'), 'a') as f:
            f.write(json.dumps(line) + '\n')
    except OSError:
        pass


def _assert_bars(name, stats):
    bars, got = BARS[name], stats.as_dict()
    for k, bar in bars.items():
        if k.startswith('_'):
            continue
        assert got[k] <= bar, (name, k, got[k], bar, got)


def test_config2_480p_clip_matches_fp32_oracle(setup):
    # the bench clip: 40 frames = 5 permanent preloads {0,20,40,60,80}, annotated frames 0 and 20, memory frames 10 and 30
    state, net = setup
    ours, auto, _ = run_lockstep(net, state, 480, 854, 40, 1, [0, 20, 40, 60, 80], None, CFG, structured=False, with_autocast=True)
    _record('config2_480p', ours, auto)
    _assert_bars('config2_480p', ours)
    # context, asserted loosely: this pipeline must not be further from fp32 than the reference's own fp16 path is
    assert ours.p50 <= 1.5 * auto.p50 + 1e-3 and ours.p99 <= 1.5 * auto.p99 + 1e-2, (ours.as_dict(), auto.as_dict())


def test_config3_portrait_two_objects_through_first_consolidation(setup):
    # 853x480 -> 54x30 grid, object 2 first appears at the second annotated frame (two value groups with suffix ranges),
    # the annotated frames are not added to the working memory (do_not_add_mask_to_memory), so it holds 10 frames at
    # ti = 140 -> first consolidation into 128 long-term prototypes, then 11 more frames read all three banks
    state, net = setup
    ours, _, (core, ocore) = run_lockstep(net, state, 853, 480, 152, 2, [0, 20, 40, 60, 80], [0, 20], CFG, structured=True)
    _record('config3_portrait_2obj', ours, None)
    assert core.memory.long_mem.size == 128 and ocore.mem.long.size == 128
    assert core.memory.permanent_work_mem.obj_groups == [[0], [1]]
    _assert_bars('config3_portrait_2obj', ours)
