"""ONE video segmented by 2 ranks with a T-sharded memory == the live reference's trace of that clip.
Needs two CUDA devices (skipped on 1-GPU boxes).  Written after the last GPU session of round 1: never run on GPUs yet
(the host logic is covered on CPU by tests/test_sharded_core_gloo.py); named to sort after the verified tests."""
import json
import os
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_ranks_segment_one_clip_like_the_reference():
    from tests import tshard_clip_worker
    out = os.path.join(tempfile.mkdtemp(), 'res.json')
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(tshard_clip_worker._spawned, args=(2, port, 'one_obj', out), nprocs=2, join=True)
    for rank in range(2):
        r = json.load(open(f'{out}.{rank}'))
        assert not r['size_mismatch'], r                 # global bank sizes follow the reference frame by frame
        assert r['long_blocks'] >= 2, r                  # consolidated (and, with max 72 long-term columns, evicted)
        assert r['worst_mean'] < 1e-2 and r['worst_agree'] > 0.95, r      # same bounds as the single-GPU clip test
        assert r['rank_spread'] < 1e-3, r                # the ranks agree on the prediction
