"""T-sharded memory read across 2 GPUs (NCCL) == the single-GPU read of the whole memory.  Needs two
CUDA devices; skipped on the 1-GPU boxes.  Also checks the degenerate 1-rank 'sharding' on any GPU box."""
import json
import os
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _launch(world, hw, frame, n_frames, n_obj):
    from tests import tshard_worker
    out = os.path.join(tempfile.mkdtemp(), 'res.json')
    port = 29600 + (os.getpid() % 1000)
    mp.spawn(tshard_worker._spawned, args=(world, port, hw, frame, n_frames, n_obj, 0, out), nprocs=world, join=True)
    return json.load(open(out))


def test_one_rank_staged_read_equals_fused_read():
    r = _launch(1, 200, 150, 5, 1)
    assert r['max_abs_diff'] <= 2e-3 * max(1.0, r['ref_absmax']) and r['oracle_err_clear'] < 2e-2, r


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_ranks_frame_sharded_read_equals_single_gpu():
    r = _launch(2, 300, 300, 7, 2)          # 7 frames: rank 0 owns 4, rank 1 owns 3
    assert r['max_abs_diff'] <= 2e-3 * max(1.0, r['ref_absmax']) and r['oracle_err_clear'] < 2e-2, r


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_rank_with_fewer_columns_than_topk():
    r = _launch(2, 128, 20, 3, 1)           # rank 1 owns a single 20-column frame (< top_k = 30)
    assert r['max_abs_diff'] <= 2e-3 * max(1.0, r['ref_absmax']) and r['oracle_err_clear'] < 2e-2, r
