"""world_size-2 gloo test (CPU) of the stream-parallel plumbing used by `bench.py --gpus N`."""
import os
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xmem2_b200.util import dist as xd


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = xd.streams_for_rank(5, rank, world)
    ms = 100.0 if rank == 0 else 250.0
    fps = xd.whole_job_throughput(frames_local=100 * len(mine), ms_local=ms, device='cpu')
    out.put((rank, mine, fps, xd.max_over_ranks(ms, 'cpu')))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_streams_and_time_by_the_slowest():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, f0, m0), (r1, s1, f1, m1) = res
    assert sorted(s0 + s1) == [0, 1, 2, 3, 4] and not set(s0) & set(s1)
    assert m0 == m1 == 250.0
    assert abs(f0 - 500 / 0.25) < 1e-6 and f0 == f1          # 500 frames / slowest rank's 250 ms


def test_single_process_is_identity():
    assert xd.max_over_ranks(3.5, 'cpu') == 3.5
    assert xd.streams_for_rank(3, 0, 1) == [0, 1, 2]
    assert xd.stream_seed(1234, 2) == 3234
