"""Worker for the T-sharded read test / benchmark (one process per GPU, NCCL).  Launched by tests/test_gpu_tshard.py
via torch.multiprocessing or by `python -m torch.distributed.run --nproc-per-node R tests/tshard_worker.py`."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import k1_ref
from xmem2_b200 import lib
from xmem2_b200.inference.tshard import ShardedReader, frames_of_rank

CK, CV = 64, 512


def build_args(case, cols, dev, top_k=30):
    """XmAffinityArgs + owned tensors for the given list of memory columns of the single working bank."""
    hw = case['hw']; hw_pad = (hw + 127) // 128 * 128
    b = case['banks'][1]
    n = len(cols)
    cap = max(64, (n + 64 + 7) // 8 * 8)
    idx = torch.tensor(cols, dtype=torch.long)
    rows = torch.zeros(cap, 2 * CK, dtype=torch.float16, device=dev)
    if n:
        lib.key_pack(b['key'][idx].to(dev).contiguous(), rows[:n])
    shr = torch.ones(cap, dtype=torch.float32, device=dev)
    val = torch.zeros(case['n_obj'], CV, cap, dtype=torch.float16, device=dev)
    if n:
        shr[:n] = b['shr'][idx].to(dev)
        val[:, :, :n] = b['val'][:, :, idx].to(dev)
    usage = torch.zeros(cap, dtype=torch.float32, device=dev)
    a = lib.XmAffinityArgs()
    a.banks[0].size = 0; a.banks[2].size = 0
    bk = a.banks[1]
    bk.keys, bk.shrinkage, bk.values, bk.usage = rows.data_ptr(), shr.data_ptr(), val.data_ptr(), usage.data_ptr()
    bk.cap, bk.n_obj_cap, bk.size = cap, case['n_obj'], n
    a.n_groups = 1; a.groups[0].obj_begin, a.groups[0].n_obj = 0, case['n_obj']
    qp, bsq = lib.query_pack(case['qk'].to(dev).contiguous(), case['qe'].to(dev).contiguous(), hw_pad)
    ws = lib.affinity_workspace(hw, case['n_obj'], dev)
    wsb = ws.numel()
    a.qp, a.bsq, a.hw, a.hw_pad, a.top_k, a.n_obj_total = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad, top_k, case['n_obj']
    a.workspace, a.workspace_bytes = ws.data_ptr(), wsb
    return a, (rows, shr, val, usage, qp, bsq, ws)


def run(rank, world, hw, frame, n_frames, n_obj, iters, out_path=None):
    torch.cuda.set_device(rank)
    dev = f'cuda:{rank}'
    case = k1_ref.make_case(hw=hw, sizes=(0, frame * n_frames, 0), n_obj=n_obj, group_begins=[(0, n_obj, [0, 0, 0])], seed=21, device=dev)
    mine = [f * frame + j for f in frames_of_rank(n_frames, rank, world) for j in range(frame)]
    a, keep = build_args(case, mine, dev)
    out = torch.zeros(n_obj, hw, CV, dtype=torch.float16, device=dev)
    reader = ShardedReader()
    out, tau = reader.read(a, out)
    torch.cuda.synchronize()
    res = {'rank': rank}
    if rank == 0:
        # single-GPU read over ALL columns (same kernels, no collectives) as the comparator
        a1, keep1 = build_args(case, list(range(frame * n_frames)), dev)
        ref = torch.zeros(n_obj, hw, CV, dtype=torch.float16, device=dev)
        a1.readout_hwc = ref.data_ptr(); a1.plan_is_resident = 0
        lib.check(lib.load().xm_affinity_readout(C.byref(a1), lib.stream_ptr()), 'xm_affinity_readout')
        tau1 = torch.frombuffer(b'', dtype=torch.float32) if False else None
        torch.cuda.synchronize()
        res['max_abs_diff'] = (out.float() - ref.float()).abs().max().item()
        res['ref_absmax'] = ref.float().abs().max().item()
        exp_out, _, amb, _ = k1_ref.expected(case)
        err = (out.float().cpu().permute(0, 2, 1).double() - exp_out).abs().amax(dim=(0, 1))
        res['oracle_err_clear'] = err[~amb].max().item()
        err1 = (ref.float().cpu().permute(0, 2, 1).double() - exp_out).abs().amax(dim=(0, 1))
        res['oracle_err_clear_1gpu'] = err1[~amb].max().item()
        res['ambiguous'] = int(amb.sum())
        bad = (err1 > 2e-2) & ~amb
        res['bad_queries_1gpu'] = [int(i) for i in torch.nonzero(bad).flatten()[:12]]
        res['n_bad_1gpu'] = int(bad.sum())
    if iters:
        for _ in range(5):                     # NCCL sets its channels up lazily: keep that out of the timed loop
            reader.read(a, out)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            reader.read(a, out)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res['ms_per_read'] = t.item()
        if os.environ.get('TSHARD_HOSTPROF'):
            import time, cProfile, pstats, io
            pr = cProfile.Profile(); pr.enable()
            t0 = time.perf_counter()
            for _ in range(5):
                reader.read(a, out)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            pr.disable()
            if rank == 0:
                sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats('cumulative').print_stats(14)
                print(f'host time per read (no sync): {(t1 - t0) / 5 * 1e3:.2f} ms'); print(sio.getvalue()[:2500], flush=True)
        st = []
        reader.read(a, out, stage_times=st)
        res['stage_ms'] = dict(zip(['stage_a', 'allreduce_max', 'stage_b', 'allgather', 'merge+stage_c', 'allreduce_sum', 'cast'],
                                   [round(x, 4) for x in st]))
        if rank == 0:
            # single-GPU read of the WHOLE memory, timed the same way
            a1, keep1 = build_args(case, list(range(frame * n_frames)), dev)
            ref = torch.zeros(n_obj, hw, CV, dtype=torch.float16, device=dev)
            a1.readout_hwc = ref.data_ptr(); a1.plan_is_resident = 0
            for _ in range(3):
                lib.check(lib.load().xm_affinity_readout(C.byref(a1), lib.stream_ptr()), 'xm_affinity_readout')
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                lib.check(lib.load().xm_affinity_readout(C.byref(a1), lib.stream_ptr()), 'xm_affinity_readout')
            e1.record(); torch.cuda.synchronize()
            res['ms_per_read_1gpu'] = e0.elapsed_time(e1) / iters
    if rank == 0:
        res.update(world=world, hw=hw, N=frame * n_frames, n_obj=n_obj)
        print(json.dumps(res), flush=True)
        if out_path:
            json.dump(res, open(out_path, 'w'))


def _spawned(rank, world, port, hw, frame, n_frames, n_obj, iters, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    try:
        run(rank, world, hw, frame, n_frames, n_obj, iters, out_path)
    finally:
        dist.destroy_process_group()


if __name__ == '__main__':
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    hw = int(sys.argv[1]) if len(sys.argv) > 1 else 8160
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{int(os.environ.get("LOCAL_RANK", 0))}'))
    run(rank, world, hw, hw, frames, 1, 10, os.environ.get('TSHARD_OUT'))
    dist.destroy_process_group()
