"""Worker for tests/test_gpu_zz_tshard_clip.py: ONE golden clip segmented SPMD by `world` ranks with a T-sharded memory
(`InferenceCore(config t_shard=True)`, NCCL), checked on every rank against the trace of the live reference
(tests/golden/clip_*.npz): global bank sizes per frame (exact), probabilities, argmax agreement."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
BASE = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True, hidden_dim=64,
            key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
            max_long_term_elements=10000)


def run(rank, world, name, out_path=None):
    from xmem2_b200.inference.inference_core import InferenceCore
    from xmem2_b200.model.network import XMem
    from xmem2_b200.util.synth import synth_state_dict, synth_frame, synth_mask
    torch.set_grad_enabled(False)
    torch.cuda.set_device(rank)
    dev = f'cuda:{rank}'
    net = XMem({}, None).to(dev).eval()
    net.load_weights(synth_state_dict(0))
    d = np.load(os.path.join(G, f'clip_{name}.npz'))
    H, W, n_frames, n_obj, save_every = [int(x) for x in d['hw']]
    cfg = dict(BASE); cfg.update({str(k): int(v) for k, v in zip(d['cfg_keys'], d['cfg_vals'])})
    cfg['t_shard'] = True
    ffo = [int(x) for x in d['first_frame_of']]; annotated = [int(x) for x in d['annotated']]
    core = InferenceCore(net, cfg)
    n_seen = 0
    for j in [int(x) for x in d['order']]:
        n_seen = max(n_seen, sum(1 for f in ffo if f <= j))
        core.set_all_labels(list(range(1, n_seen + 1)))
        core.put_to_permanent_memory(synth_frame(j, H, W, structured=True).to(dev), synth_mask(j, H, W, n_obj, ffo)[:n_seen].to(dev))
    labels = list(range(1, n_seen + 1))
    k, worst_mean, worst_agree, size_mismatch = 0, 0.0, 1.0, []
    local_cols = []
    for ti in range(n_frames):
        msk = synth_mask(ti, H, W, n_obj, ffo).to(dev) if ti in annotated else None
        p = core.step(synth_frame(ti, H, W, structured=True).to(dev), msk, labels if msk is not None else None,
                      end=(ti == n_frames - 1), do_not_add_mask_to_memory=msk is not None)
        m = core.memory
        sizes = [m.global_temp_size, len(m._perm_frames) * m.HW, m.global_long_size]
        if sizes != d['sizes'][ti][:3].tolist():
            size_mismatch.append((ti, sizes, d['sizes'][ti][:3].tolist()))
        local_cols.append(m.temporary_work_mem.size + m.permanent_work_mem.size + m.long_mem.size)
        if ti % save_every == 0:
            ref = torch.from_numpy(d['probs'][k]).float(); k += 1
            e = (p.float().cpu() - ref).abs()
            worst_mean = max(worst_mean, e.mean().item())
            worst_agree = min(worst_agree, (p.argmax(0).cpu() == ref.argmax(0)).float().mean().item())
    # every rank holds the same prediction: compare the last probability map across ranks
    last = p.float().contiguous()
    peers = [torch.empty_like(last) for _ in range(world)]
    dist.all_gather(peers, last)
    spread = max((q - last).abs().max().item() for q in peers)
    res = dict(rank=rank, world=world, clip=name, worst_mean=worst_mean, worst_agree=worst_agree, size_mismatch=size_mismatch[:3],
               rank_spread=spread, max_local_cols=max(local_cols), long_blocks=core.memory._blocks['long'])
    if out_path:
        json.dump(res, open(f'{out_path}.{rank}', 'w'))
    return res


def _spawned(rank, world, port, name, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(f'cuda:{rank}'))
    try:
        run(rank, world, name, out_path)
    finally:
        dist.destroy_process_group()


if __name__ == '__main__':          # torchrun --nproc-per-node 2 tests/tshard_clip_worker.py one_obj
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{int(os.environ.get("LOCAL_RANK", 0))}'))
    print(json.dumps(run(rank, world, sys.argv[1] if len(sys.argv) > 1 else 'one_obj')), flush=True)
    dist.destroy_process_group()
