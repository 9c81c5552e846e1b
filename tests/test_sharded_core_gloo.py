"""world_size-2 gloo test (CPU) of ONE video segmented SPMD with a T-sharded memory:
`InferenceCore` with `t_shard=True` on every rank against the ordinary single-process `InferenceCore`, same frames.

The network is a deterministic stand-in whose segmentation DEPENDS on the memory readout, the stores' CUDA entry points
are replaced by CPU doubles, and the read kernels by an oracle restatement of the read (`oracle.xmem_oracle.similarity` /
`softmax_topk`) — for the sharded cores over the shards gathered back into global column order, for the single-process
core over its own banks.  What is verified is the host logic between them: frame ownership of working / permanent /
long-term blocks, the usage bookkeeping on the owner, consolidation and eviction across ranks, hidden-state and mask
handling — the predicted probabilities, hidden state and bank contents of every rank must equal the single-process run
after every frame."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

H, W = 32, 48
h, w = H // 16, W // 16
HW = h * w


class ReadoutNet(torch.nn.Module):
    """XMem call surface; keys move with the frame, values with the mask, the segmentation with the readout."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def encode_key(self, image, need_sk=True, need_ek=True):
        m = image.mean(dim=(1, 2, 3)).view(1, 1, 1, 1)
        base = torch.sin(torch.arange(64 * h * w, dtype=torch.float32) * 0.37).view(1, 64, h, w)
        key = (base * (0.5 + m) + m).half().float()
        shr = torch.ones(1, 1, h, w) + m
        sel = (torch.full((1, 64, h, w), 0.5) + 0.1 * torch.cos(base)).half().float()
        f16 = torch.zeros(1, 8, h, w) + m
        return key, shr if need_sk else None, sel if need_ek else None, f16, f16, f16

    def segment(self, feats, readout, hidden, selector=None, h_out=True, strip_bg=True):
        n = readout.shape[1]
        r = readout.float().mean(dim=(2, 3, 4)).view(1, n, 1, 1)          # the memory readout steers the logits
        yy = torch.linspace(-2, 2, H).view(1, 1, H, 1); xx = torch.linspace(-2, 2, W).view(1, 1, 1, W)
        logits = yy + xx * torch.arange(1, n + 1).view(1, n, 1, 1) + feats[0].mean() + 1.5 * (r - 2.0)
        prob = torch.sigmoid(logits)
        bg = torch.prod(1 - prob, dim=1, keepdim=True)
        p = torch.cat([bg, prob], 1).clamp(1e-7, 1 - 1e-7)
        p = torch.softmax(torch.log(p / (1 - p)), dim=1)
        new_h = hidden + r.view(1, n, 1, 1, 1) if h_out else None
        return new_h, None, (p[:, 1:] if strip_bg else p)

    def encode_value(self, frame, f16, h16, masks, is_deep_update=True):
        n = masks.shape[1]
        ramp = torch.linspace(-1, 1, 512).view(1, 1, 512, 1, 1)
        val = (masks.mean(dim=(2, 3)).view(1, n, 1, 1, 1) * 4 + ramp * f16.mean() + torch.zeros(1, n, 512, h, w)).half().float()
        return val, (h16 * 0.5 + 1.0 if is_deep_update else h16)


def _install_cpu_doubles():
    from xmem2_b200 import lib
    from xmem2_b200.inference import kv_memory_store as kv

    from tests import cpu_doubles
    lib.require_cuda = lambda t, name: None
    for name in ('key_pack', 'usage_topk', 'usage_evict_list', 'consolidate_affinity', 'consolidate_values', 'bank_compact'):
        setattr(lib, name, getattr(cpu_doubles, name))
    kv._ARENA_POOL.clear()


def _cfg(**over):
    cfg = dict(mem_every=2, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True, hidden_dim=64,
               key_dim=64, value_dim=512, top_k=5, max_mid_term_frames=5, min_mid_term_frames=2, num_prototypes=4,
               max_long_term_elements=14, use_cuda_graph=False)
    cfg.update(over)
    return cfg


def _oracle_read(banks, qk, qe, top_k):
    """banks: list of (k [1,64,N], s [1,1,N], v [n_obj,512,N]) in the reference's order (long, working, permanent)."""
    from oracle import xmem_oracle as O
    mk = torch.cat([b[0].float() for b in banks], -1)
    ms = torch.cat([b[1].float() for b in banks], -1)
    mv = torch.cat([b[2].float() for b in banks], -1)
    sim = O.similarity(mk, ms, qk.float(), qe.float())
    aff, usage = O.softmax_topk(sim, top_k, want_usage=True)
    return mv @ aff[0], usage.flatten()


def _single_read(mm):
    def read(qk, qe, disable_usage_updates=False):
        banks, sizes = [], []
        for st in (mm.long_mem, mm.temporary_work_mem, mm.permanent_work_mem):
            if st.engaged() and st.size:
                banks.append((st.k, st.s, st.v[0])); sizes.append((st, st.size))
        out, usage = _oracle_read(banks, qk.flatten(2), qe.flatten(2), mm.top_k)
        if not disable_usage_updates:
            off = 0
            for st, n in sizes:
                if st.count_usage:
                    st.update_usage(usage[off:off + n])
                off += n
        return out.view(out.shape[0], 512, h, w)
    return read


def _sharded_read(mm, world):
    def read(qk, qe, disable_usage_updates=False):
        banks, mine = [], []
        for name, st in (('long', mm.long_mem), ('temp', mm.temporary_work_mem), ('perm', mm.permanent_work_mem)):
            pos = mm.local_positions(name)
            payload = (pos, st.k.clone(), st.s.clone(), st.v[0].clone()) if (st.engaged() and st.size) else (pos, None, None, None)
            parts = [None] * world
            dist.all_gather_object(parts, payload)
            total = sum(p[0].numel() for p in parts)
            if total == 0:
                continue
            n_obj = next(p[3].shape[0] for p in parts if p[3] is not None)
            k = torch.zeros(1, 64, total, dtype=torch.float16); s = torch.zeros(1, 1, total); v = torch.zeros(n_obj, 512, total, dtype=torch.float16)
            for p in parts:
                if p[1] is not None:
                    k[:, :, p[0]] = p[1]; s[:, :, p[0]] = p[2]; v[:, :, p[0]] = p[3]
            banks.append((k, s, v)); mine.append((st, pos, total))
        out, usage = _oracle_read(banks, qk.flatten(2), qe.flatten(2), mm.top_k)
        if not disable_usage_updates:
            off = 0
            for st, pos, total in mine:
                if st.count_usage and pos.numel():
                    st.update_usage(usage[off:off + total][pos])      # the owner of a column keeps its statistics
                elif st.count_usage:
                    st.update_usage(usage[0:0])
                off += total
        return out.view(out.shape[0], 512, h, w)
    return read


def _masks(ti, n):
    yy = torch.arange(H).view(H, 1); xx = torch.arange(W).view(1, W)
    out = torch.zeros(n, H, W)
    for o in range(n):
        out[o] = ((yy - 8 - 2 * o - ti) ** 2 + (xx - 12 - 9 * o) ** 2 <= 36).float()
    return out


def _worker(rank, world, port, out):
    try:
        os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        torch.manual_seed(0)
        _install_cpu_doubles()
        from xmem2_b200.inference.inference_core import InferenceCore
        from xmem2_b200.inference.sharded_memory import ShardedMemoryManager
        sharded = InferenceCore(ReadoutNet(), _cfg(t_shard=True))
        single = InferenceCore(ReadoutNet(), _cfg())
        assert isinstance(sharded.memory, ShardedMemoryManager) and not sharded.use_cuda_graph
        for core in (sharded, single):
            core.set_all_labels([1])
        sharded.memory.match_memory = _sharded_read(sharded.memory, world)
        single.memory.match_memory = _single_read(single.memory)
        # two annotated frames go to the permanent bank (blocks 0 and 1 -> ranks 0 and 1)
        for ti in (0, 1):
            img = torch.ones(3, H, W) * (0.1 + 0.04 * ti)
            for core in (sharded, single):
                core.put_to_permanent_memory(img, _masks(ti, 1))
        stats = dict(long_blocks=0, max_err=0.0)
        for ti in range(44):
            img = torch.ones(3, H, W) * (0.1 + 0.03 * ti) + 0.02 * torch.sin(torch.arange(W).float() * (ti + 1)).view(1, 1, W)
            msk = _masks(ti, 1) if ti == 0 else None
            kw = dict(end=(ti == 43))
            a = sharded.step(img, msk.clone() if msk is not None else None, [1] if msk is not None else None, **kw)
            b = single.step(img, msk.clone() if msk is not None else None, [1] if msk is not None else None, **kw)
            err = (a.float() - b.float()).abs().max().item()
            stats['max_err'] = max(stats['max_err'], err)
            assert err < 1e-6, (ti, err)
            ha, hb = sharded.memory.get_hidden(), single.memory.get_hidden()
            assert torch.allclose(ha.float(), hb.float(), atol=1e-6), ti
            sm, rm = sharded.memory, single.memory
            assert sm.global_temp_size == rm.temporary_work_mem.size, ti
            assert sm.global_long_size == rm.long_mem.size, ti
            for name, mine, theirs in (('temp', sm.temporary_work_mem, rm.temporary_work_mem),
                                       ('perm', sm.permanent_work_mem, rm.permanent_work_mem), ('long', sm.long_mem, rm.long_mem)):
                pos = sm.local_positions(name)
                assert pos.numel() == mine.size, (ti, name)
                if mine.size:
                    assert torch.equal(mine.k, theirs.k[..., pos]), (ti, name, 'k')
                    assert torch.equal(mine.v[0], theirs.v[0][..., pos]), (ti, name, 'v')
                    if mine.count_usage:
                        assert torch.allclose(mine.get_usage(), theirs.get_usage()[..., pos], atol=1e-6), (ti, name, 'usage')
            stats['long_blocks'] = sm._blocks['long']
        assert stats['long_blocks'] >= 6, stats           # consolidated several times, with eviction (max 14 long columns)
        assert single.memory.long_mem.size <= 14
        out.put((rank, 'ok', stats))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        import traceback
        out.put((rank, 'error', traceback.format_exc()))
        raise


def test_sharded_video_equals_single_process_video():
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in res:
        assert status == 'ok', info
