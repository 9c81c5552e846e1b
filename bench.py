#!/usr/bin/env python
"""bench.py — XMem++ per-frame memory-attention path on B200 (BASELINE.json metric: 480p frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one synthetic clip: BASELINE.json config 2 — 100 frames of 3x480x854
(ImageNet-normalised U[0,1) noise, Generator(1234+ti)), 1 object, 5 annotated frames {0,20,40,60,80} preloaded into
permanent memory in the driver's CPython-set order, then 100 x InferenceCore.step (working memory grows to 9 frames,
N <= 22 680 memory columns).  Random-init weights of the real architecture (hash-seeded, xmem2_b200.util.synth).

Printed JSON (one line, rank 0):
  value      frames/s with the clip already resident in HBM (whole job, all ranks).  STEADY STATE: recorded CUDA graphs and
             memory arenas are re-used from clip to clip (a first clip also pays 5-230 ms per graph capture).
  e2e        the same metric through the public API with HOST frames: per frame a pinned-host -> device copy of the
             3x480x854 fp32 image, InferenceCore.step, the fused resize+argmax post-process (what `_post_process` of
             run_on_video.py does) and a device -> host copy of the label map into a pinned ring (overlapped with the
             next frame; every frame's mask is on the host before the clock stops)
  roofline   the fused affinity+readout kernel (K1, one launch) at the config-2 memory size, CUDA-event timed in here;
             `traffic` is the DRAM bytes of one launch from the committed ncu capture, stamped with its git revision
  roofline_1080p_shape  the same kernel at the config-4 memory size (HW = 8160, 11 frames)
  conv_roofline  every convolution shape of one 480p frame, graph-timed per shape, summed with its multiplicity
  cpu_baseline  the oracle port (oracle/xmem_oracle.py, plain PyTorch fp32) on this box's host cores, bounded sample;
             its label maps also CHECK this pipeline's output on the same frames (`check`)
  reference_style_gpu  the oracle under CUDA fp16 autocast (cuDNN/cuBLAS, torch.cat memory) = how the reference runs on
             a GPU: two warm clips, then the median of three
With N > 1 every rank runs an independent stream (BASELINE.json config 5; no collective on the data path); the line then
also carries `tshard`: ONE long 1080p video whose memory is sharded over the N ranks (config 4: NCCL all-reduce /
all-gather between the stages of the read) — the sharded read at the config-4 shape against the same read on one GPU,
and a 1080p clip segmented SPMD by all ranks against the single-GPU run of that clip.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

H, W = 480, 854
N_FRAMES = 100
ANNOTATED = [0, 20, 40, 60, 80]
CFG = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=False, hidden_dim=64,
           key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
           max_long_term_elements=10000)
# enable_long_term_count_usage follows run_on_video.py:188-196: 100/(10-5)*128 = 2560 < 10000 -> False


def clip_inputs(seed, n_frames=N_FRAMES, h=H, w=W):
    from xmem2_b200.util.synth import synth_frame, synth_mask
    frames = torch.stack([synth_frame(ti, h, w, seed=seed, structured=False) for ti in range(n_frames)])
    masks = {ti: synth_mask(ti, h, w, 1) for ti in ANNOTATED if ti < n_frames}
    return frames, masks


def run_clip(core_factory, frames, masks, device, host_io, keep=None):
    """One step: preload permanent memory, then the frame loop (run_on_video.py:65-112 without file IO).
    host_io: frames come from pinned host memory and every label map goes back to the host (MaskDownloader).
    keep: optional list that receives (frame index, uint8 label map) of every frame (host_io) or the last probabilities."""
    core = core_factory()
    core.set_all_labels([1])
    for j in list(set(masks.keys())):
        fr = frames[j].to(device, non_blocking=True) if host_io else frames[j]
        core.put_to_permanent_memory(fr, masks[j].to(device, non_blocking=True) if host_io else masks[j])
    n = frames.shape[0]
    dl = up = None
    if host_io:
        from xmem2_b200.inference.pipeline import FrameUploader, MaskDownloader
        dl = MaskDownloader(tuple(frames.shape[-2:]), device)
        up = FrameUploader(device)
        up.prefetch(0, frames[0])
    out = None
    for ti in range(n):
        if up is not None:
            if ti + 1 < n:
                up.prefetch(ti + 1, frames[ti + 1])     # next frame's H2D copy overlaps this frame's kernels
            rgb = up.get(ti)
        else:
            rgb = frames[ti]
        msk = masks.get(ti)
        if msk is not None and host_io:
            msk = msk.to(device, non_blocking=True)
        prob = core.step(rgb, msk, [1] if msk is not None else None, end=(ti == n - 1),
                         do_not_add_mask_to_memory=msk is not None)
        if dl is not None:
            done = dl.submit(ti, prob)              # fused resize + argmax, async D2H into the pinned ring
            if keep is not None:
                keep.extend(done)
        else:
            out = prob
    if dl is not None:
        done = dl.drain()                           # every mask is on the host before the caller stops the clock
        if keep is not None:
            keep.extend(done)
    return out


class ClockSampler:
    """SM clock + throttle reasons sampled in-process through NVML every 100 ms during the timed region."""

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None
        self.max_mhz = None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self._stop.is_set():
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((mhz, r))
                self._stop.wait(0.1)
        except Exception as e:          # never let the sampler break the benchmark
            self.error = str(e)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True); self._t.start(); return self

    def __exit__(self, *a):
        self._stop.set(); self._t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'note': getattr(self, 'error', 'no samples')}
        sm = sorted(s[0] for s in self.samples)
        bits = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
        reasons = sorted({name for _, r in self.samples for bit, name in bits.items() if r & bit})
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': reasons, 'samples': len(sm)}


def k1_roofline(device, hw=1620, frames=(9, 5)):
    """CUDA-event timing of xm_affinity_readout (ONE k1_fused launch) at a full memory: `frames` = (working, permanent) frames of
    `hw` columns each — the config-2 maximum by default (HW=1620, 9 + 5 frames).  An L2-sized buffer is rewritten between launches."""
    from xmem2_b200.util import synth_memory as sm
    nw, npm = frames[0] * hw, frames[1] * hw
    case = sm.make_case(hw=hw, sizes=(0, nw, npm), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=11, device=device)
    a, keep = sm.device_args(case)
    out = torch.empty(1, hw, 512, dtype=torch.float16, device=device)
    a.readout_hwc = out.data_ptr()
    t = sm.time_readout(a, iters=10, warmup=3)
    N = nw + npm
    flops = 4 * 64 * N * hw + 2 * 512 * N * hw * 1                    # F_K1
    bytes_ = N * (2 * 64 * 2 + 4) + 512 * N * 2 + 2 * 64 * hw * 2 + 512 * hw * 2 + N * 4   # B_K1 (packed keys are 256 B/column)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('bf16_tflops', 1590.0)
    traffic, traffic_src = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(REPO, 'profiles', 'r2_k1_traffic.json')))
        if hw == 1620:
            traffic, traffic_src = tj['dram_bytes_per_launch'], tj.get('captured_at')
    except Exception:
        pass
    return {'bound': 'tensor', 'achieved': round(flops / t / 1e12, 2), 'peak': peak, 'unit': 'TFLOP/s',
            'frac': round(flops / t / 1e12 / peak, 4), 'traffic': traffic, 'traffic_from': traffic_src,
            'kernel': 'k1_fused (xm_affinity_readout: one persistent kernel, sweeps + selection + readout + reduce)',
            'launch_us': round(t * 1e6, 1), 'shape': {'N': N, 'HW': hw, 'n_obj': 1},
            'peak_source': 'MEASURED_PEAKS.json bf16 burst' if peaks else 'fallback', 'algorithmic_bytes': bytes_,
            'hbm_gbs_if_bytes_bound': round(bytes_ / t / 1e9, 1)}


def conv_roofline(device):
    """The convolution family (K2+K4+K5 GEMM-shaped work): every conv shape of one ordinary 480p frame, graph-replayed and
    CUDA-event timed per shape (xmem2_b200/util/conv_bench.py), summed with the per-frame multiplicities."""
    from xmem2_b200.util.conv_bench import conv_table
    rows, us, fl = conv_table(device)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('bf16_tflops', 1590.0)
    top = sorted(rows, key=lambda r: -r['us'] * r['per_frame'])[:6]
    return {'bound': 'tensor', 'achieved': round(fl / us / 1e6, 1), 'peak': peak, 'unit': 'TFLOP/s', 'frac': round(fl / us / 1e6 / peak, 4),
            'frame_us': round(us, 1), 'frame_gflop': round(fl / 1e9, 1), 'launches': sum(r['per_frame'] for r in rows),
            'kernels': 'conv_igemm_csk_kernel (cluster split-K) / conv_igemm_2cta_kernel (CTA pairs, BN=256) / conv3x3_c1_kernel', 'top_shapes': top}


def oracle_clip(device, n_frames, threads=None, autocast=False, seed=1234, keep=None):
    """seconds of the oracle port over the first n_frames of the config-2 clip; keep: list receiving uint8 label maps."""
    from oracle import xmem_oracle as O
    from xmem2_b200.util.synth import synth_state_dict
    if threads:
        torch.set_num_threads(threads)
    state = {k: v.to(device) for k, v in synth_state_dict(0).items()}
    frames, masks = clip_inputs(seed, n_frames)
    frames = frames.to(device); masks = {k: v.to(device) for k, v in masks.items()}
    ctx = torch.autocast('cuda', dtype=torch.float16) if autocast else torch.autocast('cpu', enabled=False)
    with ctx:
        net = O.OracleNet(state)
        core = O.OracleCore(net, dict(CFG))
        if device != 'cpu':
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        core.set_all_labels([1])
        for j in list(set(masks.keys())):
            core.put_to_permanent_memory(frames[j], masks[j])
        for ti in range(n_frames):
            msk = masks.get(ti)
            prob = core.step(frames[ti], msk, [1] if msk is not None else None, end=(ti == n_frames - 1),
                             do_not_add_mask_to_memory=msk is not None)
            if keep is not None:
                keep.append(torch.argmax(prob, dim=0).to(torch.uint8).cpu())
        if device != 'cpu':
            torch.cuda.synchronize()
        return time.perf_counter() - t0


def cpu_sample_frames(cores):
    """How many frames of the clip one CPU step covers: the whole clip when the host does it in ~25 s, else a ~20 s prefix
    (never fewer than 22 frames, so that the memory holds at least two working frames and the five preloads)."""
    t = oracle_clip('cpu', 4, threads=cores)
    fps = 4 / t
    return N_FRAMES if N_FRAMES / fps <= 25 else max(22, min(N_FRAMES, int(20 * fps)))


def tshard_measure(net, rank, world, local, device):
    """BASELINE.json config 4 on `world` GPUs: (a) the T-sharded read at the 1080p shape (HW = 8160, 11 stored frames = 89 760
    memory columns) against the same read on one GPU; (b) a 1080p clip segmented SPMD by all ranks (`t_shard=True`: memory
    sharded by stored frame, NCCL collectives between the stages of every read) against the single-GPU run of that clip."""
    import ctypes as C
    import torch.distributed as dist
    from xmem2_b200 import lib
    from xmem2_b200.inference.inference_core import InferenceCore
    from xmem2_b200.inference.tshard import ShardedReader, frames_of_rank
    from xmem2_b200.util import synth_memory as sm
    from xmem2_b200.util import dist as xd
    res = {}
    # ---- (a) the read
    hw, n_fr = 8160, 11
    case = sm.make_case(hw=hw, sizes=(0, hw * n_fr, 0), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=21, device=device)
    mine = [f * hw + j for f in frames_of_rank(n_fr, rank, world) for j in range(hw)]
    a, keep = sm.device_args(case, columns={1: mine})
    out = torch.zeros(1, hw, 512, dtype=torch.float16, device=device)
    reader = ShardedReader()
    for _ in range(5):                     # NCCL sets its channels up lazily
        reader.read(a, out)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        reader.read(a, out)
    e1.record(); torch.cuda.synchronize()
    res['ms_per_read'] = round(xd.max_over_ranks(e0.elapsed_time(e1) / iters, device), 4)
    st = []
    reader.read(a, out, stage_times=st)
    res['stage_ms'] = dict(zip(['stage_a', 'allreduce_max', 'stage_b', 'allgather', 'merge+stage_c', 'allreduce_sum', 'cast'],
                               [round(x, 4) for x in st]))
    res['read_shape'] = {'HW': hw, 'N': hw * n_fr, 'n_obj': 1}
    if rank == 0:
        a1, keep1 = sm.device_args(case)
        ref = torch.zeros(1, hw, 512, dtype=torch.float16, device=device)
        a1.readout_hwc = ref.data_ptr()
        res['ms_per_read_1gpu'] = round(sm.time_readout(a1, iters=10, warmup=3, flush_l2=False) * 1e3, 4)
        d = (out.float() - ref.float()).abs()
        # queries with an exact fp32 tie at rank top_k keep all tied columns when sharded and exactly k on one GPU
        res['read_max_abs_diff_vs_1gpu'] = round(float(d.max()), 5)
        res['read_p999_abs_diff_vs_1gpu'] = round(float(torch.quantile(d.flatten()[::7].float(), 0.999)), 6)
        del a1, keep1, ref
    del a, keep, out, case
    torch.cuda.empty_cache()
    dist.barrier()
    # ---- (b) a 1080p clip, every rank runs every frame, the memory is sharded
    hh, ww, n_frames = 1080, 1920, 56
    cfg = dict(CFG); cfg.update(mem_every=4, max_mid_term_frames=6, min_mid_term_frames=3, enable_long_term_count_usage=True)
    from xmem2_b200.util.synth import synth_frame, synth_mask
    frames = [synth_frame(ti, hh, ww, seed=77, structured=True).to(device) for ti in range(n_frames)]
    m0 = synth_mask(0, hh, ww, 1).to(device)

    def run(shard):
        c = dict(cfg); c['t_shard'] = shard
        core = InferenceCore(net, c)
        core.set_all_labels([1])
        core.put_to_permanent_memory(frames[0], m0.clone())
        probs = []
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for ti in range(n_frames):
            msk = m0.clone() if ti == 0 else None
            p = core.step(frames[ti], msk, [1] if msk is not None else None, end=(ti == n_frames - 1),
                          do_not_add_mask_to_memory=msk is not None)
            if ti % 11 == 10:
                probs.append(p.clone())
        ev1.record(); torch.cuda.synchronize()
        return ev0.elapsed_time(ev1), probs, core

    run(True)                                                   # warm-up (kernel attributes, NCCL channels)
    dist.barrier()
    ms, probs_sh, core = run(True)
    ms = xd.max_over_ranks(ms, device)
    res['clip'] = {'frames': n_frames, 'size': [hh, ww], 'fps': round(n_frames / (ms * 1e-3), 2), 'ms_per_frame': round(ms / n_frames, 3),
                   'global_columns_end': int(core.memory.global_temp_size + len(core.memory._perm_frames) * core.memory.HW + core.memory.global_long_size),
                   'long_term_blocks': int(core.memory._blocks['long'])}
    del core
    if rank == 0:
        run(False)
        ms1, probs_1, _ = run(False)
        res['clip']['fps_1gpu_unsharded'] = round(n_frames / (ms1 * 1e-3), 2)
        res['clip']['max_abs_dprob_vs_1gpu'] = round(max(float((x - y).abs().max()) for x, y in zip(probs_sh, probs_1)), 5)
        res['clip']['mean_abs_dprob_vs_1gpu'] = round(max(float((x - y).abs().mean()) for x, y in zip(probs_sh, probs_1)), 7)
    dist.barrier()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
    cores = min(os.cpu_count() or 1, 32)      # more threads only slow oneDNN down on these small convolutions
    config = {'workload': 'BASELINE.json config 2: synthetic 480p (3x480x854) 100-frame clip, 1 object, 5 permanent-memory '
                          'masks {0,20,40,60,80}, mem_every=10, top_k=30, full InferenceCore.step pipeline',
              'frames_per_step': N_FRAMES, 'streams': world, 'l2': 'inputs larger than L2 (492 MB of frames per step)',
              'parallelism': f'{world} independent stream(s), one per GPU, no collective',
              'state': 'steady state: recorded CUDA graphs and memory arenas are re-used across clips (warm-up clips record them)'}

    if args.impl == 'reference':
        if rank != 0:
            return
        nfr = cpu_sample_frames(cores)
        vals = []
        for i in range(args.warmup + args.steps):
            t = oracle_clip('cpu', nfr, threads=cores)
            if i >= args.warmup:
                vals.append(nfr / t)
        fps = statistics.median(vals)
        line = {'impl': 'reference', 'metric': 'fps_480p', 'value': round(fps, 4), 'unit': 'frames/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(1000 * nfr / fps, 2), 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': round(fps, 4), 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                 'sample': f'first {nfr} of the 100 frames of the config-2 clip per step (5 annotated frames preloaded; '
                                           f'the whole clip when the host manages it in ~25 s), median of {len(vals)} steps, '
                                           'oracle/xmem_oracle.py (PyTorch fp32 restatement of the reference) on all host threads; '
                                           'one CPU process regardless of --gpus',
                                 'steps_fps': [round(v, 3) for v in vals]},
                'e2e': {'value': round(fps, 4), 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    torch.cuda.set_device(local)
    device = f'cuda:{local}'
    from xmem2_b200 import lib
    from xmem2_b200.util import dist as xd
    from xmem2_b200.inference.inference_core import InferenceCore
    from xmem2_b200.model.network import XMem
    from xmem2_b200.util.synth import synth_state_dict
    lib.load()
    net = XMem(dict(CFG), None).to(device).eval()
    net.load_weights(synth_state_dict(0))
    frames_h, masks_h = clip_inputs(xd.stream_seed(1234, rank))
    frames_pin = frames_h.pin_memory(); masks_pin = {k: v.pin_memory() for k, v in masks_h.items()}
    frames_d = frames_h.to(device); masks_d = {k: v.to(device) for k, v in masks_h.items()}
    factory = lambda: InferenceCore(net, dict(CFG))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host_io, steps, warmup):
        fr, mk = (frames_pin, masks_pin) if host_io else (frames_d, masks_d)
        for _ in range(warmup):
            run_clip(factory, fr, mk, device, host_io)
        barrier()
        l0 = lib.load().xm_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run_clip(factory, fr, mk, device, host_io)
        e1.record()
        barrier()
        ms = xd.max_over_ranks(e0.elapsed_time(e1), device)          # the slowest rank's device time
        launches = lib.load().xm_launch_count() - l0
        return ms, launches

    with ClockSampler(local) as clk:
        ms_dev, launches = timed(False, args.steps, args.warmup)
    clocks = clk.summary()
    ms_e2e, _ = timed(True, args.steps, max(1, min(args.warmup, 2)))
    total_frames = N_FRAMES * args.steps * world
    value = total_frames / (ms_dev * 1e-3)
    e2e = total_frames / (ms_e2e * 1e-3)
    line = {'metric': 'fps_480p', 'value': round(value, 2), 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(ms_dev / args.steps, 2), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic', 'config': config,
            'e2e': {'value': round(e2e, 2), 'unit': 'frames/s', 'h2d_bytes_per_step': N_FRAMES * 3 * H * W * 4 + len(ANNOTATED) * 2 * H * W * 4
                    + len(ANNOTATED) * 3 * H * W * 4, 'd2h_bytes_per_step': N_FRAMES * H * W,
                    'what': 'pinned frame -> device (side stream, one frame ahead), InferenceCore.step, fused resize+argmax (xm_resize_argmax), label map -> pinned ring '
                            '(overlapped; all masks on the host before the clock stops)'},
            'gpu_launches': int(launches), 'clocks': clocks}
    if world > 1:
        try:
            ts = tshard_measure(net, rank, world, local, device)
            if rank == 0:
                line['tshard'] = ts
        except Exception as e:
            if rank == 0:
                line['tshard'] = {'error': str(e)[:300]}
    if rank == 0:
        if world == 1:
            try:
                line['roofline'] = k1_roofline(device)
            except Exception as e:                       # never lose the headline number to the side measurement
                line['roofline'] = {'error': str(e)[:200]}
            try:   # the same kernel at the 1080p memory size (config 4: HW=8160, 6 working + 5 permanent frames)
                r4 = k1_roofline(device, hw=8160, frames=(6, 5))
                line['roofline_1080p_shape'] = {k: r4[k] for k in ('achieved', 'peak', 'unit', 'frac', 'launch_us', 'shape', 'algorithmic_bytes')}
                line['conv_roofline'] = conv_roofline(device)
            except Exception as e:
                line['roofline_1080p_shape'] = {'error': str(e)[:200]}
            try:
                nfr = cpu_sample_frames(cores)
                ref_masks = []
                t_cpu = oracle_clip('cpu', nfr, threads=cores, keep=ref_masks)
                line['cpu_baseline'] = {'value': round(nfr / t_cpu, 4), 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                        'sample': f'first {nfr} of the 100 frames of the config-2 clip (5 annotated frames preloaded), '
                                                  'oracle/xmem_oracle.py fp32 on all host threads, one run'}
                # the same frames through this pipeline (host frames in, label maps out): output check against the oracle
                mine = []
                run_clip(factory, frames_pin[:nfr], {k: v for k, v in masks_pin.items() if k < nfr}, device, True, keep=mine)
                agree = [float((torch.from_numpy(m) == ref_masks[ti]).float().mean()) for ti, m in sorted(mine)]
                line['check'] = {'frames': len(agree), 'argmax_agreement_mean': round(sum(agree) / len(agree), 5),
                                 'argmax_agreement_min': round(min(agree), 5),
                                 'against': 'fp32 CPU oracle label maps of the same frames (fp16 tensor-core pipeline vs fp32: see '
                                            'tests/test_gpu_baseline_shapes.py for the calibrated bars)'}
                assert min(agree) > 0.98, f'bench output check failed: {line["check"]}'
            except AssertionError:
                raise
            except Exception as e:
                line['cpu_baseline'] = {'error': str(e)[:200]}
            try:
                for _ in range(2):
                    oracle_clip(device, N_FRAMES, autocast=True)
                vals = [N_FRAMES / oracle_clip(device, N_FRAMES, autocast=True) for _ in range(3)]
                line['reference_style_gpu'] = {'value': round(statistics.median(vals), 2), 'unit': 'frames/s', 'runs': [round(v, 1) for v in vals],
                                               'what': 'oracle port on this GPU under torch fp16 autocast (cuDNN/cuBLAS library kernels, '
                                                       'torch.cat memory, materialised affinity) = how the reference runs on a GPU; '
                                                       'two warm clips, then the median of three'}
                line['speedup_vs_reference_style_gpu'] = {'device_resident': round(value / statistics.median(vals), 2),
                                                          'e2e': round(e2e / statistics.median(vals), 2)}
            except Exception as e:
                line['reference_style_gpu'] = {'error': str(e)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
