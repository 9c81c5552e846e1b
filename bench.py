#!/usr/bin/env python
"""bench.py — XMem++ per-frame memory-attention path on B200 (BASELINE.json metric: 480p frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one synthetic clip: BASELINE.json config 2 — 100 frames of 3x480x854
(ImageNet-normalised U[0,1) noise, Generator(1234+ti)), 1 object, 5 annotated frames {0,20,40,60,80} preloaded into
permanent memory in the driver's CPython-set order, then 100 x InferenceCore.step (working memory grows to 9 frames,
N <= 22 680 memory columns).  Random-init weights of the real architecture (hash-seeded, xmem2_b200.util.synth).

Printed JSON (one line, rank 0):
  value      frames/s with the clip already resident in HBM (whole job, all ranks)
  e2e        the same metric through the public API with HOST frames: per frame a pinned-host -> device copy of the
             3x480x854 fp32 image and a device -> host read of the argmax label map (what run_on_video.py does)
  roofline   the fused affinity+readout kernel group (K1) at the config-2 memory size, CUDA-event timed in here
  cpu_baseline  the oracle port (oracle/xmem_oracle.py, plain PyTorch fp32) on this box's host cores, bounded sample
With N > 1 every rank runs an independent stream (BASELINE.json config 5; no collective on the data path).
"""
import argparse
import json
import os
import subprocess
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


def clip_inputs(seed, n_frames=N_FRAMES):
    from xmem2_b200.util.synth import synth_frame, synth_mask
    frames = torch.stack([synth_frame(ti, H, W, seed=seed, structured=False) for ti in range(n_frames)])
    masks = {ti: synth_mask(ti, H, W, 1) for ti in ANNOTATED if ti < n_frames}
    return frames, masks


def run_clip(core_factory, frames, masks, device, host_io):
    """One step: preload permanent memory, then the frame loop (run_on_video.py:65-112 without file IO)."""
    core = core_factory()
    core.set_all_labels([1])
    for j in list(set(masks.keys())):
        fr = frames[j].to(device, non_blocking=True) if host_io else frames[j]
        core.put_to_permanent_memory(fr, masks[j].to(device, non_blocking=True) if host_io else masks[j])
    n = frames.shape[0]
    out = None
    for ti in range(n):
        rgb = frames[ti].to(device, non_blocking=True) if host_io else frames[ti]
        msk = masks.get(ti)
        if msk is not None and host_io:
            msk = msk.to(device, non_blocking=True)
        prob = core.step(rgb, msk, [1] if msk is not None else None, end=(ti == n - 1),
                         do_not_add_mask_to_memory=msk is not None)
        if host_io:
            out = torch.argmax(prob, dim=0).to(torch.uint8).cpu()      # D2H + sync, as _post_process (run_on_video.py:171-172)
        else:
            out = prob
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


def k1_roofline(device):
    """CUDA-event timing of xm_affinity_readout at the config-2 maximum (HW=1620, 9 working + 5 permanent frames)."""
    import ctypes as C
    from xmem2_b200 import lib
    from tests import k1_ref
    hw, nw, npm = 1620, 9 * 1620, 5 * 1620
    case = k1_ref.make_case(hw=hw, sizes=(0, nw, npm), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=11, device=device)
    # build the device-side state once (same code path as the parity tests), then time repeated calls
    L = lib.load()
    hw_pad = (hw + 127) // 128 * 128
    a = lib.XmAffinityArgs(); keep = []
    for bi, b in enumerate(case['banks']):
        if b is None:
            a.banks[bi].size = 0; continue
        rows = torch.zeros(b['cap'], 128, dtype=torch.float16, device=device)
        lib.key_pack(b['key'].to(device).contiguous(), rows[:b['n']])
        shr = torch.ones(b['cap'], dtype=torch.float32, device=device); shr[:b['n']] = b['shr'].to(device)
        val = b['val'].to(device).contiguous(); usage = torch.zeros(b['cap'], dtype=torch.float32, device=device)
        keep += [rows, shr, val, usage]
        bk = a.banks[bi]
        bk.keys, bk.shrinkage, bk.values, bk.usage = rows.data_ptr(), shr.data_ptr(), val.data_ptr(), usage.data_ptr()
        bk.cap, bk.n_obj_cap, bk.size = b['cap'], 1, b['n']
    a.n_groups = 1; a.groups[0].obj_begin, a.groups[0].n_obj = 0, 1
    qp, bsq = lib.query_pack(case['qk'].to(device).contiguous(), case['qe'].to(device).contiguous(), hw_pad)
    ws = lib.affinity_workspace(hw, 1, device); wsb = ws.numel()
    out = torch.empty(1, hw, 512, dtype=torch.float16, device=device)
    a.qp, a.bsq, a.hw, a.hw_pad, a.top_k, a.n_obj_total = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad, 30, 1
    a.readout_hwc, a.workspace, a.workspace_bytes = out.data_ptr(), ws.data_ptr(), wsb
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)      # > 126 MB L2
    times = []
    for it in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(L.xm_affinity_readout(C.byref(a), lib.stream_ptr()), 'xm_affinity_readout')
        e1.record(); torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1) * 1e-3)
    t = sum(times) / len(times)
    N = nw + npm
    flops = 4 * 64 * N * hw + 2 * 512 * N * hw * 1                    # F_K1
    bytes_ = N * (2 * 64 * 2 + 4) + 512 * N * 2 + 2 * 64 * hw * 2 + 512 * hw * 2 + N * 4   # B_K1 (packed keys are 256 B/column)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('bf16_tflops', 1590.0)
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the six kernels of one call, from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(REPO, 'profiles', 'r1_k1_traffic.json')))['dram_bytes_per_call']
    except Exception:
        pass
    return {'bound': 'tensor', 'achieved': round(flops / t / 1e12, 2), 'peak': peak, 'unit': 'TFLOP/s',
            'frac': round(flops / t / 1e12 / peak, 4), 'traffic': traffic, 'kernel': 'xm_affinity_readout (pass1+merge+pass2+finish)',
            'launch_us': round(t * 1e6, 1), 'shape': {'N': N, 'HW': hw, 'n_obj': 1},
            'peak_source': 'MEASURED_PEAKS.json bf16 burst' if peaks else 'fallback', 'algorithmic_bytes': bytes_,
            'hbm_gbs_if_bytes_bound': round(bytes_ / t / 1e9, 1)}


def oracle_fps(device, n_frames, threads=None, autocast=False, seed=1234):
    """frames/s of the oracle port over the first n_frames of the config-2 clip (bounded sample)."""
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
        run_clip(lambda: O.OracleCore(net, dict(CFG)), frames[:2], {0: masks[0]}, device, False)      # warm-up
        if device != 'cpu':
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_clip(lambda: O.OracleCore(net, dict(CFG)), frames, masks, device, False)
        if device != 'cpu':
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return n_frames / dt


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
              'parallelism': f'{world} independent stream(s), one per GPU, no collective'}

    if args.impl == 'reference':
        if rank != 0:
            return
        nfr = 12                      # bounded sample: 1 preload (frame 0) + 12 frames of the same clip, ~10-20 s per step
        vals = []
        for i in range(args.warmup + args.steps):
            v = oracle_fps('cpu', nfr, threads=cores)
            if i >= args.warmup:
                vals.append(v)
        fps = sum(vals) / len(vals)
        line = {'impl': 'reference', 'metric': 'fps_480p', 'value': round(fps, 4), 'unit': 'frames/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(1000 * nfr / fps, 2), 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': round(fps, 4), 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                 'sample': f'first {nfr} frames of the config-2 clip (frame 0 annotated+preloaded) per step, '
                                           'oracle/xmem_oracle.py (PyTorch fp32 restatement of the reference) on all host threads'},
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
    ms_e2e, _ = timed(True, args.steps, 1)
    total_frames = N_FRAMES * args.steps * world
    value = total_frames / (ms_dev * 1e-3)
    e2e = total_frames / (ms_e2e * 1e-3)
    line = {'metric': 'fps_480p', 'value': round(value, 2), 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(ms_dev / args.steps, 2), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic', 'config': config,
            'e2e': {'value': round(e2e, 2), 'unit': 'frames/s', 'h2d_bytes_per_step': N_FRAMES * 3 * H * W * 4 + len(ANNOTATED) * 2 * H * W * 4
                    + len(ANNOTATED) * 3 * H * W * 4, 'd2h_bytes_per_step': N_FRAMES * H * W},
            'gpu_launches': int(launches), 'clocks': clocks}
    if rank == 0:
        if world == 1:
            try:
                line['roofline'] = k1_roofline(device)
            except Exception as e:                       # never lose the headline number to the side measurement
                line['roofline'] = {'error': str(e)[:200]}
            try:
                nfr = 12
                fps_cpu = oracle_fps('cpu', nfr, threads=cores)
                line['cpu_baseline'] = {'value': round(fps_cpu, 4), 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                        'sample': f'first {nfr} frames of the config-2 clip, oracle/xmem_oracle.py fp32 on all host threads'}
            except Exception as e:
                line['cpu_baseline'] = {'error': str(e)[:200]}
            try:
                fps_ref_gpu = oracle_fps(device, 100, autocast=True)
                line['reference_style_gpu'] = {'value': round(fps_ref_gpu, 2), 'unit': 'frames/s',
                                               'what': 'oracle port on this GPU under torch fp16 autocast (cuDNN/cuBLAS library kernels, '
                                                       'torch.cat memory, materialised affinity) = how the reference runs on a GPU'}
            except Exception as e:
                line['reference_style_gpu'] = {'error': str(e)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
