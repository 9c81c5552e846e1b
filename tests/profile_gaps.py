"""Where does a frame's wall time go?  Runs the config-2 clip from recorded graphs under torch.profiler (CUPTI activity
records, no ncu serialisation) and prints: GPU busy time vs span, idle gaps, per-kernel warm totals.
    python tests/profile_gaps.py [n_frames]"""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict
from torch.profiler import profile, ProfilerActivity

torch.set_grad_enabled(False)
if os.environ.get('XMEM_NO_PDL'):
    from xmem2_b200 import lib as _lib
    _lib.load().xm_set_pdl(0)          # plain launches: kernel durations are not inflated by waiting in griddepcontrol.wait
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = 'cuda:0'
net = XMem(dict(bench.CFG), None).to(dev).eval(); net.load_weights(synth_state_dict(0))
frames, masks = bench.clip_inputs(1234)
frames = frames.to(dev); masks = {k: v.to(dev) for k, v in masks.items()}
fac = lambda: InferenceCore(net, dict(bench.CFG))
for _ in range(2):
    bench.run_clip(fac, frames[:n], {k: v for k, v in masks.items() if k < n}, dev, False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    bench.run_clip(fac, frames[:n], {k: v for k, v in masks.items() if k < n}, dev, False)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda x: x[0])
span = (ks[-1][1] - ks[0][0])
busy, last_end, gaps = 0.0, ks[0][0], []
tot = collections.defaultdict(lambda: [0.0, 0])
for s, e, name in ks:
    tot[name][0] += e - s; tot[name][1] += 1
    if s > last_end:
        gaps.append(s - last_end)
    busy += max(0.0, e - max(s, last_end))
    last_end = max(last_end, e)
print(f'{n} frames: span {span / 1e3:.1f} ms, GPU busy (union of kernels/copies) {busy / 1e3:.1f} ms = {100 * busy / span:.1f} %, '
      f'{len(ks)} activities, {len(gaps)} gaps totalling {sum(gaps) / 1e3:.1f} ms')
import numpy as np
g = np.array(gaps)
for lo, hi in ((0, 1), (1, 2), (2, 4), (4, 8), (8, 20), (20, 100), (100, 1e9)):
    m = (g >= lo) & (g < hi)
    print(f'  gaps {lo:>4}-{hi:<6} us: n={int(m.sum()):6d}  total {g[m].sum() / 1e3:7.2f} ms')
print('per kernel (warm, inside graphs): total ms | n | avg us')
for name, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f'  {t / 1e3:8.2f} | {c:5d} | {t / c:7.1f} | {name[:90]}')

# one ordinary frame in launch order: the kernels between two consecutive k1_fused launches late in the clip
k1_idx = [i for i, (s_, e_, nm) in enumerate(ks) if 'k1_fused' in nm]
if len(k1_idx) > 60:
    a, b = k1_idx[57], k1_idx[58]
    print(f'--- kernels from one memory read to the next (frame ~58), {b - a} activities, {(ks[b][0] - ks[a][0]):.1f} us:')
    for s_, e_, nm in ks[a:b]:
        short = nm.replace('(anonymous namespace)::', '').replace('void ', '')[:70]
        print(f'   +{s_ - ks[a][0]:8.1f} us  {e_ - s_:7.1f} us  {short}')

# one MEMORY frame: the window between the memory reads around a value-encoder stem launch
v_idx = [i for i, (s_, e_, nm) in enumerate(ks) if 'stem7x7_kernel<5>' in nm]
if len(v_idx) > 8 and len(k1_idx) > 60:
    vi = v_idx[8]
    a = max(i for i in k1_idx if i < vi); later = [i for i in k1_idx if i > vi]
    b = later[0] if later else len(ks) - 1
    print(f'--- a memory frame: kernels from the memory read before the value encoder to the next read, {b - a} activities, {(ks[b][0] - ks[a][0]):.1f} us:')
    for s_, e_, nm in ks[a:b]:
        short = nm.replace('(anonymous namespace)::', '').replace('void ', '')[:70]
        print(f'   +{s_ - ks[a][0]:8.1f} us  {e_ - s_:7.1f} us  {short}')
