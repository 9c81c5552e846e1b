"""Phase timeline of the fused read kernel (diagnostics): per-phase durations (max / median over CTAs) at a given shape.
    python tests/profile_k1_timeline.py [hw] [n_work_frames] [n_perm_frames] [n_obj]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import k1_ref
from xmem2_b200 import lib

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 1620
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 9
npm = int(sys.argv[3]) if len(sys.argv) > 3 else 5
n_obj = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = 'cuda'
case = k1_ref.make_case(hw=hw, sizes=(0, nw * hw, npm * hw), n_obj=n_obj, group_begins=[(0, n_obj, [0, 0, 0])], seed=11, device=dev)
L = lib.load()
L.xm_affinity_debug_timeline.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
hw_pad = (hw + 127) // 128 * 128
a = lib.XmAffinityArgs(); keep = []
for bi, b in enumerate(case['banks']):
    if b is None:
        a.banks[bi].size = 0; continue
    rows = torch.zeros(b['cap'], 128, dtype=torch.float16, device=dev)
    lib.key_pack(b['key'].to(dev).contiguous(), rows[:b['n']])
    shr = torch.ones(b['cap'], dtype=torch.float32, device=dev); shr[:b['n']] = b['shr'].to(dev)
    val = b['val'].to(dev).contiguous(); usage = torch.zeros(b['cap'], dtype=torch.float32, device=dev)
    keep += [rows, shr, val, usage]
    bk = a.banks[bi]
    bk.keys, bk.shrinkage, bk.values, bk.usage = rows.data_ptr(), shr.data_ptr(), val.data_ptr(), usage.data_ptr()
    bk.cap, bk.n_obj_cap, bk.size = b['cap'], n_obj, b['n']
a.n_groups = 1; a.groups[0].obj_begin, a.groups[0].n_obj = 0, n_obj
qp, bsq = lib.query_pack(case['qk'].to(dev).contiguous(), case['qe'].to(dev).contiguous(), hw_pad)
ws = lib.affinity_workspace(hw, n_obj, dev)
out = torch.empty(n_obj, hw, 512, dtype=torch.float16, device=dev)
a.qp, a.bsq, a.hw, a.hw_pad, a.top_k, a.n_obj_total = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad, 30, n_obj
a.readout_hwc, a.workspace, a.workspace_bytes = out.data_ptr(), ws.data_ptr(), ws.numel()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
names = ['prologue->start', 'sweep A', 'barrier 1', 'merge A + barrier 2', 'sweep B', 'publish + barrier 3', 'merge B', 'grid barrier',
         'readout', 'grid barrier 2', 'reduce']
for rep in range(3):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.check(L.xm_affinity_readout(C.byref(a), lib.stream_ptr()), 'xm_affinity_readout')
    e1.record(); torch.cuda.synchronize()
    buf = np.zeros((160, 16), dtype=np.uint64)
    n = L.xm_affinity_debug_timeline(ws.data_ptr(), hw, n_obj, buf.ctypes.data, 160)
    t = buf[:n, :11].astype(np.int64)
    tb = buf[:n, 11:16].astype(np.int64)
    t0 = t[:, 0].min()
    if rep < 2 or True and rep < 2:
        continue
    print(f'--- rep {rep}: event time {e0.elapsed_time(e1) * 1e3:.1f} us, N = {(nw + npm) * hw}, hw = {hw}, n_obj = {n_obj}, CTAs = {n}')
    print(f'  start skew across CTAs: {(t[:, 0].max() - t0) / 1e3:.1f} us')
    for i in range(1, 11):
        d = (t[:, i] - t[:, i - 1]) / 1e3
        ok = t[:, i] > 0
        print(f'  {names[i]:24s} median {np.median(d[ok]):7.1f}  max {d[ok].max():7.1f}  min {d[ok].min():7.1f} us   (phase ends at {(t[ok, i].max() - t0) / 1e3:7.1f})')
cnts = np.zeros(hw, dtype=np.int32)
L.xm_affinity_debug_counts.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
L.xm_affinity_debug_counts(ws.data_ptr(), hw, n_obj, cnts.ctypes.data)
print(f'candidates per query after sweep B: mean {cnts.mean():.1f}, median {np.median(cnts):.0f}, max {cnts.max()}, min {cnts.min()}')
t0 = t[:, 0].min()
rel = lambda x: (x - t0) / 1e3
sw_ctas = t[:, 5] > 0
print(f'merge B (worker stamp 11) ends: median {np.median(rel(tb[sw_ctas, 0])):.1f} max {rel(tb[sw_ctas, 0]).max():.1f} | barrier 3 released (stamp 5) median {np.median(rel(t[sw_ctas, 5])):.1f}')
print('readout detail per CTA (us): tables built - barrier released | MMAs done - tables built | drained - MMAs done | stamp 8 - drained')
d1, d2, d3, d4 = (tb[:, 1] - t[:, 7]) / 1e3, (tb[:, 2] - tb[:, 1]) / 1e3, (tb[:, 3] - tb[:, 2]) / 1e3, (t[:, 8] - tb[:, 3]) / 1e3
d0 = (tb[:, 4] - t[:, 7]) / 1e3
for nm, d in (('counted', d0), ('tables', d1), ('mma', d2), ('drain', d3), ('tail', d4)):
    print(f'  {nm:8s} median {np.median(d):6.1f}  min {d.min():6.1f}  max {d.max():6.1f}')
order = np.argsort(d2)
print('  slowest MMA phases (cta, us):', [(int(c), round(float(d2[c]), 1)) for c in order[-8:]], ' fastest:', [(int(c), round(float(d2[c]), 1)) for c in order[:6]])
print('  MMA phase by CTA index (every 4th):', [round(float(x), 1) for x in d2[::4]])
if os.environ.get('K1_TRACE'):
    # cycle accounting of CTA 0 (library built with -DK1_TRACE)
    L.xm_affinity_debug_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    tb = np.zeros(4096, dtype=np.uint64)
    L.xm_affinity_debug_trace(ws.data_ptr(), hw, n_obj, tb.ctypes.data)
    tbm = tb[18 * 8:34 * 8].reshape(16, 2, 4).astype(np.int64)
    tb = tb[:18 * 8].reshape(18, 2, 4).astype(np.int64)
    print('merge B per warp (gather, select, rest clks):', [tuple(int(x) for x in tbm[w_, 0, :3]) for w_ in range(0, 16, 3)])
    for sw_ in range(2):
        print(f'sweep {"AB"[sw_]}: producer wait kempty {tb[0, sw_, 0]} clks over {tb[0, sw_, 1]} tiles | mma: wait kfull {tb[1, sw_, 0]}, wait sempty {tb[1, sw_, 1]}, '
              f'issue {tb[1, sw_, 2]}')
        for w_ in range(16):
            print(f'   scan warp {w_:2d} (h{w_ >> 3},b{(w_ >> 2) & 1},quad{(w_ + 2) & 3}): wait {tb[2 + w_, sw_, 0]:7d}  scan {tb[2 + w_, sw_, 1]:7d} clks over {tb[2 + w_, sw_, 2]} tiles')
