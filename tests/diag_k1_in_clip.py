"""K1 inside the real pipeline: phase timeline and candidate counts of the LAST memory read of a config-2 clip prefix
(real keys / values from the network, not the synthetic bank of the roofline measurement).
    python tests/diag_k1_in_clip.py [n_frames]"""
import ctypes as C
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xmem2_b200 import lib
from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict

torch.set_grad_enabled(False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 98
dev = 'cuda:0'
net = XMem(dict(bench.CFG), None).to(dev).eval(); net.load_weights(synth_state_dict(0))
frames, masks = bench.clip_inputs(1234)
frames = frames.to(dev); masks = {k: v.to(dev) for k, v in masks.items()}
cfg = dict(bench.CFG); cfg['use_cuda_graph'] = False
core = InferenceCore(net, cfg)
core.set_all_labels([1])
for j in masks:
    core.put_to_permanent_memory(frames[j], masks[j])
L = lib.load()
L.xm_affinity_debug_timeline.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
L.xm_affinity_debug_counts.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
names = ['', 'sweep A', 'barrier 1', 'merge A + barrier 2', 'sweep B', 'barrier 3', '(stamp 6)', 'merge B + grid barrier', 'readout', 'row barrier', 'reduce']
for ti in range(n):
    msk = masks.get(ti)
    core.step(frames[ti], msk, [1] if msk is not None else None, end=False, do_not_add_mask_to_memory=msk is not None)
    if ti in (12, 45, n - 1) and msk is None:
        torch.cuda.synchronize()
        mm = core.memory
        ws = mm._ws
        hw = 30 * 54
        buf = np.zeros((160, 16), dtype=np.uint64)
        g = L.xm_affinity_debug_timeline(ws.data_ptr(), hw, 1, buf.ctypes.data, 160)
        t = buf[:g, :11].astype(np.int64); t0 = t[:, 0].min()
        cnts = np.zeros(hw, dtype=np.int32)
        L.xm_affinity_debug_counts(ws.data_ptr(), hw, 1, cnts.ctypes.data)
        sizes = (mm.long_mem.size if getattr(mm, 'long_mem', None) is not None else 0, mm.temporary_work_mem.size, mm.permanent_work_mem.size)
        print(f'frame {ti}: banks (long, work, perm) = {sizes}; kernel end at {(t[:, 10].max() - t0) / 1e3:.1f} us')
        print('   phase ends:', ', '.join(f'{names[i]} {(t[t[:, i] > 0, i].max() - t0) / 1e3:.1f}' for i in range(1, 11)))
        print(f'   candidates per query after sweep B: mean {cnts.mean():.1f} median {np.median(cnts):.0f} max {cnts.max()} | >128: {(cnts > 128).sum()} queries, >512: {(cnts > 512).sum()}')
