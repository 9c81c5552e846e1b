import os, sys, time, torch
sys.path.insert(0,'/root/repo')
import bench
from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict
torch.set_grad_enabled(False)
dev='cuda:0'
net = XMem(dict(bench.CFG), None).to(dev).eval(); net.load_weights(synth_state_dict(0))
frames, masks = bench.clip_inputs(1234); frames=frames.to(dev); masks={k:v.to(dev) for k,v in masks.items()}
fac=lambda: InferenceCore(net, dict(bench.CFG))
for rep in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    bench.run_clip(fac, frames, masks, dev, False)
    torch.cuda.synchronize(); print('clip', time.perf_counter()-t0)
# steady-state per-frame time of graph replay: one long-lived core
core=fac(); core.set_all_labels([1])
for j in masks: core.put_to_permanent_memory(frames[j], masks[j])
core.step(frames[0], masks[0], [1], do_not_add_mask_to_memory=True)
for ti in range(1,6): core.step(frames[ti])
torch.cuda.synchronize(); t0=time.perf_counter()
for ti in range(6,10): core.step(frames[ti])
torch.cuda.synchronize(); print('graph frame ms', (time.perf_counter()-t0)/4*1e3)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): core._graph.replay()
e1.record(); torch.cuda.synchronize(); print('pure replay ms', e0.elapsed_time(e1)/20)
