"""Profiling helper (run under ncu with --profile-from-start off): one warm clip, then a profiled slice of the
config-2 clip (frames 0..N-1) so the launch list covers ordinary frames, memory frames and annotated frames."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict

torch.set_grad_enabled(False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = 'cuda:0'
net = XMem(dict(bench.CFG), None).to(dev).eval(); net.load_weights(synth_state_dict(0))
frames, masks = bench.clip_inputs(1234)
frames = frames.to(dev); masks = {k: v.to(dev) for k, v in masks.items()}
cfg = dict(bench.CFG); cfg['use_cuda_graph'] = False      # eager launches so that ncu sees every kernel individually
fac = lambda: InferenceCore(net, dict(cfg))
bench.run_clip(fac, frames[:12], {0: masks[0]}, dev, False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
bench.run_clip(fac, frames[:n], {k: v for k, v in masks.items() if k < n}, dev, False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
