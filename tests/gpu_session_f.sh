#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stem.py tests/test_gpu_network.py tests/test_gpu_clip.py -x -q > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.txt
timeout 100 python - <<'PY' 2>&1 | tail -5
import torch, time
from xmem2_b200 import lib
torch.set_grad_enabled(False)
dev='cuda'
H,W=480,864
img=torch.randn(3,H,W,device=dev); w=torch.randn(64,192,device=dev).half(); b=torch.zeros(64,device=dev)
out=torch.empty(1,H//2,W//2,64,dtype=torch.float16,device=dev)
L=lib.load()
def run():
    lib.check(L.xm_stem7x7(img.data_ptr(), None, 1, H, W, w.data_ptr(), b.data_ptr(), 192, 1, out.data_ptr(), lib.stream_ptr()), 'stem')
for _ in range(5): run()
torch.cuda.synchronize()
g=torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20): run()
g.replay(); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): g.replay()
e1.record(); torch.cuda.synchronize()
print('fused key stem: %.2f us per launch' % (e0.elapsed_time(e1)*1e3/200))
mk=torch.rand(2,H,W,device=dev); w5=torch.randn(64,256,device=dev).half(); out5=torch.empty(2,H//2,W//2,64,dtype=torch.float16,device=dev)
def run5():
    lib.check(L.xm_stem7x7(img.data_ptr(), mk.data_ptr(), 2, H, W, w5.data_ptr(), b.data_ptr(), 256, 0, out5.data_ptr(), lib.stream_ptr()), 'stem')
for _ in range(5): run5()
torch.cuda.synchronize()
e0.record()
for _ in range(50): run5()
e1.record(); torch.cuda.synchronize()
print('fused value stem (2 objects): %.2f us per launch (eager)' % (e0.elapsed_time(e1)*1e3/50))
PY
echo done
