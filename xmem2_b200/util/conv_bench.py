"""Micro-benchmark of the convolution family: every distinct convolution shape of the 480p network (reference model/resnet.py,
model/modules.py), each recorded 20x in a CUDA graph and CUDA-event timed over 10 replays.  `conv_table()` returns per-layer
microseconds / TFLOP/s and the frame-level sum weighted by how often a shape occurs in one ordinary frame; `python -m
xmem2_b200.util.conv_bench [substring]` prints it (bench.py reports the sum as `conv_roofline`)."""
import sys

import torch

from .. import lib
from ..model.packing import pack_conv

# (name, batch, H, W, [cin...], cout, k, stride, count_per_frame)
h, w = 30, 54
SHAPES = [
    ('stem 7x7 s2 3->64 (fused)', 1, 480, 864, [3], 64, 7, 2, 1),
    ('res2 1x1 64->64',     1, 120, 216, [64], 64, 1, 1, 1),
    ('res2 3x3 64',         1, 120, 216, [64], 64, 3, 1, 3),
    ('res2 1x1 64->256',    1, 120, 216, [64], 256, 1, 1, 4),
    ('res2 1x1 256->64',    1, 120, 216, [256], 64, 1, 1, 2),
    ('l2 1x1 256->128',     1, 120, 216, [256], 128, 1, 1, 1),
    ('l2 3x3 128 s2',       1, 120, 216, [128], 128, 3, 2, 1),
    ('l2 ds 256->512 s2',   1, 120, 216, [256], 512, 1, 2, 1),
    ('l2 1x1 128->512',     1, 60, 108, [128], 512, 1, 1, 4),
    ('l2 1x1 512->128',     1, 60, 108, [512], 128, 1, 1, 3),
    ('l2 3x3 128',          1, 60, 108, [128], 128, 3, 1, 3),
    ('l3 1x1 512->256',     1, 60, 108, [512], 256, 1, 1, 1),
    ('l3 3x3 256 s2',       1, 60, 108, [256], 256, 3, 2, 1),
    ('l3 ds 512->1024 s2',  1, 60, 108, [512], 1024, 1, 2, 1),
    ('l3 1x1 256->1024',    1, h, w, [256], 1024, 1, 1, 6),
    ('l3 1x1 1024->256',    1, h, w, [1024], 256, 1, 1, 5),
    ('l3 3x3 256',          1, h, w, [256], 256, 3, 1, 5),
    ('keyproj 3x3 1024->129', 1, h, w, [1024], 129, 3, 1, 1),
    ('fuser 3x3 1600->512', 1, h, w, [1024, 512, 64], 512, 3, 1, 2),
    ('fuser 3x3 512->512',  1, h, w, [512], 512, 3, 1, 3),
    ('up16 skip 3x3 512',   1, 60, 108, [512], 512, 3, 1, 1),
    ('up16 3x3 512->256',   1, 60, 108, [512], 256, 3, 1, 2),
    ('up16 3x3 256->256',   1, 60, 108, [256], 256, 3, 1, 1),
    ('up8 3x3 256 @1/4',    1, 120, 216, [256], 256, 3, 1, 3),
    ('pred 3x3 256->1',     1, 120, 216, [256], 1, 3, 1, 1),
    ('hu 1x1 (512|256|320)->256', 1, h, w, [512, 256, 320], 256, 1, 1, 1),
    ('hu 3x3 320->192',     1, h, w, [256, 64], 192, 3, 1, 1),
]


def _time(launch, reps, n):
    """microseconds per launch: n back-to-back launches recorded in a CUDA graph (so the host launch path is not what gets
    timed), CUDA-event timed over `reps` replays"""
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(n):
            launch()
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)


def conv_table(dev='cuda', only=None, reps=10, n=20):
    g = torch.Generator().manual_seed(0)
    rows, tot_us, tot_fl = [], 0.0, 0.0
    for name, B, H, W, cins, cout, k, s, cnt in SHAPES:
        if only and only not in name:
            continue
        if k == 7:                                   # the fused stem kernel (csrc/stem7x7.cu), not an implicit GEMM launch
            img = torch.randn(3, H, W, generator=g).to(dev)
            wgt = (torch.randn(64, 192, generator=g) * 0.08).half().to(dev); bias = torch.zeros(64, device=dev)
            out = torch.empty(1, H // 2, W // 2, 64, dtype=torch.float16, device=dev)
            L = lib.load()
            launch = lambda: lib.check(L.xm_stem7x7(img.data_ptr(), None, 1, H, W, wgt.data_ptr(), bias.data_ptr(), 192, 1, out.data_ptr(),
                                                     lib.stream_ptr()), 'xm_stem7x7')
            us = _time(launch, reps, n)
            fl = 2.0 * (H // 2) * (W // 2) * 64 * 147
            tot_us += us * cnt; tot_fl += fl * cnt
            rows.append(dict(layer=name, us=round(us, 2), tflops=round(fl / us / 1e6, 1), per_frame=cnt))
            continue
        srcs = [((torch.randn(B, H, W, c, generator=g)).half().to(dev), False) for c in cins]
        cin = sum(cins)
        wgt = torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k) ** 0.5)
        wp, bp, _ = pack_conv(wgt, torch.zeros(cout), device=dev)
        out = torch.empty(B, H // s, W // s, cout, dtype=torch.float16, device=dev)
        us = _time(lambda: lib.conv2d_nhwc(srcs, wp, bp, cout, ksize=k, stride=s, relu=True, out=out), reps, n)
        fl = 2.0 * B * (H // s) * (W // s) * cout * cin * k * k
        tot_us += us * cnt; tot_fl += fl * cnt
        rows.append(dict(layer=name, us=round(us, 2), tflops=round(fl / us / 1e6, 1), per_frame=cnt))
    return rows, tot_us, tot_fl


if __name__ == '__main__':
    rows, tot_us, tot_fl = conv_table(only=sys.argv[1] if len(sys.argv) > 1 else None)
    for r in rows:
        print(f"{r['layer']:26s} {r['us']:8.1f} us  {r['tflops']:8.1f} TFLOP/s  x{r['per_frame']}")
    print(f'frame conv sum: {tot_us:.0f} us, {tot_fl / 1e9:.1f} GFLOP, {tot_fl / tot_us / 1e6:.1f} TFLOP/s average')
