"""Post-processing kernel math on the CPU (reference inference/run_on_video.py:165-173 `_post_process`):
the per-pixel function the CUDA kernel executes (xmem2_b200/csrc/postproc_math.h) is compiled with gcc into a host harness and
compared with F.interpolate(mode='bilinear', align_corners=False) + argmax for up-scaling, down-scaling, identity, a strided
(unpadded) view and a label look-up table."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp('harness') / 'postproc_host.so')
    subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-o', so, os.path.join(ROOT, 'tests', 'host_harness', 'postproc_host.c')])
    lib = C.CDLL(so)
    lib.resize_argmax_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p]
    return lib


def _run(host, prob, out_hw, lut=None):
    c, h, w = prob.shape
    out = np.zeros(out_hw, dtype=np.uint8)
    assert prob.stride(2) == 1
    host.resize_argmax_host(prob.data_ptr(), c, h, w, prob.stride(0), prob.stride(1), out_hw[0], out_hw[1],
                            lut.ctypes.data if lut is not None else None, out.ctypes.data)
    return torch.from_numpy(out)


def _reference(prob, out_hw):
    up = F.interpolate(prob.unsqueeze(1), out_hw, mode='bilinear', align_corners=False)[:, 0]
    top2 = up.topk(2, dim=0).values
    return up.argmax(0).to(torch.uint8), top2[0] - top2[1]


@pytest.mark.parametrize('in_hw,out_hw', [((30, 54), (480, 854)), ((480, 864), (480, 864)), ((96, 128), (57, 75)),
                                          ((37, 53), (111, 160)), ((64, 64), (1, 1)), ((1, 7), (5, 9))])
def test_pixel_math_equals_torch_bilinear_argmax(host, in_hw, out_hw):
    g = torch.Generator().manual_seed(in_hw[0] * 1000 + out_hw[1])
    for n_obj in (1, 3):
        logits = torch.randn(n_obj + 1, *in_hw, generator=g) * 2
        prob = torch.softmax(logits, dim=0).contiguous()
        got = _run(host, prob, out_hw)
        want, gap = _reference(prob, out_hw)
        clear = gap > 1e-6
        assert torch.equal(got[clear], want[clear])
        assert clear.float().mean() > 0.999


def test_strided_view_and_label_table(host):
    g = torch.Generator().manual_seed(5)
    padded = torch.softmax(torch.randn(3, 48, 64, generator=g), dim=0)
    view = padded[:, 3:45, 5:60]                      # what unpad() hands out: same strides, smaller extent
    got = _run(host, view, (84, 110))
    want, gap = _reference(view.contiguous(), (84, 110))
    assert torch.equal(got[gap > 1e-6], want[gap > 1e-6])
    lut = np.arange(256, dtype=np.uint8)
    lut[1], lut[2] = 7, 200                            # MaskMapper.remap_index_mask: internal index -> original label
    mapped = _run(host, view, (84, 110), lut)
    assert torch.equal(mapped, torch.from_numpy(lut)[got.long()])


def test_python_wrapper_has_no_cpu_path_and_builds_the_label_table():
    from xmem2_b200.inference import postprocess as pp
    t = pp.label_table({5: 1, 9: 2})                   # MaskMapper.remappings: original label -> internal index
    assert t[0] == 0 and t[1] == 5 and t[2] == 9 and t[3] == 3 and t.dtype == torch.uint8 and t.numel() == 256
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pp.post_process(torch.rand(2, 8, 8))
    from xmem2_b200 import lib
    L = lib.load()
    assert L.xm_resize_argmax(None, 2, 8, 8, 64, 8, 8, 8, None, None, None) != 0
    assert b'null pointer' in L.xm_last_error()
