"""GPU parity of the tcgen05 implicit-GEMM convolution (xm_conv2d_nhwc) against torch.nn.functional.conv2d
in fp32 on the same fp16-rounded operands (reference call sites: model/resnet.py, model/modules.py)."""
import pytest
import torch
import torch.nn.functional as F

from xmem2_b200 import lib
from xmem2_b200.model.packing import pack_conv

pytestmark = pytest.mark.gpu


def _run(batch, H, W, cins, cout, ksize, stride, relu=False, residual=False, bcast0=False, relu_copy=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    dev = 'cuda'
    srcs, refs = [], []
    for i, c in enumerate(cins):
        nb = 1 if (bcast0 and i == 0) else batch
        x = (torch.randn(nb, H, W, c, generator=g)).half()
        srcs.append((x.to(dev), nb == 1 and batch > 1))
        refs.append(x.float().expand(batch, -1, -1, -1))
    cin = sum(cins)
    w = torch.randn(cout, cin, ksize, ksize, generator=g) * (1.0 / (cin * ksize * ksize) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    wp, bp, _ = pack_conv(w, b, device=dev)
    Ho, Wo = H // stride, W // stride
    res = torch.randn(batch, Ho, Wo, cout, generator=g).half() if residual else None
    out, out_relu = lib.conv2d_nhwc(srcs, wp, bp, cout, ksize=ksize, stride=stride, relu=relu,
                                    residual=res.to(dev) if residual else None, want_relu_copy=relu_copy)
    torch.cuda.synchronize()
    x = torch.cat(refs, 3).permute(0, 3, 1, 2)
    ref = F.conv2d(x, w.half().float(), b, stride=stride, padding=ksize // 2)
    if residual:
        ref = ref + res.float().permute(0, 3, 1, 2)
    raw = ref
    if relu:
        ref = F.relu(ref)
    got = out.float().cpu().permute(0, 3, 1, 2)
    tol = 2e-2 + 4e-3 * ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err < tol, (err, tol)
    if relu_copy:
        err2 = (out_relu.float().cpu().permute(0, 3, 1, 2) - F.relu(raw)).abs().max().item()
        assert err2 < tol, (err2, tol)
    return err


def test_conv1x1_gemm():
    _run(1, 16, 24, [64], 64, 1, 1)


def test_conv1x1_wide_relu():
    _run(1, 30, 54, [256], 1024, 1, 1, relu=True)


def test_conv3x3_padding_edges():
    _run(1, 30, 54, [128], 128, 3, 1, relu=True)


def test_conv3x3_residual_and_relu_copy():
    _run(2, 30, 54, [64], 192, 3, 1, residual=True, relu_copy=True)


def test_conv3x3_concat_sources_with_broadcast():
    # FeatureFusionBlock input: cat([f16 broadcast over objects, readout, hidden]) (modules.py:33-36)
    _run(2, 30, 54, [128, 64, 64], 128, 3, 1, bcast0=True)


def test_conv3x3_stride2():
    _run(1, 60, 108, [128], 128, 3, 2, relu=True)


def test_conv1x1_stride2():
    _run(1, 60, 108, [256], 512, 1, 2)


def test_conv_cout_not_multiple_of_64():
    _run(1, 24, 40, [256], 1, 3, 1)
    _run(1, 24, 40, [64], 129, 3, 1)


def test_relu_copy_through_the_tma_epilogue():
    # decoder.up_8_4 out_conv.conv2 (group_modules.py:47-54: g and relu(g) both live on): BN = 128, 3 stages, residual
    _run(1, 120, 216, [256], 256, 3, 1, residual=True, relu_copy=True, seed=3)
    # BN = 64 variants: short and long K loops, two objects
    _run(2, 30, 54, [128], 64, 3, 1, residual=True, relu_copy=True, seed=4)
    _run(1, 60, 108, [64], 128, 1, 1, relu_copy=True, seed=5)
