"""Host-side weight preparation for the tcgen05 implicit-GEMM convolution (csrc/conv_igemm.cu).

BatchNorm (eval mode) is folded into the preceding convolution (reference: nn.BatchNorm2d after every
ResNet conv, model/resnet.py:46-114), weights are laid out [cout_pad][kh*kw][cin_pad] in fp16 so one TMA
box of 64 input channels of one filter tap is a K-step of the GEMM, and biases stay fp32.
"""
from __future__ import annotations

import torch


def _round_up(x, m):
    return (x + m - 1) // m * m


def fold_bn(weight, bias, gamma, beta, mean, var, eps=1e-5):
    scale = gamma / torch.sqrt(var + eps)
    w = weight * scale.view(-1, 1, 1, 1)
    b = beta - mean * scale if bias is None else (bias - mean) * scale + beta
    return w, b


def pack_conv(weight: torch.Tensor, bias, cin_pad=None, device=None):
    """weight [cout, cin, kh, kw] fp32 -> (wp [cout_pad, kh*kw*cin_pad] fp16, bp [cout_pad] fp32, cout)."""
    cout, cin, kh, kw = weight.shape
    cin_pad = cin_pad or _round_up(cin, 64)
    cout_pad = _round_up(cout, 64)
    wp = torch.zeros((cout_pad, kh * kw, cin_pad), dtype=torch.float32)
    wp[:cout, :, :cin] = weight.float().permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    bp = torch.zeros((cout_pad,), dtype=torch.float32)
    if bias is not None:
        bp[:cout] = bias.float()
    wp = wp.reshape(cout_pad, kh * kw * cin_pad).half().contiguous()
    if device is not None:
        wp, bp = wp.to(device), bp.to(device)
    return wp, bp, cout
