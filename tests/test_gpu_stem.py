"""GPU parity of the fused 7x7/stride-2 stem convolution (xm_stem7x7, csrc/stem7x7.cu) against torch.nn.functional.conv2d in
fp32 on the same fp16-rounded operands.  Reference call sites: KeyEncoder conv1+bn1+relu (model/modules.py:165-168) and
ValueEncoder `torch.cat([image, mask, others], 1)` -> conv1 -> bn1 (model/modules.py:124-137)."""
import pytest
import torch
import torch.nn.functional as F

from xmem2_b200 import lib
from xmem2_b200.model.packing import pack_conv

pytestmark = pytest.mark.gpu


def _pack_stem(w, b, dev):
    """the K layouts of XMem._ensure_packed (model/network.py put_stem)"""
    if w.shape[1] == 3:
        rows = torch.zeros((w.shape[0], 8, 24))
        rows[:, :7, :21] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 7, 21)
        flat, kpad = rows.reshape(w.shape[0], 192), 192
    else:
        flat, kpad = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), 256
    wp, bp, _ = pack_conv(flat[:, :, None, None], b, cin_pad=kpad, device=dev)
    return wp, bp, kpad


def _run(H, W, n_obj, with_masks, relu, seed=0):
    g = torch.Generator().manual_seed(seed)
    dev = 'cuda'
    C = 5 if with_masks else 3
    img = torch.randn(3, H, W, generator=g)
    masks = torch.rand(n_obj, H, W, generator=g) if with_masks else None
    w = torch.randn(64, C, 7, 7, generator=g) * (1.0 / (C * 49) ** 0.5)
    b = torch.randn(64, generator=g) * 0.1
    wp, bp, kpad = _pack_stem(w, b, dev)
    out = torch.empty(n_obj, H // 2, W // 2, 64, dtype=torch.float16, device=dev)
    img_d = img.to(dev).contiguous()
    mk_d = masks.to(dev).contiguous() if with_masks else None
    lib.check(lib.load().xm_stem7x7(img_d.data_ptr(), mk_d.data_ptr() if with_masks else None, n_obj, H, W, wp.data_ptr(), bp.data_ptr(),
                                    kpad, 1 if relu else 0, out.data_ptr(), lib.stream_ptr()), 'xm_stem7x7')
    torch.cuda.synchronize()
    # fp32 reference on the operands as the kernel sees them (fp16-rounded inputs and weights)
    xs = []
    for o in range(n_obj):
        chans = [img]
        if with_masks:
            others = torch.zeros(H, W)          # the kernel sums the other masks in fp32 in index order
            for j in range(n_obj):
                if j != o:
                    others = others + masks[j]
            chans += [masks[o:o + 1], others[None]]
        xs.append(torch.cat(chans, 0))
    x = torch.stack(xs).half().float()
    ref = F.conv2d(x, w.half().float(), b, stride=2, padding=3)
    if relu:
        ref = F.relu(ref)
    got = out.float().cpu().permute(0, 3, 1, 2)
    tol = 2e-2 + 4e-3 * ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err < tol, (err, tol)
    # and against the round-1 path (materialised im2col + 1x1 GEMM) on the same packed weights
    col = torch.empty(n_obj, H // 2, W // 2, kpad, dtype=torch.float16, device=dev)
    lib.check(lib.load().xm_im2col_stem(img_d.data_ptr(), mk_d.data_ptr() if with_masks else None, n_obj, H, W, kpad, col.data_ptr(),
                                        lib.stream_ptr()), 'xm_im2col_stem')
    old, _ = lib.conv2d_nhwc([(col, False)], wp, bp, 64, ksize=1, stride=1, relu=relu)
    torch.cuda.synchronize()
    assert (old.float() - out.float()).abs().max().item() < 1e-2
    return err


def test_key_stem_480p():
    _run(480, 864, 1, False, True)


def test_key_stem_ragged_edges():
    # output 24 x 40: the last tile column is half empty, rows are a multiple of 8
    _run(48, 80, 1, False, True, seed=1)
    # output 20 x 24: partial tiles in both directions
    _run(40, 48, 1, False, True, seed=2)


def test_value_stem_one_object():
    _run(96, 160, 1, True, False, seed=3)


def test_value_stem_three_objects():
    _run(64, 96, 3, True, False, seed=4)


def test_value_stem_480p_two_objects():
    _run(480, 864, 2, True, False, seed=5)
