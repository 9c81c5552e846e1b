"""Fused resize + argmax (+ label table) kernel against torch on the GPU.  The per-pixel math is
verified on the CPU (tests/test_postproc.py).  Round 1's failure of the 480x864 -> 1080x1920 case came from
--use_fast_math's approximate division in the source-coordinate scale (fixed in csrc/postproc_math.h)."""
import pytest
import torch
import torch.nn.functional as F

from xmem2_b200.inference import postprocess as pp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('in_hw,out_hw', [((480, 864), (480, 864)), ((480, 864), (1080, 1920)), ((96, 128), (57, 75))])
def test_resize_argmax_equals_torch(in_hw, out_hw):
    g = torch.Generator().manual_seed(3)
    prob = torch.softmax(torch.randn(3, *in_hw, generator=g) * 2, dim=0).cuda()
    view = prob[:, : in_hw[0] - 4, 5:]                   # an unpadded view (strides of the padded tensor)
    for p in (prob, view):
        shape = out_hw if p is prob else (out_hw[0] - 3, out_hw[1] - 7)
        got = pp.post_process(p, shape)
        up = F.interpolate(p.unsqueeze(1), shape, mode='bilinear', align_corners=False)[:, 0]
        top2 = up.topk(2, dim=0).values
        clear = (top2[0] - top2[1]) > 1e-5
        assert got.dtype == torch.uint8 and tuple(got.shape) == tuple(shape)
        assert torch.equal(got[clear], up.argmax(0).to(torch.uint8)[clear]) and clear.float().mean() > 0.999
    table = pp.label_table({7: 1, 200: 2})
    mapped = pp.post_process(prob, out_hw, label_table=table)
    assert torch.equal(mapped, table.cuda()[pp.post_process(prob, out_hw).long()])
