"""Synthetic parameters and clips for the XMem++ hot path.

The reference ships no weights (`scripts/download_models.sh:1` needs network) and no
tests, so every parity run in this repo uses *named, hash-seeded* synthetic
parameters: each tensor is drawn from its own `torch.Generator` seeded by the
CRC32 of its state-dict key, so the reference network (golden generation), the
CPU oracle and the CUDA implementation all see bit-identical parameters without
shipping a 250 MB checkpoint.

`xmem_param_spec` enumerates the upstream `XMem.pth` state-dict layout
(reference `model/network.py:18-38`, `model/modules.py`, `model/resnet.py:117-164`,
`model/cbam.py`), which is what `XMem.load_weights` must accept.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bn(spec, prefix, c):
    spec[prefix + '.weight'] = ((c,), 'bn_gamma')
    spec[prefix + '.bias'] = ((c,), 'bn_beta')
    spec[prefix + '.running_mean'] = ((c,), 'bn_mean')
    spec[prefix + '.running_var'] = ((c,), 'bn_var')
    spec[prefix + '.num_batches_tracked'] = ((), 'bn_count')


def _conv(spec, prefix, cout, cin, k, bias=True):
    spec[prefix + '.weight'] = ((cout, cin, k, k), 'conv_w')
    if bias:
        spec[prefix + '.bias'] = ((cout,), 'conv_b')


def _bottleneck_layer(spec, prefix, inplanes, planes, blocks, stride):
    for b in range(blocks):
        p = f'{prefix}.{b}'
        cin = inplanes if b == 0 else planes * 4
        _conv(spec, p + '.conv1', planes, cin, 1, bias=False); _bn(spec, p + '.bn1', planes)
        _conv(spec, p + '.conv2', planes, planes, 3, bias=False); _bn(spec, p + '.bn2', planes)
        _conv(spec, p + '.conv3', planes * 4, planes, 1, bias=False); _bn(spec, p + '.bn3', planes * 4)
        if b == 0 and (stride != 1 or inplanes != planes * 4):
            _conv(spec, p + '.downsample.0', planes * 4, inplanes, 1, bias=False)
            _bn(spec, p + '.downsample.1', planes * 4)
    return planes * 4


def _basic_layer(spec, prefix, inplanes, planes, blocks, stride):
    for b in range(blocks):
        p = f'{prefix}.{b}'
        cin = inplanes if b == 0 else planes
        _conv(spec, p + '.conv1', planes, cin, 3, bias=False); _bn(spec, p + '.bn1', planes)
        _conv(spec, p + '.conv2', planes, planes, 3, bias=False); _bn(spec, p + '.bn2', planes)
        if b == 0 and (stride != 1 or inplanes != planes):
            _conv(spec, p + '.downsample.0', planes, inplanes, 1, bias=False)
            _bn(spec, p + '.downsample.1', planes)
    return planes


def _group_res_block(spec, prefix, cin, cout):
    if cin != cout:
        _conv(spec, prefix + '.downsample', cout, cin, 3)
    _conv(spec, prefix + '.conv1', cout, cin, 3)
    _conv(spec, prefix + '.conv2', cout, cout, 3)


def _fusion_block(spec, prefix, x_in, g_in, g_mid, g_out):
    _group_res_block(spec, prefix + '.block1', x_in + g_in, g_mid)
    spec[prefix + '.attention.ChannelGate.mlp.1.weight'] = ((g_mid // 16, g_mid), 'linear_w')
    spec[prefix + '.attention.ChannelGate.mlp.1.bias'] = ((g_mid // 16,), 'conv_b')
    spec[prefix + '.attention.ChannelGate.mlp.3.weight'] = ((g_mid, g_mid // 16), 'linear_w')
    spec[prefix + '.attention.ChannelGate.mlp.3.bias'] = ((g_mid,), 'conv_b')
    _conv(spec, prefix + '.attention.SpatialGate.spatial.conv', 1, 2, 7)
    _group_res_block(spec, prefix + '.block2', g_mid, g_out)


def xmem_param_spec(key_dim=64, value_dim=512, hidden_dim=64, single_object=False):
    """Ordered {state_dict key: (shape, kind)} of the upstream XMem checkpoint."""
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    # key encoder: ResNet-50 up to layer3 (modules.py:153-175; layer1 is renamed res2)
    _conv(spec, 'key_encoder.conv1', 64, 3, 7, bias=False); _bn(spec, 'key_encoder.bn1', 64)
    c = _bottleneck_layer(spec, 'key_encoder.res2', 64, 64, 3, 1)
    c = _bottleneck_layer(spec, 'key_encoder.layer2', c, 128, 4, 2)
    c = _bottleneck_layer(spec, 'key_encoder.layer3', c, 256, 6, 2)
    # value encoder: ResNet-18 up to layer3 with a (3 + 1|2)-channel stem (modules.py:102-122)
    extra = 1 if single_object else 2
    _conv(spec, 'value_encoder.conv1', 64, 3 + extra, 7, bias=False); _bn(spec, 'value_encoder.bn1', 64)
    c = _basic_layer(spec, 'value_encoder.layer1', 64, 64, 2, 1)
    c = _basic_layer(spec, 'value_encoder.layer2', c, 128, 2, 2)
    c = _basic_layer(spec, 'value_encoder.layer3', c, 256, 2, 2)
    _fusion_block(spec, 'value_encoder.fuser', 1024, 256, value_dim, value_dim)
    if hidden_dim > 0:
        _conv(spec, 'value_encoder.hidden_reinforce.transform', hidden_dim * 3, value_dim + hidden_dim, 3)
    # key projection (modules.py:194-211)
    _conv(spec, 'key_proj.key_proj', key_dim, 1024, 3)
    _conv(spec, 'key_proj.d_proj', 1, 1024, 3)
    _conv(spec, 'key_proj.e_proj', key_dim, 1024, 3)
    # decoder (modules.py:214-250)
    _fusion_block(spec, 'decoder.fuser', 1024, value_dim + hidden_dim, 512, 512)
    if hidden_dim > 0:
        _conv(spec, 'decoder.hidden_update.g16_conv', 256, 512, 1)
        _conv(spec, 'decoder.hidden_update.g8_conv', 256, 256, 1)
        _conv(spec, 'decoder.hidden_update.g4_conv', 256, 257, 1)
        _conv(spec, 'decoder.hidden_update.transform', hidden_dim * 3, 256 + hidden_dim, 3)
    _conv(spec, 'decoder.up_16_8.skip_conv', 512, 512, 3)
    _group_res_block(spec, 'decoder.up_16_8.out_conv', 512, 256)
    _conv(spec, 'decoder.up_8_4.skip_conv', 256, 256, 3)
    _group_res_block(spec, 'decoder.up_8_4.out_conv', 256, 256)
    _conv(spec, 'decoder.pred', 1, 256, 3)
    return spec


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_state_dict(seed: int = 0, **dims) -> "OrderedDict[str, torch.Tensor]":
    """Well-conditioned synthetic parameters keyed like the upstream checkpoint.

    Scales are chosen so activations stay O(1..10) through the 60-odd layers in
    fp16 and the decoder emits a non-degenerate mask (not 99 % background as a
    default-initialised network does, SURVEY.md section 7 'hard parts').
    """
    out = OrderedDict()
    for name, (shape, kind) in xmem_param_spec(**dims).items():
        g = _gen(name, seed)
        if kind == 'conv_w':
            fan_in = shape[1] * shape[2] * shape[3]
            gain = math.sqrt(2.0 / fan_in)
            if name.endswith(('block1.conv2.weight', 'block2.conv2.weight', 'out_conv.conv2.weight')):
                gain *= 0.5      # residual branch of GroupResBlock (no norm layer there)
            if name.startswith('key_proj.key_proj'):
                gain = 0.18 * math.sqrt(1.0 / fan_in)   # |key| ~ 0.7: similarities stay in exp() range
            if name.startswith('key_proj.d_proj'):
                gain = 0.15 * math.sqrt(1.0 / fan_in)   # shrinkage = d^2+1 in ~[1,3]
            if name.startswith('key_proj.e_proj'):
                gain = 0.25 * math.sqrt(1.0 / fan_in)   # selection = sigmoid(.) not saturated
            if name == 'decoder.pred.weight':
                gain = 2.0 * math.sqrt(1.0 / fan_in)
            t = torch.randn(shape, generator=g) * gain
        elif kind == 'linear_w':
            t = torch.randn(shape, generator=g) * math.sqrt(1.0 / shape[1])
        elif kind == 'conv_b':
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == 'bn_gamma':
            t = 0.8 + 0.4 * torch.rand(shape, generator=g)
            if name.endswith(('bn3.weight',)) or ('value_encoder.layer' in name and name.endswith('bn2.weight')):
                t = t * 0.3      # last norm of each residual block: keep the trunk variance tame
        elif kind == 'bn_beta':
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == 'bn_mean':
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == 'bn_var':
            t = 0.8 + 0.4 * torch.rand(shape, generator=g)
        elif kind == 'bn_count':
            t = torch.zeros(shape, dtype=torch.long)
        else:
            raise KeyError(kind)
        out[name] = t
    return out


def synth_frame(ti: int, height: int, width: int, seed: int = 1234, structured: bool = False) -> torch.Tensor:
    """3xHxW ImageNet-normalised frame. `structured=False` is BASELINE.json config 2
    (U[0,1) noise per pixel from Generator(seed+ti)); `structured=True`
    adds slowly moving low-frequency content so features vary across the image."""
    g = torch.Generator(device='cpu'); g.manual_seed(seed + ti)
    img = torch.rand((3, height, width), generator=g)
    if structured:
        yy = torch.linspace(0, 1, height).view(1, height, 1)
        xx = torch.linspace(0, 1, width).view(1, 1, width)
        ph = torch.tensor([0.0, 2.1, 4.2]).view(3, 1, 1) + 0.07 * ti
        low = 0.5 + 0.25 * torch.sin(6.28 * (1.5 * xx + 0.5 * yy) + ph) + 0.25 * torch.cos(6.28 * (2.5 * yy - xx) - ph)
        img = 0.7 * low + 0.3 * img
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    return (img - mean) / std


def synth_mask(ti: int, height: int, width: int, num_objects: int = 1, first_frame_of=None) -> torch.Tensor:
    """[num_objects,H,W] one-hot float masks: one moving ellipse per object.
    `first_frame_of[o]` = first ti at which object o exists (all-zero before)."""
    yy = torch.arange(height, dtype=torch.float32).view(height, 1)
    xx = torch.arange(width, dtype=torch.float32).view(1, width)
    out = torch.zeros((num_objects, height, width))
    taken = torch.zeros((height, width), dtype=torch.bool)
    for o in range(num_objects):
        if first_frame_of is not None and ti < first_frame_of[o]:
            continue
        cy = height * (0.35 + 0.3 * o + 0.1 * math.sin(0.05 * ti + o))
        cx = width * (0.3 + 0.35 * o + 0.15 * math.cos(0.04 * ti + 2 * o))
        ry, rx = height * 0.18, width * 0.12
        m = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1.0
        m = m & ~taken
        taken |= m
        out[o] = m.float()
    return out
