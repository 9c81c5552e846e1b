"""XMem network on the B200-native kernels — drop-in for the reference `model/network.py` (XMem, :17-198).

Same constructor, same state-dict keys (so upstream `XMem.pth` loads unchanged), same
`encode_key / encode_value / segment / load_weights` signatures and tensor SHAPES.  Internally every
activation is an NHWC fp16 device buffer and every layer is a call into libxmem2_b200.so (tcgen05
implicit-GEMM convolutions + small fused element-wise kernels); the tensors handed back are NCHW-shaped
*views* of those buffers, so the reference drivers (`inference/run_on_video.py`) can index them as before.

There is no CPU path: calling any forward method with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from .. import lib
from ..util.synth import xmem_param_spec
from .packing import fold_bn, pack_conv

Tensor = torch.Tensor


def _as_nhwc(t: Tensor, dtype=torch.float16) -> Tensor:
    """[B,C,H,W]-shaped tensor -> contiguous [B,H,W,C] (zero-copy when `t` is already a view of NHWC memory)."""
    p = t.permute(0, 2, 3, 1)
    if p.dtype != dtype:
        p = p.to(dtype)
    return p if p.is_contiguous() else p.contiguous()


def _nchw_view(t: Tensor) -> Tensor:
    return t.permute(0, 3, 1, 2)


class XMem(nn.Module):
    def __init__(self, config, model_path=None, map_location=None, pretrained_key_encoder=True, pretrained_value_encoder=True):
        super().__init__()
        # pretrained_*_encoder are accepted for signature compatibility only: the reference downloads ImageNet ResNet
        # weights for them (model/resnet.py:154-164); here weights always come from `model_path` / `load_weights`.
        del pretrained_key_encoder, pretrained_value_encoder
        model_weights = self.init_hyperparameters(config, model_path, map_location)
        self.single_object = config.get('single_object', False)
        if self.single_object:
            raise NotImplementedError('single_object checkpoints are not supported on this path')
        self._spec = xmem_param_spec(self.key_dim, self.value_dim, self.hidden_dim, self.single_object)
        if self.key_dim != 64 or self.value_dim != 512 or self.hidden_dim != 64:
            raise NotImplementedError('kernels are specialised for key_dim=64, value_dim=512, hidden_dim=64')
        for name, (shape, kind) in self._spec.items():
            self._register(name, shape, kind)
        self._pk: Dict[str, tuple] = {}
        self._pk_device = None
        self._h16_cache = None
        if model_weights is not None:
            self.load_weights(model_weights, init_as_zero_if_needed=True)

    # ------------------------------------------------------------------ parameters / checkpoint compat
    def _register(self, name: str, shape, kind: str):
        parts = name.split('.')
        mod = self
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        if kind in ('bn_mean', 'bn_var'):
            mod.register_buffer(parts[-1], torch.zeros(shape) if kind == 'bn_mean' else torch.ones(shape))
        elif kind == 'bn_count':
            mod.register_buffer(parts[-1], torch.zeros(shape, dtype=torch.long))
        else:
            t = torch.zeros(shape)
            if kind in ('conv_w', 'linear_w'):
                fan_in = t[0].numel()
                t.normal_(0, (2.0 / fan_in) ** 0.5)
            elif kind == 'bn_gamma':
                t.fill_(1.0)
            mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))

    def init_hyperparameters(self, config, model_path=None, map_location=None):
        """reference network.py:134-182: read key/value/hidden dims from the checkpoint or the config and
        write them back into `config`."""
        model_weights = None
        if model_path is not None:
            model_weights = torch.load(model_path, map_location=map_location)
            self.key_dim = model_weights['key_proj.key_proj.weight'].shape[0]
            self.value_dim = model_weights['value_encoder.fuser.block2.conv2.weight'].shape[0]
            self.disable_hidden = 'decoder.hidden_update.transform.weight' not in model_weights
            self.hidden_dim = 0 if self.disable_hidden else model_weights['decoder.hidden_update.transform.weight'].shape[0] // 3
        else:
            self.key_dim = config.get('key_dim', 64)
            self.value_dim = config.get('value_dim', 512)
            self.hidden_dim = config.get('hidden_dim', 64)
            self.disable_hidden = self.hidden_dim <= 0
        config['key_dim'] = self.key_dim
        config['value_dim'] = self.value_dim
        config['hidden_dim'] = self.hidden_dim
        return model_weights

    def load_weights(self, src_dict, init_as_zero_if_needed=False):
        """reference network.py:184-198: single-object checkpoints get an extra stem channel."""
        for k in list(src_dict.keys()):
            if k == 'value_encoder.conv1.weight' and src_dict[k].shape[1] == 4:
                pads = torch.zeros((64, 1, 7, 7), device=src_dict[k].device)
                if not init_as_zero_if_needed:
                    nn.init.orthogonal_(pads)
                src_dict[k] = torch.cat([src_dict[k], pads], 1)
        self.load_state_dict(src_dict)
        self._pk = {}

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._pk = {}
        return out

    # ------------------------------------------------------------------ weight packing
    def _w(self, name):
        return self._sd[name].detach().float().cpu()

    def _folded(self, conv, bn=None):
        w = self._w(conv + '.weight')
        b = self._w(conv + '.bias') if (conv + '.bias') in self._spec else None
        if bn is not None:
            w, b = fold_bn(w, b, self._w(bn + '.weight'), self._w(bn + '.bias'), self._w(bn + '.running_mean'),
                           self._w(bn + '.running_var'))
        return w, b

    def _ensure_packed(self, device):
        if self._pk and self._pk_device == device:
            return
        pk = {}
        self._sd = self.state_dict()

        def put(key, w, b, cin_pad=None):
            wp, bp, cout = pack_conv(w, b, cin_pad=cin_pad, device=device)
            pk[key] = (wp, bp, cout, w.shape[-1])

        def put_stem(key, conv, bn, kpad):
            w, b = self._folded(conv, bn)                         # [64, C, 7, 7]
            if w.shape[1] == 3:
                # key stem: k = kh*24 + kw*3 + c (each kernel row padded to 24), matches im2col_stem3_kernel
                rows = torch.zeros((w.shape[0], 8, 24))
                rows[:, :7, :21] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 7, 21)
                flat = rows.reshape(w.shape[0], 192)
            else:
                flat = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)  # k = (kh*7+kw)*C + c, matches xm_im2col_stem
            put(key, flat[:, :, None, None], b, cin_pad=kpad)

        put_stem('key_encoder.conv1', 'key_encoder.conv1', 'key_encoder.bn1', 192)
        put_stem('value_encoder.conv1', 'value_encoder.conv1', 'value_encoder.bn1', 256)
        for name in self._spec:
            if not name.endswith('.weight') or self._spec[name][1] != 'conv_w':
                continue
            conv = name[:-len('.weight')]
            if conv in ('key_encoder.conv1', 'value_encoder.conv1') or conv.startswith('key_proj') or 'SpatialGate' in conv:
                continue
            bn = None
            if conv.startswith(('key_encoder.', 'value_encoder.layer')):
                bn = conv.replace('conv', 'bn') if '.downsample.0' not in conv else conv.replace('downsample.0', 'downsample.1')
            w, b = self._folded(conv, bn)
            put(conv, w, b, cin_pad=320 if conv == 'decoder.hidden_update.g4_conv' else None)
        # HiddenUpdater (modules.py:52-61): g16_conv(g16) + g8_conv(down(g8)) + g4_conv(down(g4)) = ONE 1x1 conv over the channel
        # concatenation [512 | 256 | 257 -> 320] (three sources of the implicit GEMM, one fp32 accumulation, summed bias)
        if not self.disable_hidden and 'decoder.hidden_update.g16_conv.weight' in self._spec:
            w16, b16 = self._folded('decoder.hidden_update.g16_conv'); w8, b8 = self._folded('decoder.hidden_update.g8_conv')
            w4, b4 = self._folded('decoder.hidden_update.g4_conv')
            c16, c8, c4 = w16.shape[1], w8.shape[1], w4.shape[1]
            wf = torch.zeros((w16.shape[0], c16 + c8 + 320, 1, 1))
            wf[:, :c16] = w16; wf[:, c16:c16 + c8] = w8; wf[:, c16 + c8:c16 + c8 + c4] = w4
            put('decoder.hidden_update.g_fused', wf, b16 + b8 + b4)
        # key projection: key | d | e in one GEMM (modules.py:194-211)
        wk, bk = self._folded('key_proj.key_proj'); wd, bd = self._folded('key_proj.d_proj'); we, be = self._folded('key_proj.e_proj')
        put('key_proj', torch.cat([wk, wd, we], 0), torch.cat([bk, bd, be], 0))
        for fz in ('value_encoder.fuser', 'decoder.fuser'):
            a = fz + '.attention.'
            pk[fz + '.cbam'] = tuple(self._w(a + n).contiguous().to(device) for n in (
                'ChannelGate.mlp.1.weight', 'ChannelGate.mlp.1.bias', 'ChannelGate.mlp.3.weight', 'ChannelGate.mlp.3.bias')) + (
                self._w(a + 'SpatialGate.spatial.conv.weight').reshape(-1).contiguous().to(device),
                float(self._w(a + 'SpatialGate.spatial.conv.bias').item()))
        wpred, bpred = self._folded('decoder.pred')
        pk['decoder.pred.direct'] = (wpred[0].permute(1, 2, 0).reshape(9, -1).half().contiguous().to(device), float(bpred.item()))
        self._pk, self._pk_device = pk, device
        self._sd = None

    # ------------------------------------------------------------------ kernel wrappers
    def _conv(self, name, srcs, stride=1, relu=False, residual=None, res_bcast=False, relu_copy=False, want_out=True):
        wp, bp, cout, ks = self._pk[name]
        srcs = [(t, bc) for t, bc in srcs]
        batch = max(t.shape[0] for t, _ in srcs)
        if stride == 2 and batch > 1:
            H, W = srcs[0][0].shape[1:3]
            out = torch.empty((batch, H // 2, W // 2, cout), dtype=torch.float16, device=wp.device)
            for i in range(batch):
                lib.conv2d_nhwc([(t[i:i + 1], False) for t, _ in srcs], wp, bp, cout, ksize=ks, stride=2, relu=relu,
                                residual=None if residual is None else residual[i:i + 1], out=out[i:i + 1])
            return out
        out, out_relu = lib.conv2d_nhwc(srcs, wp, bp, cout, ksize=ks, stride=stride, relu=relu, residual=residual,
                                        residual_broadcast=res_bcast, want_out=want_out, want_relu_copy=relu_copy)
        return (out, out_relu) if relu_copy else out

    @staticmethod
    def _relu(t):
        out = torch.empty_like(t)
        lib.check(lib.load().xm_relu(t.data_ptr(), out.data_ptr(), t.numel(), lib.stream_ptr()), 'xm_relu')
        return out

    def _fusion(self, x, parts, prefix):
        """FeatureFusionBlock (modules.py:22-41): x [1,h,w,1024] shared by all objects, parts: list of [n,h,w,C]."""
        n = parts[0].shape[0]
        bc = n > 1
        raw = [(x, bc)] + [(t, False) for t in parts]
        rel = [(x, bc)] + [(self._relu(t), False) for t in parts]       # x is post-ReLU already (resnet.py:112)
        t1 = self._conv(prefix + '.block1.conv1', rel, relu=True)
        ds = self._conv(prefix + '.block1.downsample', raw)
        g1 = self._conv(prefix + '.block1.conv2', [(t1, False)], residual=ds)
        w1, b1, w2, b2, w7, b7 = self._pk[prefix + '.cbam']
        B, H, W, Cc = g1.shape
        scratch = torch.empty(33 * B * Cc + 2 * B * H * W, dtype=torch.float32, device=g1.device)
        gs = torch.empty_like(g1); gsr = torch.empty_like(g1)
        lib.check(lib.load().xm_cbam(g1.data_ptr(), B, H, W, Cc, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                     w7.data_ptr(), C.c_float(b7), scratch.data_ptr(), gs.data_ptr(), gsr.data_ptr(),
                                     lib.stream_ptr()), 'xm_cbam')
        t2 = self._conv(prefix + '.block2.conv1', [(gsr, False)], relu=True)
        return self._conv(prefix + '.block2.conv2', [(t2, False)], residual=gs)

    def _gru(self, values, h32):
        npix = values.shape[0] * values.shape[1] * values.shape[2]
        h_new = torch.empty_like(h32)
        h16 = torch.empty(h32.shape, dtype=torch.float16, device=h32.device)
        lib.check(lib.load().xm_gru(values.data_ptr(), h32.data_ptr(), npix, self.hidden_dim, h_new.data_ptr(), h16.data_ptr(),
                                    lib.stream_ptr()), 'xm_gru')
        self._h16_cache = (h_new, h16)       # holding h_new keeps its storage from being recycled under the cached pointer
        return h_new

    def _hidden_pair(self, hidden5):
        """hidden [1,n,64,h,w] fp32 (any layout) -> (h32 NHWC [n,h,w,64] fp32 contiguous, h16 same fp16)."""
        h32 = _as_nhwc(hidden5[0], torch.float32)
        c = self._h16_cache
        if c is not None and c[0].data_ptr() == h32.data_ptr() and c[0].shape == h32.shape:
            return h32, c[1]
        return h32, h32.half()

    # ------------------------------------------------------------------ public passes
    def encode_key(self, frame, need_sk=True, need_ek=True):
        """reference network.py:40-70.  frame [1,3,H,W] fp32 (H, W multiples of 16)."""
        if frame.dim() != 4:
            raise NotImplementedError('only b*c*h*w frames are supported on the inference path')
        lib.require_cuda(frame, 'frame')
        B, _, H, W = frame.shape
        if B != 1:
            raise NotImplementedError('batch size 1 (one video stream per call)')
        dev = frame.device
        self._ensure_packed(dev)
        L = lib.load()
        img = frame[0].float().contiguous()
        x = self._stem('key_encoder.conv1', img, None, 1, H, W, relu=True)
        p = torch.empty((1, H // 4, W // 4, 64), dtype=torch.float16, device=dev)
        lib.check(L.xm_maxpool3x3s2(x.data_ptr(), 1, H // 2, W // 2, 64, 0, p.data_ptr(), lib.stream_ptr()), 'xm_maxpool3x3s2')
        x = p
        feats = []
        for layer, blocks, stride in (('res2', 3, 1), ('layer2', 4, 2), ('layer3', 6, 2)):
            for i in range(blocks):
                pre = f'key_encoder.{layer}.{i}'
                s = stride if i == 0 else 1
                o = self._conv(pre + '.conv1', [(x, False)], relu=True)
                o = self._conv(pre + '.conv2', [(o, False)], stride=s, relu=True)
                res = self._conv(pre + '.downsample.0', [(x, False)], stride=s) if (pre + '.downsample.0') in self._pk else x
                x = self._conv(pre + '.conv3', [(o, False)], residual=res, relu=True)
            feats.append(x)
        f4, f8, f16 = feats
        h, w = H // 16, W // 16
        hw = h * w
        proj = self._conv('key_proj', [(f16, False)])                      # [1,h,w,129]
        key = torch.empty((1, h, w, 64), dtype=torch.float16, device=dev)
        sel = torch.empty((1, h, w, 64), dtype=torch.float16, device=dev)
        shr = torch.empty((1, h, w, 1), dtype=torch.float32, device=dev)
        lib.check(L.xm_keyproj_post(proj.data_ptr(), proj.shape[3], hw, hw, key.data_ptr(), sel.data_ptr(), shr.data_ptr(), None, None,
                                    lib.stream_ptr()), 'xm_keyproj_post')
        return (_nchw_view(key), _nchw_view(shr) if need_sk else None, _nchw_view(sel) if need_ek else None,
                _nchw_view(f16), _nchw_view(f8), _nchw_view(f4))

    def _stem(self, name, img, masks, n, H, W, relu):
        """7x7/s2 stem conv + folded bn (+relu) as one kernel (csrc/stem7x7.cu); img fp32 [3,H,W], masks fp32 [n,H,W] or None."""
        wp, bp, cout, _ = self._pk[name]
        assert cout == 64 and wp.shape[0] == 64
        out = torch.empty((n, H // 2, W // 2, 64), dtype=torch.float16, device=img.device)
        lib.check(lib.load().xm_stem7x7(img.data_ptr(), masks.data_ptr() if masks is not None else None, n, H, W, wp.data_ptr(),
                                        bp.data_ptr(), wp.shape[1], 1 if relu else 0, out.data_ptr(), lib.stream_ptr()), 'xm_stem7x7')
        return out

    def encode_value(self, frame, image_feat_f16, h16, masks, is_deep_update=True):
        """reference network.py:72-85 + ValueEncoder.forward modules.py:124-150.  masks [1,n,H,W]."""
        lib.require_cuda(frame, 'frame')
        dev = frame.device
        self._ensure_packed(dev)
        L = lib.load()
        _, n, H, W = masks.shape
        img = frame[0].float().contiguous()
        mk = masks[0].float().contiguous()
        x = self._stem('value_encoder.conv1', img, mk, n, H, W, relu=False)  # conv + bn, no relu yet
        p = torch.empty((n, H // 4, W // 4, 64), dtype=torch.float16, device=dev)
        lib.check(L.xm_maxpool3x3s2(x.data_ptr(), n, H // 2, W // 2, 64, 1, p.data_ptr(), lib.stream_ptr()), 'xm_maxpool3x3s2')
        x = p
        for layer, stride in (('layer1', 1), ('layer2', 2), ('layer3', 2)):
            for i in range(2):
                pre = f'value_encoder.{layer}.{i}'
                s = stride if i == 0 else 1
                o = self._conv(pre + '.conv1', [(x, False)], stride=s, relu=True)
                res = self._conv(pre + '.downsample.0', [(x, False)], stride=s) if (pre + '.downsample.0') in self._pk else x
                x = self._conv(pre + '.conv2', [(o, False)], residual=res, relu=True)
        f16 = _as_nhwc(image_feat_f16)
        g = self._fusion(f16, [x], 'value_encoder.fuser')                   # [n,h,w,512]
        if is_deep_update and self.hidden_dim > 0:
            h32, h16h = self._hidden_pair(h16)
            vals = self._conv('value_encoder.hidden_reinforce.transform', [(g, False), (h16h, False)])
            h16 = self._gru(vals, h32).permute(0, 3, 1, 2).unsqueeze(0)
        return g.permute(0, 3, 1, 2).unsqueeze(0), h16

    def segment(self, multi_scale_features, memory_readout, hidden_state, selector=None, h_out=True, strip_bg=True):
        """reference network.py:107-120 + Decoder.forward modules.py:229-250.
        memory_readout [1,n,512,h,w], hidden_state [1,n,64,h,w] -> (hidden, logits, prob)."""
        if selector is not None:
            raise NotImplementedError('selector is a training-time argument')
        f16, f8, f4 = [_as_nhwc(t) for t in multi_scale_features]
        lib.require_cuda(f16, 'features')
        dev = f16.device
        self._ensure_packed(dev)
        L = lib.load()
        ro = _as_nhwc(memory_readout[0])
        n, h, w, _ = ro.shape
        h32, h16h = self._hidden_pair(hidden_state)
        g16 = self._fusion(f16, [ro, h16h], 'decoder.fuser')

        def up_block(skip_feat, up_g, pre, last):
            skip = self._conv(pre + '.skip_conv', [(skip_feat, False)])
            B, hh, ww, Cc = up_g.shape
            g = torch.empty((B, 2 * hh, 2 * ww, Cc), dtype=torch.float16, device=dev); gr = torch.empty_like(g)
            lib.check(L.xm_upsample2x_add(up_g.data_ptr(), skip.data_ptr(), B, hh, ww, Cc, g.data_ptr(), gr.data_ptr(), lib.stream_ptr()),
                      'xm_upsample2x_add')
            t = self._conv(pre + '.out_conv.conv1', [(gr, False)], relu=True)
            res = self._conv(pre + '.out_conv.downsample', [(g, False)]) if (pre + '.out_conv.downsample') in self._pk else g
            return self._conv(pre + '.out_conv.conv2', [(t, False)], residual=res, relu_copy=last)

        g8 = up_block(f8, g16, 'decoder.up_16_8', False)
        g4, g4r = up_block(f4, g8, 'decoder.up_8_4', True)
        wpd, bpd = self._pk['decoder.pred.direct']
        logits4 = torch.empty((n, 4 * h, 4 * w, 1), dtype=torch.float16, device=dev)
        lib.check(L.xm_conv3x3_c1(g4r.data_ptr(), wpd.data_ptr(), C.c_float(bpd), n, 4 * h, 4 * w, g4r.shape[3], logits4.data_ptr(),
                                  lib.stream_ptr()), 'xm_conv3x3_c1')
        new_hidden = None
        if h_out and self.hidden_dim > 0:
            g8d = torch.empty((n, h, w, 256), dtype=torch.float16, device=dev)
            lib.check(L.xm_area_down(g8.data_ptr(), None, n, 2 * h, 2 * w, 256, 2, 256, g8d.data_ptr(), lib.stream_ptr()), 'xm_area_down')
            g4d = torch.empty((n, h, w, 320), dtype=torch.float16, device=dev)
            lib.check(L.xm_area_down(g4.data_ptr(), logits4.data_ptr(), n, 4 * h, 4 * w, 256, 4, 320, g4d.data_ptr(), lib.stream_ptr()),
                      'xm_area_down')
            c = self._conv('decoder.hidden_update.g_fused', [(g16, False), (g8d, False), (g4d, False)])
            vals = self._conv('decoder.hidden_update.transform', [(c, False), (h16h, False)])
            new_hidden = self._gru(vals, h32).permute(0, 3, 1, 2).unsqueeze(0)
        H, W = 16 * h, 16 * w
        prob = torch.empty((n + 1, H, W), dtype=torch.float32, device=dev)
        logits = torch.empty((n + 1, H, W), dtype=torch.float32, device=dev)
        lib.check(L.xm_upsample4x_aggregate(logits4.data_ptr(), n, 4 * h, 4 * w, prob.data_ptr(), logits.data_ptr(), lib.stream_ptr()),
                  'xm_upsample4x_aggregate')
        prob = prob.unsqueeze(0); logits = logits.unsqueeze(0)
        if strip_bg:
            prob = prob[:, 1:]
        return new_hidden, logits, prob

    def forward(self, mode, *args, **kwargs):
        if mode == 'encode_key':
            return self.encode_key(*args, **kwargs)
        if mode == 'encode_value':
            return self.encode_value(*args, **kwargs)
        if mode == 'segment':
            return self.segment(*args, **kwargs)
        raise NotImplementedError(mode)
