"""Checkpoint compatibility of `XMem` on CPU (reference model/network.py:134-198): hyper-parameters are read from a
checkpoint file and written back into the config; single-object checkpoints (4-channel value stem) are widened to the
multi-object layout; unknown layouts are rejected loudly."""
import os
import tempfile

import pytest
import torch

from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict


def test_hyperparameters_come_from_the_checkpoint_file():
    sd = synth_state_dict(0)
    path = os.path.join(tempfile.mkdtemp(), 'XMem.pth')
    torch.save(sd, path)
    cfg = {'key_dim': 1, 'value_dim': 2, 'hidden_dim': 3}            # must be overwritten from the weights (network.py:146-180)
    net = XMem(cfg, path, map_location='cpu')
    assert (cfg['key_dim'], cfg['value_dim'], cfg['hidden_dim']) == (64, 512, 64)
    for k, v in net.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_single_object_checkpoint_is_widened():
    sd = synth_state_dict(0)
    sd['value_encoder.conv1.weight'] = sd['value_encoder.conv1.weight'][:, :4].clone()
    net = XMem({}, None)
    net.load_weights(dict(sd), init_as_zero_if_needed=True)
    w = net.state_dict()['value_encoder.conv1.weight']
    assert w.shape == (64, 5, 7, 7) and torch.equal(w[:, :4], sd['value_encoder.conv1.weight']) and float(w[:, 4].abs().sum()) == 0.0


def test_unsupported_dimensions_are_rejected():
    with pytest.raises(NotImplementedError):
        XMem({'key_dim': 32}, None)
    with pytest.raises(NotImplementedError):
        XMem({'single_object': True}, None)


def test_packing_layouts_on_cpu():
    # BN folding + K-order of the packed weights, checked without a GPU against torch convolutions on small inputs
    import torch.nn.functional as F
    net = XMem({}, None)
    sd = synth_state_dict(0)
    net.load_weights(dict(sd))
    net._ensure_packed(torch.device('cpu'))
    pk = net._pk
    # (1) a BN-folded 3x3 conv: y = conv(x, w_folded) + b_folded must equal bn(conv(x, w))
    name = 'key_encoder.res2.0.conv2'
    wp, bp, cout, ks = pk[name]
    cin = sd[name + '.weight'].shape[1]
    x = torch.randn(1, cin, 6, 7)
    w_fold = wp.float().view(wp.shape[0], ks * ks, -1)[:cout, :, :cin].reshape(cout, ks, ks, cin).permute(0, 3, 1, 2)
    y = F.conv2d(x, w_fold, bp[:cout], padding=1)
    bn = name.replace('conv', 'bn')
    ref = F.batch_norm(F.conv2d(x, sd[name + '.weight'], None, padding=1), sd[bn + '.running_mean'], sd[bn + '.running_var'],
                       sd[bn + '.weight'], sd[bn + '.bias'], False, 0.0, 1e-5)
    assert torch.allclose(y, ref, atol=2e-2, rtol=2e-2)
    # (2) key projection: rows 0..63 key, 64 shrinkage pre-activation, 65..128 selection pre-activation
    wp, bp, cout, ks = pk['key_proj']
    assert cout == 129 and wp.shape[0] == 192
    wk = wp.float().view(192, 9, 1024)
    assert torch.allclose(wk[64, :, :].reshape(3, 3, 1024).permute(2, 0, 1), sd['key_proj.d_proj.weight'][0], atol=1e-3)
    assert torch.allclose(wk[65, :, :].reshape(3, 3, 1024).permute(2, 0, 1), sd['key_proj.e_proj.weight'][0], atol=1e-3)
    # (3) key stem: row-padded layout k = kh*24 + kw*3 + c
    wp, bp, cout, ks = pk['key_encoder.conv1']
    assert wp.shape == (64, 192) and ks == 1
    assert float(wp[:, 21:24].abs().sum()) == 0.0 and float(wp[:, 168:].abs().sum()) == 0.0
    # (4) hidden-update 1x1 on 257 channels is padded to 320
    assert pk['decoder.hidden_update.g4_conv'][0].shape == (256, 320)
    assert float(pk['decoder.hidden_update.g4_conv'][0][:, 257:].abs().sum()) == 0.0
