"""The round-2 head-start kernels (xmem2_b200/csrc/experimental/) must keep compiling for sm_100a and exporting their
entry points, and the development switch XMEM_CONV_IMPL must leave the product path alone unless it is set.  No kernel of
that library has run on a GPU yet; nothing here launches one."""
import ctypes as C

import pytest

from xmem2_b200 import build as xb
from xmem2_b200 import lib


def test_experimental_library_builds_and_exports_its_entry_points():
    path = xb.build_experimental()
    L = C.CDLL(path)
    for name in ('xm_conv2d_nhwc_csk', 'xm_conv2d_nhwc_2cta', 'xm_conv2d_nhwc_mc', 'xm_conv2d_nhwc_halo', 'xm_pair_dissimilarity',
                 'xm_last_error'):
        assert hasattr(L, name), name
    L.xm_last_error.restype = C.c_char_p
    # argument checks run before any CUDA call
    for name in ('xm_conv2d_nhwc_csk', 'xm_conv2d_nhwc_2cta', 'xm_conv2d_nhwc_mc', 'xm_conv2d_nhwc_halo'):
        assert getattr(L, name)(None, None) != 0
        assert b'null args' in L.xm_last_error()
    assert L.xm_pair_dissimilarity(None, None, None, None, 1, 54, 128, None, None, 1, None, None, None) != 0
    assert b'null pointer' in L.xm_last_error()


def test_switch_is_off_by_default_and_routes_only_supported_shapes(monkeypatch):
    assert lib._CONV_IMPL == ''                      # the validated configuration: every convolution -> product library

    class Fake:
        calls = []

        def __getattr__(self, name):
            def f(a, stream):
                Fake.calls.append(name)
                return 0
            return f
    monkeypatch.setattr(lib, 'load_experimental', lambda: Fake())
    monkeypatch.setattr(lib, 'stream_ptr', lambda: None)
    a = lib.XmConvArgs()
    a.cout_pad = 64
    monkeypatch.setattr(lib, '_CONV_IMPL', '2cta')
    assert lib._experimental_conv(a) is False        # 64-channel layers stay on the product kernel
    a.cout_pad = 256
    assert lib._experimental_conv(a) is True and Fake.calls == ['xm_conv2d_nhwc_2cta']
    monkeypatch.setattr(lib, '_CONV_IMPL', 'csk')
    a.cout_pad = 64
    assert lib._experimental_conv(a) is True and Fake.calls[-1] == 'xm_conv2d_nhwc_csk'
    monkeypatch.setattr(lib, '_CONV_IMPL', 'halo')
    a.ksize, a.stride = 1, 1
    assert lib._experimental_conv(a) is False        # 1x1 layers stay on the product kernel
    a.ksize = 3
    assert lib._experimental_conv(a) is True and Fake.calls[-1] == 'xm_conv2d_nhwc_halo'
    monkeypatch.setattr(lib, '_CONV_IMPL', 'typo')
    with pytest.raises(RuntimeError, match='expected csk'):
        lib._experimental_conv(a)
