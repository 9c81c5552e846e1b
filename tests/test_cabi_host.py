"""CPU-only checks of the C-ABI library: it loads without a GPU, exports every symbol declared in
include/xmem2_b200.h, and its host-side planning logic (xm_affinity_plan: column ranges of object groups over the
three banks, reference inference/memory_manager.py:98-128,162-182) behaves as specified.  No kernels run here."""
import ctypes as C
import os
import re

import pytest

from xmem2_b200 import lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    header = open(os.path.join(REPO, 'include', 'xmem2_b200.h')).read()
    names = set(re.findall(r'\b(xm_[a-z0-9_]+)\s*\(', header))
    assert len(names) >= 18
    for n in sorted(names):
        assert hasattr(L, n), f'{n} declared in the header but not exported'
    assert L.xm_version() >= 100


def _args(sizes, groups, top_k=30, n_obj=2):
    a = lib.XmAffinityArgs()
    for i, n in enumerate(sizes):
        b = a.banks[i]
        b.size = n
        if n:
            b.keys, b.shrinkage, b.values, b.usage = 0x1000 * (i + 1), 0x2000 * (i + 1), 0x3000 * (i + 1), 0x4000 * (i + 1)
            b.cap, b.n_obj_cap = (n + 71) // 8 * 8, n_obj
    a.n_groups = len(groups)
    for gi, (ob, no, begins) in enumerate(groups):
        a.groups[gi].obj_begin, a.groups[gi].n_obj = ob, no
        for bi in range(3):
            a.groups[gi].begin[bi] = begins[bi]
    a.top_k, a.n_obj_total = top_k, n_obj
    return a


def test_affinity_plan_is_host_only_and_validates():
    L = lib.load()
    buf = (C.c_uint8 * 4096)()
    a = _args((128, 300, 405), [(0, 1, [0, 0, 0]), (1, 1, [128, 150, 135])])
    assert L.xm_affinity_plan(C.byref(a), buf, 4096) == 0
    ints = (C.c_int32 * 1024).from_buffer(buf)
    assert ints[0] == 3                               # group 0 reads three column ranges (long, working, permanent)
    # a group that sees fewer columns than top_k must be rejected like torch.topk would
    a = _args((0, 0, 20), [(0, 1, [0, 0, 0])], n_obj=1)
    assert L.xm_affinity_plan(C.byref(a), buf, 4096) != 0
    assert b'top_k' in L.xm_last_error()
    # begin outside the bank
    a = _args((0, 100, 100), [(0, 1, [0, 101, 0])], n_obj=1)
    assert L.xm_affinity_plan(C.byref(a), buf, 4096) != 0
    # too small an output buffer
    a = _args((0, 100, 100), [(0, 1, [0, 0, 0])], n_obj=1)
    assert L.xm_affinity_plan(C.byref(a), buf, 100) != 0


def test_entry_points_reject_bad_arguments_without_touching_the_gpu():
    L = lib.load()
    assert L.xm_query_pack(None, None, 10, 128, None, None, None) != 0
    assert L.xm_conv2d_nhwc(None, None) != 0
    assert L.xm_affinity_workspace_bytes(1620, 1) > 4096


def test_missing_cuda_is_a_loud_error():
    import torch
    from xmem2_b200.model.network import XMem
    net = XMem({}, None)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net.encode_key(torch.zeros(1, 3, 32, 32))
