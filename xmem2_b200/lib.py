"""ctypes binding of libxmem2_b200.so (the C ABI declared in include/xmem2_b200.h).

There is no CPU or library fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  PyTorch is used only to own device memory and the CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libxmem2_b200.so')

XM_MAX_GROUPS = 8
XM_MAX_TOPK = 32
CK = 64
CV = 512


class XmBank(C.Structure):
    _fields_ = [('keys', C.c_void_p), ('shrinkage', C.c_void_p), ('values', C.c_void_p), ('usage', C.c_void_p),
                ('cap', C.c_int64), ('n_obj_cap', C.c_int32), ('size', C.c_int32)]


class XmGroup(C.Structure):
    _fields_ = [('obj_begin', C.c_int32), ('n_obj', C.c_int32), ('begin', C.c_int32 * 3)]


class XmAffinityArgs(C.Structure):
    _fields_ = [('banks', XmBank * 3), ('n_groups', C.c_int32), ('groups', XmGroup * XM_MAX_GROUPS),
                ('qp', C.c_void_p), ('bsq', C.c_void_p), ('hw', C.c_int32), ('hw_pad', C.c_int32),
                ('top_k', C.c_int32), ('n_obj_total', C.c_int32), ('readout_chw', C.c_void_p),
                ('readout_hwc', C.c_void_p), ('workspace', C.c_void_p), ('workspace_bytes', C.c_int64),
                ('debug_scores', C.c_void_p), ('plan_is_resident', C.c_int32)]


class XmConvSrc(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('channels', C.c_int32), ('broadcast', C.c_int32)]


class XmConvArgs(C.Structure):
    _fields_ = [('src', XmConvSrc * 3), ('n_src', C.c_int32), ('batch', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('ksize', C.c_int32), ('stride', C.c_int32), ('weight', C.c_void_p), ('bias', C.c_void_p),
                ('cout', C.c_int32), ('cout_pad', C.c_int32), ('residual', C.c_void_p), ('residual_broadcast', C.c_int32),
                ('relu', C.c_int32), ('out', C.c_void_p), ('out_relu', C.c_void_p), ('out_stride', C.c_int32),
                ('out_offset', C.c_int32), ('workspace', C.c_void_p), ('workspace_bytes', C.c_int64)]


_conv_ws = {}


def conv_workspace(device):
    """per-device split-K scratch (zeroed counters + partial tiles); allocated once."""
    key = str(device)
    if key not in _conv_ws:
        _conv_ws[key] = torch.zeros(65536 * 4 + 48 * 1024 * 1024, dtype=torch.uint8, device=device)
    return _conv_ws[key]


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: build it with `python -m xmem2_b200.build` '
                               '(there is no CPU fallback for the XMem++ hot path)')
        lib = C.CDLL(LIB_PATH)
        lib.xm_last_error.restype = C.c_char_p
        lib.xm_version.restype = C.c_int
        lib.xm_affinity_workspace_bytes.restype = C.c_int64
        lib.xm_affinity_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
        lib.xm_affinity_workspace_init.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
        lib.xm_affinity_readout.argtypes = [C.POINTER(XmAffinityArgs), C.c_void_p]
        lib.xm_affinity_plan.argtypes = [C.POINTER(XmAffinityArgs), C.c_void_p, C.c_int64]
        lib.xm_query_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.xm_key_pack.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.xm_conv2d_nhwc.argtypes = [C.POINTER(XmConvArgs), C.c_void_p]
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        lib.xm_debug_last_trap.argtypes = [C.POINTER(C.c_int)]
        lib.xm_launch_count.restype = C.c_longlong
        lib.xm_add_launch_count.argtypes = [C.c_int]
        lib.xm_im2col_stem.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
        lib.xm_stem7x7.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, i32, vp, vp]
        lib.xm_maxpool3x3s2.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
        lib.xm_relu.argtypes = [vp, vp, i64, vp]
        lib.xm_keyproj_post.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
        lib.xm_cbam.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]
        lib.xm_upsample2x_add.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp]
        lib.xm_area_down.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
        lib.xm_gru.argtypes = [vp, vp, i64, i32, vp, vp, vp]
        lib.xm_conv3x3_c1.argtypes = [vp, vp, f32, i32, i32, i32, i32, vp, vp]
        lib.xm_upsample4x_aggregate.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        lib.xm_resize_argmax.argtypes = [vp, i32, i32, i32, i64, i64, i32, i32, vp, vp, vp]
        lib.xm_value_append.argtypes = [vp, i32, i32, vp, i64, i32, vp]
        lib.xm_pair_dissimilarity.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, i32, vp, vp, vp]
        lib.xm_usage_topk.argtypes = [vp, vp, i32, i32, vp, vp]
        lib.xm_usage_evict_list.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        lib.xm_consolidate_affinity.argtypes = [vp, vp, vp, i32, vp, i32, i32, vp, i64, vp, vp]
        lib.xm_consolidate_scratch_bytes.restype = C.c_int64
        lib.xm_consolidate_scratch_bytes.argtypes = [i32, i32]
        lib.xm_consolidate_values.argtypes = [vp, i64, i32, i32, i32, vp, i64, vp, i32, vp, i64, vp, vp]
        lib.xm_bank_compact_tmp_bytes.restype = C.c_int64
        lib.xm_bank_compact_tmp_bytes.argtypes = [i32, i32]
        lib.xm_bank_compact.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, vp, i32, i32, i32, vp, i64, vp]
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f'{what} failed ({rc}): {load().xm_last_error().decode()}')


def last_trap():
    """{tag, block, thread, parity} of the last in-kernel mbarrier timeout, or None (diagnostics)."""
    buf = (C.c_int * 7)()
    if load().xm_debug_last_trap(buf):
        return dict(tag=buf[1], block=(buf[2], buf[3], buf[4]), thread=buf[5], parity=buf[6])
    return None


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must live on a CUDA device: the XMem++ hot path has no CPU fallback')


# ------------------------------------------------------------------------------------------------
# thin typed wrappers
# ------------------------------------------------------------------------------------------------
def affinity_workspace(hw: int, n_obj: int, device) -> torch.Tensor:
    """Workspace of the fused read kernel for `hw` query positions and `n_obj` value planes.  Its inter-CTA barrier counters
    have to be zero before the first launch (the kernel re-arms them itself on exit): xm_affinity_workspace_init."""
    wsb = load().xm_affinity_workspace_bytes(hw, n_obj)
    ws = torch.empty(wsb, dtype=torch.uint8, device=device)
    with torch.cuda.device(ws.device):
        check(load().xm_affinity_workspace_init(ws.data_ptr(), wsb, hw, n_obj, stream_ptr()), 'xm_affinity_workspace_init')
    return ws


def query_pack(key_hwc: torch.Tensor, sel_hwc: torch.Tensor, hw_pad: int):
    """key/sel [hw,64] fp16 contiguous -> (qp [hw_pad,128] fp16, bsq [hw_pad] fp32)."""
    require_cuda(key_hwc, 'key')
    hw = key_hwc.shape[0]
    assert key_hwc.dtype == torch.float16 and sel_hwc.dtype == torch.float16
    assert key_hwc.is_contiguous() and sel_hwc.is_contiguous() and key_hwc.shape == sel_hwc.shape == (hw, CK)
    qp = torch.empty((hw_pad, 2 * CK), dtype=torch.float16, device=key_hwc.device)
    bsq = torch.empty((hw_pad,), dtype=torch.float32, device=key_hwc.device)
    check(load().xm_query_pack(ptr(key_hwc), ptr(sel_hwc), hw, hw_pad, ptr(qp), ptr(bsq), stream_ptr()), 'xm_query_pack')
    return qp, bsq


def key_pack(key_hwc: torch.Tensor, dst_rows: torch.Tensor):
    """key [n,64] fp16 -> dst_rows [n,128] fp16 (a contiguous row slice of a bank's packed-key arena)."""
    require_cuda(key_hwc, 'key')
    n = key_hwc.shape[0]
    assert key_hwc.dtype == torch.float16 and key_hwc.is_contiguous() and dst_rows.is_contiguous()
    assert dst_rows.shape == (n, 2 * CK) and dst_rows.dtype == torch.float16
    check(load().xm_key_pack(ptr(key_hwc), n, ptr(dst_rows), stream_ptr()), 'xm_key_pack')


def conv2d_nhwc(srcs, weight, bias, cout, ksize=3, stride=1, relu=False, residual=None, residual_broadcast=False,
                out=None, out_relu=None, out_offset=0, want_out=True, want_relu_copy=False):
    """srcs: list of (tensor [B,H,W,C] fp16, broadcast flag).  weight fp16 [cout_pad, k*k*cin], bias fp32 [cout_pad].
    Returns (out, out_relu) NHWC fp16."""
    t0 = srcs[0][0]
    require_cuda(t0, 'conv input')
    batch = max(t.shape[0] for t, _ in srcs)
    H, W = t0.shape[1], t0.shape[2]
    Ho, Wo = H // stride, W // stride
    a = XmConvArgs()
    a.n_src = len(srcs)
    for i, (t, bc) in enumerate(srcs):
        assert t.dtype == torch.float16 and t.is_contiguous() and t.shape[1] == H and t.shape[2] == W
        a.src[i].ptr = ptr(t); a.src[i].channels = t.shape[3]; a.src[i].broadcast = 1 if bc else 0
    a.batch, a.H, a.W, a.ksize, a.stride = batch, H, W, ksize, stride
    a.weight, a.bias, a.cout, a.cout_pad = ptr(weight), ptr(bias), cout, weight.shape[0]
    a.residual = ptr(residual); a.residual_broadcast = 1 if residual_broadcast else 0
    a.relu = 1 if relu else 0
    if out is None and want_out:
        out = torch.empty((batch, Ho, Wo, cout), dtype=torch.float16, device=t0.device)
    if out_relu is None and want_relu_copy:
        out_relu = torch.empty((batch, Ho, Wo, cout), dtype=torch.float16, device=t0.device)
    ref = out if out is not None else out_relu
    a.out, a.out_relu = ptr(out), ptr(out_relu)
    a.out_stride, a.out_offset = ref.shape[3], out_offset
    ws = conv_workspace(t0.device)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    check(load().xm_conv2d_nhwc(C.byref(a), stream_ptr()), 'xm_conv2d_nhwc')
    return out, out_relu


# ------------------------------------------------------------------------------------------------
# long-term memory maintenance (csrc/consolidate.cu); tensor-level wrappers so that the CPU host-logic tests can replace them
# ------------------------------------------------------------------------------------------------
def usage_topk(use: torch.Tensor, life: torch.Tensor, k: int) -> torch.Tensor:
    """indices (int32 [k]) of the k largest use/life, descending, ties by ascending index (memory_manager.py:355)."""
    require_cuda(use, 'usage')
    n = use.numel()
    out = torch.empty(k, dtype=torch.int32, device=use.device)
    check(load().xm_usage_topk(ptr(use), ptr(life), n, k, ptr(out), stream_ptr()), 'xm_usage_topk')
    return out


def usage_evict_list(use: torch.Tensor, life: torch.Tensor, n: int, n_remove: int):
    """(keep_idx int32 [n] of which the first `count` are valid, count): columns with use/life above the n_remove-th smallest."""
    keep = torch.empty(n, dtype=torch.int32, device=use.device)
    count = torch.zeros(1, dtype=torch.int32, device=use.device)
    check(load().xm_usage_evict_list(ptr(use), ptr(life), n, n_remove, ptr(keep), ptr(count), stream_ptr()), 'xm_usage_evict_list')
    return keep, int(count.item())


def consolidate_affinity(kp, s, e, proto, col_begin: int, aff, shr_out):
    """aff[q, col_begin:n] <- softmax_n similarity(candidate n, prototype q) for prototypes with proto[q] >= col_begin; shr_out optional."""
    n = kp.shape[0]
    check(load().xm_consolidate_affinity(ptr(kp), ptr(s), ptr(e), n, ptr(proto), proto.numel(), col_begin, ptr(aff), aff.stride(0),
                                         ptr(shr_out), stream_ptr()), 'xm_consolidate_affinity')


def consolidate_values(gv, aff, col_begin: int, valid, n_valid: int) -> torch.Tensor:
    """gv [n_g, CV, N_g] fp16 (last dim contiguous, channel pitch gv.stride(1)) @ aff[valid][:, col_begin:col_begin+N_g]^T -> [n_g, CV, n_valid]."""
    L = load()
    n_g, _, ng = gv.shape
    sb = L.xm_consolidate_scratch_bytes(n_g, n_valid)
    scratch = torch.empty(sb, dtype=torch.uint8, device=gv.device)
    out = torch.empty((n_g, CV, n_valid), dtype=torch.float16, device=gv.device)
    check(L.xm_consolidate_values(ptr(gv), gv.stride(1), n_g, 0, ng, aff.data_ptr() + 4 * col_begin, aff.stride(0), ptr(valid), n_valid,
                                  ptr(scratch), sb, ptr(out), stream_ptr()), 'xm_consolidate_values')
    return out


def bank_compact(kp, s, e, use, life, v, keep_idx, shift: int, first: int, m: int):
    """columns [first, m) of every arena of a bank <- columns keep_idx[i] (or i + shift), in place."""
    L = load()
    n_planes, cap = v.shape[0], v.shape[2]
    tb = L.xm_bank_compact_tmp_bytes(n_planes, m - first)
    tmp = torch.empty(tb, dtype=torch.uint8, device=kp.device)
    check(L.xm_bank_compact(ptr(kp), ptr(s), ptr(e), ptr(use), ptr(life), ptr(v), cap, n_planes, ptr(keep_idx), shift, first, m, ptr(tmp), tb,
                            stream_ptr()), 'xm_bank_compact')

