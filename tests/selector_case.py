"""Seeded inputs of the annotation-candidate selector cases (shared by tests/golden/make_golden.py, which runs the live
reference on them, and by the tests that replay the committed fixture tests/golden/selector.npz)."""
import torch


def selector_inputs():
    """seeded keys / shrinkage / selection of 14 'frames' on a 6x9 grid and full-resolution masks (some too small, some
    with two objects); shared by the fixture generator and the tests."""
    g = torch.Generator().manual_seed(11)
    N, h, w, H, W = 14, 6, 9, 96, 144
    base = torch.randn(1, 64, h, w, generator=g) * 0.6
    drift = torch.cumsum(torch.randn(N, 64, h, w, generator=g) * (0.15 + 0.25 * torch.rand(N, 1, 1, 1, generator=g)), dim=0)
    keys = (base + drift).half().float()
    shr = (torch.rand(N, 1, h, w, generator=g) + 1).float()
    sel = torch.rand(N, 64, h, w, generator=g).half().float()
    yy = torch.arange(H).view(H, 1); xx = torch.arange(W).view(1, W)
    masks = []
    for i in range(N):
        r = 3 if i in (4, 9) else 150 + 40 * i                      # frames 4 and 9: mask below the presence threshold
        m = (((yy - 30 - 2 * i) ** 2 + (xx - 40 - 5 * i) ** 2) <= r).float().unsqueeze(0)
        if i % 5 == 2:
            m = torch.cat([m, (((yy - 70) ** 2 + (xx - 110) ** 2) <= 120).float().unsqueeze(0)], 0)
        masks.append(m)
    return keys, shr, sel, masks


SELECTOR_CASES = [dict(alpha=0.5, prev=[0], k=5), dict(alpha=1.0, prev=[0, 6], k=4), dict(alpha=0.0, prev=[3], k=6),
                  dict(alpha=0.7, prev=[13, 1, 7], k=3)]


def pair_scores_via_score_dump(packed, chosen, candidates):
    """Test-side reference of the fused pair kernel (csrc/pair_dissim.cu): the same quantity composed from two full
    read-kernel calls with the similarity dump (debug_scores) per ordered pair, as round 1 computed it."""
    import ctypes as C
    import torch
    import torch.nn.functional as F
    from xmem2_b200 import lib
    hw, hw_pad, cap = packed.hw, packed.hw_pad, packed.cap
    dev = next(iter(packed.frames.values()))[0].device
    values = torch.zeros((1, lib.CV, cap), dtype=torch.float16, device=dev)
    readout = torch.empty((1, hw, lib.CV), dtype=torch.float16, device=dev)
    ws = lib.affinity_workspace(hw, 1, dev)
    s_ab, s_ba = [torch.empty((hw, hw_pad), dtype=torch.float32, device=dev) for _ in range(2)]
    top_k = min(30, hw)

    def similarity(mem, query, out):
        kp, ms, _, _ = packed.frames[mem]
        _, _, qp, bsq = packed.frames[query]
        a = lib.XmAffinityArgs()
        a.banks[0].size = 0; a.banks[2].size = 0
        b = a.banks[1]
        b.keys, b.shrinkage, b.values, b.usage = kp.data_ptr(), ms.data_ptr(), values.data_ptr(), None
        b.cap, b.n_obj_cap, b.size = cap, 1, hw
        a.n_groups = 1
        a.groups[0].obj_begin, a.groups[0].n_obj = 0, 1
        a.qp, a.bsq, a.hw, a.hw_pad, a.top_k, a.n_obj_total = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad, top_k, 1
        a.readout_chw, a.readout_hwc = None, readout.data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.debug_scores = out.data_ptr()
        a.plan_is_resident = 0
        lib.check(lib.load().xm_affinity_readout(C.byref(a), lib.stream_ptr()), 'xm_affinity_readout')

    out = torch.empty(len(candidates), dtype=torch.float32, device=dev)
    for n, j in enumerate(candidates):
        similarity(chosen, j, s_ab)
        similarity(j, chosen, s_ba)
        out[n] = F.relu(s_ab[:, :hw] - s_ba[:, :hw]).sum() / (hw * hw)
    return out
