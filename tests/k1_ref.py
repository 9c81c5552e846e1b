"""Test helpers for the fused memory-read kernel (K1): build random banks in the kernel's arena layout,
run the C ABI, and compute the expected result with the oracle's attention math on the same
fp16-rounded operands (oracle/xmem_oracle.py: similarity + softmax_topk)."""
import ctypes as C
import torch

from xmem2_b200 import lib

CK, CV = 64, 512


def make_case(hw, sizes, n_obj, group_begins, seed=0, device='cuda', key_scale=0.6):
    """sizes = (long, work, perm) columns; group_begins = list of (obj_begin, n_obj, [b_long, b_work, b_perm])."""
    g = torch.Generator().manual_seed(seed)
    banks = []
    for n in sizes:
        cap = max(8, (n + 64 + 7) // 8 * 8)
        if n == 0:
            banks.append(None); continue
        key = (torch.randn(n, CK, generator=g) * key_scale).half()
        shr = (torch.rand(n, generator=g) * 2 + 1).float()
        val = torch.zeros(n_obj, CV, cap).half()
        val[:, :, :n] = torch.randn(n_obj, CV, n, generator=g).half()
        banks.append(dict(key=key, shr=shr, val=val, cap=cap, n=n))
    qk = (torch.randn(hw, CK, generator=g) * key_scale).half()
    qe = torch.rand(hw, CK, generator=g).half()
    return dict(hw=hw, banks=banks, n_obj=n_obj, groups=group_begins, qk=qk, qe=qe, device=device)


def expected(case, top_k=30):
    """fp64 evaluation of memory_util.py:7-65 on the fp16-rounded operands the kernel consumes."""
    qk, qe = case['qk'].double(), case['qe'].double()
    keh = (case['qk'] * case['qe'])          # fp16 product like the kernel / autocast
    two_ke = (keh + keh).double()
    bsq = (case['qe'].float() * case['qk'].float() ** 2).sum(1).double()
    sims = []
    for b in case['banks']:
        if b is None:
            sims.append(None); continue
        k = b['key']
        k2 = (k.float() ** 2).half().double()
        sp = -(k2 @ qe.t()) + k.double() @ two_ke.t()            # [n, hw]
        s = (sp - bsq[None, :]) * b['shr'].double()[:, None] / 8.0
        sims.append(s)
    outs, usages, ambiguous = [], None, torch.zeros(case['hw'], dtype=torch.bool)
    for gi, (ob, no, begins) in enumerate(case['groups']):
        cols, vals = [], []
        for bi, b in enumerate(case['banks']):
            if b is None or begins[bi] >= b['n']:
                continue
            cols.append(sims[bi][begins[bi]:])
            vals.append(b['val'][ob:ob + no, :, begins[bi]:b['n']].double())
        s = torch.cat(cols, 0)                                    # [Ng, hw]
        v = torch.cat(vals, 2)                                    # [no, CV, Ng]
        tv, ti = torch.topk(s, min(top_k + 1, s.shape[0]), dim=0)
        if s.shape[0] > top_k:
            ambiguous |= (tv[top_k - 1] - tv[top_k]) < 3e-5
        e = tv[:top_k].exp()
        w = e / e.sum(0, keepdim=True)
        aff = torch.zeros_like(s).scatter_(0, ti[:top_k], w)
        outs.append(torch.einsum('ocn,nq->ocq', v, aff.half().double()))
        if gi == 0:
            usages = aff.sum(1)
            scores0 = s
    return torch.cat(outs, 0), usages, ambiguous, scores0


def run_kernel(case, top_k=30, want_debug=False):
    dev = case['device']
    hw = case['hw']; hw_pad = (hw + 127) // 128 * 128
    L = lib.load()
    a = lib.XmAffinityArgs()
    keep = []
    usage_bufs = []
    for bi, b in enumerate(case['banks']):
        if b is None:
            a.banks[bi].size = 0; usage_bufs.append(None); continue
        rows = torch.zeros(b['cap'], 2 * CK, dtype=torch.float16, device=dev)
        lib.key_pack(b['key'].to(dev).contiguous(), rows[:b['n']])
        shr = torch.ones(b['cap'], dtype=torch.float32, device=dev); shr[:b['n']] = b['shr'].to(dev)
        val = b['val'].to(dev).contiguous()
        usage = torch.zeros(b['cap'], dtype=torch.float32, device=dev)
        keep += [rows, shr, val, usage]; usage_bufs.append(usage)
        bk = a.banks[bi]
        bk.keys, bk.shrinkage, bk.values, bk.usage = rows.data_ptr(), shr.data_ptr(), val.data_ptr(), usage.data_ptr()
        bk.cap, bk.n_obj_cap, bk.size = b['cap'], case['n_obj'], b['n']
    a.n_groups = len(case['groups'])
    for gi, (ob, no, begins) in enumerate(case['groups']):
        a.groups[gi].obj_begin, a.groups[gi].n_obj = ob, no
        for bi in range(3):
            a.groups[gi].begin[bi] = begins[bi]
    qp, bsq = lib.query_pack(case['qk'].to(dev).contiguous(), case['qe'].to(dev).contiguous(), hw_pad)
    wsb = L.xm_affinity_workspace_bytes(hw, case['n_obj'])
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out_chw = torch.zeros(case['n_obj'], CV, hw, dtype=torch.float16, device=dev)
    out_hwc = torch.zeros(case['n_obj'], hw, CV, dtype=torch.float16, device=dev)
    a.qp, a.bsq, a.hw, a.hw_pad, a.top_k, a.n_obj_total = qp.data_ptr(), bsq.data_ptr(), hw, hw_pad, top_k, case['n_obj']
    a.readout_chw, a.readout_hwc = out_chw.data_ptr(), out_hwc.data_ptr()
    a.workspace, a.workspace_bytes = ws.data_ptr(), wsb
    dbg = None
    if want_debug:
        ob, no, begins = case['groups'][0]
        n0 = sum(b['n'] - begins[bi] for bi, b in enumerate(case['banks']) if b is not None and begins[bi] < b['n'])
        dbg = torch.full((n0, hw_pad), float('nan'), dtype=torch.float32, device=dev)
        a.debug_scores = dbg.data_ptr()
    lib.check(L.xm_affinity_readout(C.byref(a), lib.stream_ptr()), 'xm_affinity_readout')
    try:
        torch.cuda.synchronize()
    except Exception as e:
        raise RuntimeError(f'kernel fault: {e}; last trap: {lib.last_trap()}')
    return out_chw, out_hwc, usage_bufs, dbg


def compare(case, top_k=30, verbose=False):
    exp_out, exp_usage, amb, exp_scores = expected(case, top_k)
    out_chw, out_hwc, usage_bufs, dbg = run_kernel(case, top_k, want_debug=True)
    res = {}
    hw = case['hw']
    res['score_err'] = (dbg[:, :hw].double().cpu() - exp_scores).abs().max().item()
    got = out_chw.double().cpu()
    err = (got - exp_out).abs()                       # [n_obj, CV, hw]
    per_q = err.amax(dim=(0, 1))
    res['ambiguous'] = int(amb.sum())
    res['out_err_clear'] = per_q[~amb].max().item() if (~amb).any() else 0.0
    res['out_err_all'] = per_q.max().item()
    res['layout_err'] = (out_hwc.transpose(1, 2).float() - out_chw.float()).abs().max().item()
    ob, no, begins = case['groups'][0]
    u = torch.cat([usage_bufs[bi][begins[bi]:b['n']].cpu().double() for bi, b in enumerate(case['banks'])
                   if b is not None and begins[bi] < b['n']])
    res['usage_err'] = (u - exp_usage).abs().max().item()
    res['usage_sum'] = u.sum().item()
    if verbose:
        print(res)
    return res
