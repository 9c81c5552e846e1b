"""Test helpers for the fused memory-read kernel (K1): build random banks in the kernel's arena layout,
run the C ABI, and compute the expected result with the oracle's attention math on the same
fp16-rounded operands (oracle/xmem_oracle.py: similarity + softmax_topk)."""
import ctypes as C
import torch

from xmem2_b200 import lib

CK, CV = 64, 512


from xmem2_b200.util.synth_memory import make_case      # noqa: E402,F401  (shared with bench.py)


def case_from_attention_golden(device='cuda'):
    """tests/golden/attention.npz (inputs + outputs of the LIVE reference's get_similarity / do_softmax / readout,
    tests/golden/make_golden.py:68-88) as a kernel case: operands rounded to fp16, the fixture's 32 value channels
    embedded in the kernel's 512 (the rest zero), its two value planes as two objects of one group."""
    import os
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'attention.npz'))
    n, hw = d['mk'].shape[2], d['qk'].shape[2]
    cap = (n + 64 + 7) // 8 * 8
    val = torch.zeros(2, CV, cap).half()
    val[:, :d['v'].shape[1], :n] = torch.from_numpy(d['v']).half()
    bank = dict(key=torch.from_numpy(d['mk'][0]).t().contiguous().half(), shr=torch.from_numpy(d['ms'][0, 0]).float(), val=val, cap=cap, n=n)
    return dict(hw=hw, banks=[None, None, bank], n_obj=2, groups=[(0, 2, [0, 0, 0])],
                qk=torch.from_numpy(d['qk'][0]).t().contiguous().half(), qe=torch.from_numpy(d['qe'][0]).t().contiguous().half(),
                device=device)


def kernel_operand_similarity(case):
    """fp64 evaluation of get_similarity (memory_util.py:7-39) on the operands as the kernel (and the reference under
    autocast) rounds them: k^2 and 2*k*e are fp16 values, b_sq is fp32.  One [n, hw] matrix per bank (None = empty)."""
    qe = case['qe'].double()
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
        sims.append((sp - bsq[None, :]) * b['shr'].double()[:, None] / 8.0)
    return sims


def oracle_similarity(case):
    """oracle.similarity (the pinned restatement of memory_util.py:7-39) in fp64 on the same fp16-valued keys: differs
    from `kernel_operand_similarity` only by the fp16 rounding of k^2 and 2*k*e (tests/test_k1_ref_vs_oracle.py)."""
    from oracle import xmem_oracle as O
    qk = case['qk'].double().t().unsqueeze(0); qe = case['qe'].double().t().unsqueeze(0)
    out = []
    for b in case['banks']:
        if b is None:
            out.append(None); continue
        out.append(O.similarity(b['key'].double().t().unsqueeze(0), b['shr'].double().view(1, 1, -1), qk, qe)[0])
    return out


def expected(case, top_k=30):
    """Expected readout / usage: similarity on the kernel's fp16 operands (above), then the ORACLE's do_softmax
    (oracle.softmax_topk = memory_util.py:41-65) and the reference's `v @ affinity` with the affinity cast to fp16
    (memory_manager.py:57-59 under autocast)."""
    from oracle import xmem_oracle as O
    sims = kernel_operand_similarity(case)
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
        if s.shape[0] > top_k:
            tv = torch.topk(s, top_k + 1, dim=0).values
            ambiguous |= (tv[top_k - 1] - tv[top_k]) < 3e-5
        aff, usage = O.softmax_topk(s.unsqueeze(0), top_k, want_usage=True)
        aff = aff[0]
        outs.append(torch.einsum('ocn,nq->ocq', v, aff.half().double()))
        if gi == 0:
            usages = usage[0]
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
    ws = lib.affinity_workspace(hw, case['n_obj'], dev)
    wsb = ws.numel()
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
