"""Diagnostic: stage-by-stage error of the CUDA network vs the oracle / golden fixtures (prints, no asserts)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xmem_oracle as O
from xmem2_b200.model.network import XMem
from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.util.synth import synth_state_dict, synth_frame, synth_mask

torch.set_grad_enabled(False)
G = os.path.join(os.path.dirname(__file__), 'golden')
dev = 'cuda'


def rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return f'max|d|={(a - b).abs().max().item():.3e} mean|d|={(a - b).abs().mean().item():.3e} ref_max={b.abs().max().item():.3e}'


def main():
    state = synth_state_dict(0)
    cfg = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=True, hidden_dim=64,
               key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
               max_long_term_elements=10000)
    net = XMem(dict(cfg), None).to(dev).eval()
    net.load_weights(dict(state))
    on = O.OracleNet(state)
    H, W = 64, 96
    img = synth_frame(0, H, W, structured=True)[None]
    masks = synth_mask(0, H, W, 2)[None]
    ok, os_, oe, of16, of8, of4 = on.encode_key(img)
    key, shr, sel, f16, f8, f4 = net.encode_key(img.to(dev))
    torch.cuda.synchronize()
    for n, a, b in (('key', key, ok), ('shr', shr, os_), ('sel', sel, oe), ('f16', f16, of16), ('f8', f8, of8), ('f4', f4, of4)):
        print(f'{n:8s}', tuple(a.shape), rel(a, b))
    d = np.load(os.path.join(G, 'network.npz'))
    hid = torch.from_numpy(d['hid']); ro = torch.from_numpy(d['ro'])
    ov, oh2 = on.encode_value(img, of16, hid, masks, True)
    hid_dev = hid.to(dev).permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    v, h2 = net.encode_value(img.to(dev), f16, hid_dev, masks.to(dev), True)
    torch.cuda.synchronize()
    print('value   ', tuple(v.shape), rel(v, ov)); print('hid2    ', tuple(h2.shape), rel(h2, oh2))
    onh, ologits, oprob = on.segment((of16, of8, of4), ro, hid, True, False)
    ro_dev = ro.to(dev).half().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    nh, logits, prob = net.segment((f16, f8, f4), ro_dev, hid_dev, h_out=True, strip_bg=False)
    torch.cuda.synchronize()
    print('new_hid ', tuple(nh.shape), rel(nh, onh)); print('logits  ', tuple(logits.shape), rel(logits, ologits))
    print('prob    ', tuple(prob.shape), rel(prob, oprob))
    print('argmax agreement', (prob.argmax(1).cpu() == oprob.argmax(1)).float().mean().item())
    e = (logits.float().cpu() - ologits).abs()
    idx = torch.nonzero(e == e.max())[0].tolist()
    print('worst logit at', idx, 'mine', logits[tuple(idx)].item(), 'oracle', ologits[tuple(idx)].item())
    mid = ologits.abs() < 8
    print('logit err where |logit|<8: max', e[mid].max().item(), 'mean', e[mid].mean().item(), 'count', int(mid.sum()))
    # calibration: the oracle itself run the way the reference runs on a GPU (fp16 autocast, run_on_video.py:76)
    og = O.OracleNet({k: v.to(dev) for k, v in state.items()})
    with torch.autocast('cuda', dtype=torch.float16):
        gk, gs, ge, gf16, gf8, gf4 = og.encode_key(img.to(dev))
        gv, gh2 = og.encode_value(img.to(dev), gf16, hid.to(dev), masks.to(dev), True)
        gnh, glogits, gprob = og.segment((gf16, gf8, gf4), ro.to(dev), hid.to(dev), True, False)
    print('--- oracle under cuda fp16 autocast vs oracle fp32 (what the reference GPU path itself deviates by)')
    for n, a, b in (('key', gk, ok), ('f16', gf16, of16), ('value', gv, ov), ('hid2', gh2, oh2), ('new_hid', gnh, onh),
                    ('logits', glogits, ologits), ('prob', gprob, oprob)):
        print(f'{n:8s}', rel(a, b))
    eg = (glogits.float().cpu() - ologits).abs()
    print('autocast logit err where |logit|<8: max', eg[mid].max().item(), 'mean', eg[mid].mean().item())
    print('autocast argmax agreement', (gprob.argmax(1).cpu() == oprob.argmax(1)).float().mean().item())
    print('mine vs autocast: logits', rel(logits, glogits), ' prob', rel(prob, gprob))

    for name in ('one_obj', 'two_obj'):
        d = np.load(os.path.join(G, f'clip_{name}.npz'))
        Hc, Wc, n_frames, n_obj, save_every = [int(x) for x in d['hw']]
        c = dict(cfg); c.update({str(k): int(v) for k, v in zip(d['cfg_keys'], d['cfg_vals'])})
        ffo = [int(x) for x in d['first_frame_of']]; annotated = [int(x) for x in d['annotated']]
        core = InferenceCore(net, dict(c))
        n_seen = 0
        for j in [int(x) for x in d['order']]:
            n_seen = max(n_seen, sum(1 for f in ffo if f <= j))
            core.set_all_labels(list(range(1, n_seen + 1)))
            core.put_to_permanent_memory(synth_frame(j, Hc, Wc, structured=True).to(dev), synth_mask(j, Hc, Wc, n_obj, ffo)[:n_seen].to(dev))
        labels = list(range(1, n_seen + 1))
        k = 0; worst = 0.0; worst_mean = 0.0; agree = 1.0
        t0 = time.time()
        for ti in range(n_frames):
            msk = synth_mask(ti, Hc, Wc, n_obj, ffo).to(dev) if ti in annotated else None
            p = core.step(synth_frame(ti, Hc, Wc, structured=True).to(dev), msk, labels if msk is not None else None,
                          end=(ti == n_frames - 1), do_not_add_mask_to_memory=msk is not None)
            m = core.memory
            sizes = [m.temporary_work_mem.size, m.permanent_work_mem.size, m.long_mem.size]
            if sizes != d['sizes'][ti][:3].tolist():
                print(f'  SIZE MISMATCH ti={ti} got {sizes} want {d["sizes"][ti][:3].tolist()}')
            if ti % save_every == 0:
                ref = torch.from_numpy(d['probs'][k]).float(); k += 1
                e = (p.float().cpu() - ref).abs()
                worst = max(worst, e.max().item()); worst_mean = max(worst_mean, e.mean().item())
                agree = min(agree, (p.argmax(0).cpu() == ref.argmax(0)).float().mean().item())
        torch.cuda.synchronize()
        print(f'clip {name}: max|dprob|={worst:.3e} worst frame mean={worst_mean:.3e} min argmax agreement={agree:.5f} '
              f'groups={core.memory.permanent_work_mem.num_groups} time={time.time() - t0:.2f}s')
        fh = torch.from_numpy(d['final_hidden']).float()
        print('   final hidden', rel(core.memory.get_hidden(), fh))
        if len(d['temp_usage']):
            print('   temp usage  ', rel(core.memory.temporary_work_mem.get_usage(), torch.from_numpy(d['temp_usage'])))


if __name__ == '__main__':
    main()
