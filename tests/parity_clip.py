"""Lock-step end-to-end parity harness: the SAME synthetic clip through this package's InferenceCore (CUDA kernels) and
through the oracle (oracle/xmem_oracle.py, the pinned restatement of the reference) in fp32 on the same GPU, optionally
also through the oracle under fp16 autocast (= how the reference itself runs on a GPU, inference/run_on_video.py:76).

Compared per frame (reference inference/inference_core.py:62-152 outputs):
  * memory-bank sizes (bit-exact),
  * pre-argmax logits: the log-odds of every object against the background, log p_k - log p_0 (the quantity the
    argmax of `_post_process` (run_on_video.py:165-173) decides on), on pixels where the oracle is unsaturated
    (|log-odds| < 8, i.e. away from the 1e-7 clamp of aggregate(), model/aggregate.py:10): p50 / p99 / max of |delta|,
  * argmax label maps: number of disagreeing pixels and, among them, the LARGEST fp32-oracle top-2 log-probability
    margin (= the epsilon above which the argmax is identical on every pixel).
Test infrastructure only (imports oracle/)."""
import contextlib

import torch

from oracle import xmem_oracle as O
from xmem2_b200.inference.inference_core import InferenceCore
from xmem2_b200.util.synth import synth_frame, synth_mask


def _log_odds(p):
    lp = p.float().clamp_min(1e-30).log()
    return lp[1:] - lp[0:1]


class Stats:
    def __init__(self):
        self.p50 = self.p99 = self.max = 0.0
        self.mismatch = 0
        self.pixels = 0
        self.eps = 0.0
        self.mean_dprob = 0.0
        self.frames = 0
        self.per_frame = []          # (frame, argmax mismatch fraction, largest oracle margin among the mismatches)
        self.mismatch_unsat = 0      # disagreements on pixels whose oracle top-2 classes are both away from the 1e-7 clamp
        self.pixels_unsat = 0
        self.eps_unsat = 0.0

    def add(self, p, po):
        d, do = _log_odds(p), _log_odds(po)
        unsat = do.abs() < 8
        if unsat.any():
            e = (d - do).abs()[unsat]
            if e.numel() > 4_000_000:
                e = e[:: e.numel() // 4_000_000 + 1]
            q = torch.quantile(e, torch.tensor([0.5, 0.99], device=e.device))
            self.p50 = max(self.p50, q[0].item()); self.p99 = max(self.p99, q[1].item()); self.max = max(self.max, e.max().item())
        am, amo = p.argmax(0), po.argmax(0)
        bad = am != amo
        self.mismatch += int(bad.sum()); self.pixels += bad.numel()
        eps_f = 0.0
        lpo = po.float().clamp_min(1e-30).log()
        top2 = lpo.topk(2, dim=0).values
        # "saturated tie": the background probability has vanished (some object's sigmoid sits at the 1e-7 clamp of aggregate(),
        # aggregate.py:10) but a second class still holds > 1e-3: two objects whose sigmoids are BOTH within a few fp32 ulps
        # of 1.  Their odds are small integer multiples of 6e-8, so the "margins" ln 2 / ln 3 there are float artefacts.
        unsat_px = ~((po[0].float() < 1e-6) & (top2[1] > -6.9))
        self.pixels_unsat += int(unsat_px.sum()); self.mismatch_unsat += int((bad & unsat_px).sum())
        if (bad & unsat_px).any():
            self.eps_unsat = max(self.eps_unsat, (top2[0] - top2[1])[bad & unsat_px].max().item())
        if bad.any():
            eps_f = (top2[0] - top2[1])[bad].max().item()
            self.eps = max(self.eps, eps_f)
        self.per_frame.append((self.frames, round(float(bad.float().mean()), 6), round(eps_f, 4)))
        self.mean_dprob = max(self.mean_dprob, (p.float() - po.float()).abs().mean().item())
        self.frames += 1

    def as_dict(self):
        return dict(dlogit_p50=self.p50, dlogit_p99=self.p99, dlogit_max=self.max, argmax_mismatch_frac=self.mismatch / max(1, self.pixels),
                    argmax_eps=self.eps, mean_dprob=self.mean_dprob, frames=self.frames,
                    argmax_mismatch_unsat_frac=self.mismatch_unsat / max(1, self.pixels_unsat), argmax_eps_unsat=self.eps_unsat,
                    unsat_pixel_frac=self.pixels_unsat / max(1, self.pixels),
                    worst_frames=sorted(self.per_frame, key=lambda t: -t[1])[:8])


def run_lockstep(net, state, H, W, n_frames, n_obj, annotated, first_frame_of, cfg, structured, seed=1234, with_autocast=False,
                 dev='cuda'):
    """Returns (stats_ours_vs_fp32, stats_autocast_vs_fp32 or None, cores).  Raises AssertionError on a bank-size mismatch."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    state_d = {k: v.to(dev) for k, v in state.items()}
    onet = O.OracleNet(state_d)
    core = InferenceCore(net, dict(cfg))
    ocore = O.OracleCore(onet, dict(cfg))
    acore = O.OracleCore(onet, dict(cfg)) if with_autocast else None
    auto = (lambda: torch.autocast('cuda', dtype=torch.float16)) if with_autocast else None
    frame = lambda ti: synth_frame(ti, H, W, seed=seed, structured=structured).to(dev)
    mask = lambda ti: synth_mask(ti, H, W, n_obj, first_frame_of).to(dev)
    n_seen = 0
    for j in list(set(annotated)):                       # CPython set order, run_on_video.py:45,65-66
        n_seen = max(n_seen, n_obj if first_frame_of is None else sum(1 for f in first_frame_of if f <= j))
        labels = list(range(1, n_seen + 1))
        for c in (core, ocore, acore):
            if c is None:
                continue
            c.set_all_labels(labels)
            with (auto() if c is acore else contextlib.nullcontext()):
                c.put_to_permanent_memory(frame(j), mask(j)[:n_seen].clone())
    labels = list(range(1, n_seen + 1))
    s_ours, s_auto = Stats(), (Stats() if with_autocast else None)
    for ti in range(n_frames):
        img = frame(ti)
        m = mask(ti)[:n_seen] if ti in annotated else None
        kw = dict(end=(ti == n_frames - 1), do_not_add_mask_to_memory=m is not None)
        p = core.step(img, m.clone() if m is not None else None, labels if m is not None else None, **kw)
        po = ocore.step(img, m.clone() if m is not None else None, labels if m is not None else None, **kw)
        sizes = [core.memory.temporary_work_mem.size, core.memory.permanent_work_mem.size, core.memory.long_mem.size]
        osizes = [ocore.mem.temp.size, ocore.mem.perm.size, ocore.mem.long.size]
        assert sizes == osizes, (ti, sizes, osizes)
        assert p.shape == po.shape and torch.isfinite(p).all()
        s_ours.add(p, po)
        if acore is not None:
            with auto():
                pa = acore.step(img, m.clone() if m is not None else None, labels if m is not None else None, **kw)
            s_auto.add(pa, po)
    return s_ours, s_auto, (core, ocore)
