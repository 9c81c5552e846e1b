"""Diagnostic: where do argmax disagreements of the 2-object portrait clip come from?  (run on a GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity_clip as pc
from xmem2_b200.model.network import XMem
from xmem2_b200.util.synth import synth_state_dict

torch.set_grad_enabled(False)
CFG = dict(mem_every=10, deep_update_every=-1, enable_long_term=True, enable_long_term_count_usage=False, hidden_dim=64,
           key_dim=64, value_dim=512, top_k=30, max_mid_term_frames=10, min_mid_term_frames=5, num_prototypes=128,
           max_long_term_elements=10000)
state = synth_state_dict(0)
net = XMem(dict(CFG), None).to('cuda').eval(); net.load_weights(dict(state))
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (853, 480)
n_frames = int(sys.argv[3]) if len(sys.argv) > 3 else 30

orig_add = pc.Stats.add
def add(self, p, po):
    orig_add(self, p, po)
    am, amo = p.argmax(0), po.argmax(0)
    conf = torch.zeros(3, 3, dtype=torch.long)
    for a in range(3):
        for b in range(3):
            conf[a, b] = int(((amo == a) & (am == b)).sum())
    dp = (p.float() - po.float()).abs()
    print(f'frame {self.frames - 1:3d}: mismatch {float((am != amo).float().mean()):.4f}  confusion(oracle->ours) {conf.tolist()}  '
          f'mean|dp| per class {[round(float(dp[k].mean()), 5) for k in range(3)]}  max|dp| {float(dp.max()):.4f}  '
          f'oracle mean p {[round(float(po[k].mean()), 4) for k in range(3)]}')
pc.Stats.add = add
pc.run_lockstep(net, state, H, W, n_frames, 2, [0, 20, 40, 60, 80], [0, 20], CFG, structured=True)
