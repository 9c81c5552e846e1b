"""Profiling helper: the fused memory read (xm_affinity_readout) alone at the BASELINE config-2 maximum
(HW=1620, 9 working + 5 permanent frames = 22 680 columns, 1 object)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import k1_ref
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 3
case = k1_ref.make_case(hw=1620, sizes=(0, 9 * 1620, 5 * 1620), n_obj=1, group_begins=[(0, 1, [0, 0, 0])], seed=11)
for _ in range(n_iter):
    k1_ref.run_kernel(case)
torch.cuda.synchronize()
print('done')
