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
