"""Back-of-the-envelope model of the convolution kernel (no GPU needed): replays the host tile plan of
csrc/conv_igemm.cu for every convolution shape of a 480p frame and predicts its time from ONE measured constant, the
bytes per second an SM pulls through its L2 port (NOTES.md: ~70 GB/s per active SM, from the ncu launch list),
plus a fixed launch/prologue/epilogue cost.  Then predicts what the experimental variants would buy:
  csk  : cluster split-K (2-3 CTAs per output tile, DSMEM fix-up ~1.5 us) for grids that leave SMs idle
  2cta : CTA pairs, M=256, each CTA loads half of the weight tile (BN=128 or 256)
  halo : 3x3 stride-1 layers load one haloed 36 KB input tile per 64-channel block instead of nine 16 KB boxes
Usage: python tests/conv_model.py [GBps_per_SM] [fixed_us]"""
import sys

SMS = 148
R = float(sys.argv[1]) if len(sys.argv) > 1 else 72.0      # GB/s per SM
T0 = float(sys.argv[2]) if len(sys.argv) > 2 else 3.5      # us
MMA_FLOP_PER_US = 2.25e15 / SMS / 1e6                       # dense fp16 per SM

h, w = 30, 54
SHAPES = [  # (name, H, W, cin, cout, k, stride, count per frame) — xmem2_b200/util/conv_bench.py
    ('stem 1x1 K192', 240, 432, 192, 64, 1, 1, 1), ('res2 1x1 64->64', 120, 216, 64, 64, 1, 1, 1),
    ('res2 3x3 64', 120, 216, 64, 64, 3, 1, 3), ('res2 1x1 64->256', 120, 216, 64, 256, 1, 1, 4),
    ('res2 1x1 256->64', 120, 216, 256, 64, 1, 1, 2), ('l2 1x1 256->128', 120, 216, 256, 128, 1, 1, 1),
    ('l2 3x3 128 s2', 120, 216, 128, 128, 3, 2, 1), ('l2 ds 256->512 s2', 120, 216, 256, 512, 1, 2, 1),
    ('l2 1x1 128->512', 60, 108, 128, 512, 1, 1, 4), ('l2 1x1 512->128', 60, 108, 512, 128, 1, 1, 3),
    ('l2 3x3 128', 60, 108, 128, 128, 3, 1, 3), ('l3 1x1 512->256', 60, 108, 512, 256, 1, 1, 1),
    ('l3 3x3 256 s2', 60, 108, 256, 256, 3, 2, 1), ('l3 ds 512->1024 s2', 60, 108, 512, 1024, 1, 2, 1),
    ('l3 1x1 256->1024', h, w, 256, 1024, 1, 1, 6), ('l3 1x1 1024->256', h, w, 1024, 256, 1, 1, 5),
    ('l3 3x3 256', h, w, 256, 256, 3, 1, 5), ('keyproj 3x3 1024->129', h, w, 1024, 129, 3, 1, 1),
    ('fuser 3x3 1600->512', h, w, 1600, 512, 3, 1, 2), ('fuser 3x3 512->512', h, w, 512, 512, 3, 1, 3),
    ('up16 skip 3x3 512', 60, 108, 512, 512, 3, 1, 1), ('up16 3x3 512->256', 60, 108, 512, 256, 3, 1, 2),
    ('up16 3x3 256->256', 60, 108, 256, 256, 3, 1, 1), ('up8 3x3 256 @1/4', 120, 216, 256, 256, 3, 1, 3),
    ('pred 3x3 256->1', 120, 216, 256, 1, 3, 1, 1), ('hu 1x1 512->256', h, w, 512, 256, 1, 1, 1),
    ('hu 1x1 256->256', h, w, 256, 256, 1, 1, 1), ('hu 1x1 320->256', h, w, 320, 256, 1, 1, 1),
    ('hu 3x3 320->192', h, w, 320, 192, 3, 1, 1),
]


def tiles_of(Ho, Wo):
    best = None
    for tw in (8, 16, 32):
        th = 128 // tw
        area = -(-Wo // tw) * -(-Ho // th)
        if best is None or area < best:
            best = area
    return best


def plan(Ho, Wo, cin, cout, k):
    """the production heuristic (conv_igemm.cu: BN, split-K, pipeline depth)"""
    cout_pad = -(-cout // 64) * 64
    tiles = tiles_of(Ho, Wo)
    ksteps = k * k * (cin // 64)
    bn = 128 if cout_pad % 128 == 0 else 64
    if bn == 128 and ksteps <= 128 and tiles * (cout_pad // 128) * 2 <= SMS:
        bn = 64
    ctas = tiles * (cout_pad // bn)
    splits = 1
    if (ctas * 3 <= SMS and ksteps >= 48) or (ctas * 2 <= SMS and ksteps > 128):
        splits = max(1, min(SMS // ctas, ksteps // 16, 4))
    kps = -(-ksteps // splits)
    splits = -(-ksteps // kps)
    depth = 2 if kps <= 4 else (6 if ctas * splits < 2 * SMS else 3)
    return dict(bn=bn, ctas=ctas, splits=splits, kps=kps, depth=depth, ksteps=ksteps, cout_pad=cout_pad, tiles=tiles)


def time_us(ctas_total, bytes_per_cta, flop_per_cta, fixup_us=0.0):
    on_busiest_sm = -(-ctas_total // SMS)
    ingest = on_busiest_sm * bytes_per_cta / (R * 1e3)            # us  (GB/s = 1e3 bytes/us)
    mma = on_busiest_sm * flop_per_cta / MMA_FLOP_PER_US
    return T0 + max(ingest, mma) + fixup_us


def main():
    tot = dict(now=0.0, csk=0.0, cta2=0.0, both=0.0, halo=0.0, all=0.0)
    print(f'{"layer":24s} {"grid":>14s} {"now":>7s} {"csk":>7s} {"2cta":>7s} {"halo":>7s}   (us per launch; R={R} GB/s/SM, T0={T0} us)')
    for name, H, W, cin, cout, k, s, cnt in SHAPES:
        Ho, Wo = H // s, W // s
        p = plan(Ho, Wo, cin, cout, k)
        kb = 16384 + p['bn'] * 128
        flop_step = 2.0 * 128 * p['bn'] * 64
        now = time_us(p['ctas'] * p['splits'], p['kps'] * kb, p['kps'] * flop_step, fixup_us=10.0 if p['splits'] > 1 else 0.0)
        # cluster split-K: as many splits (<= 3 for BN=64, 2 for BN=128) as fit in one wave, >= 8 k-steps each
        bn_c = p['bn']
        ctas_c = p['tiles'] * (p['cout_pad'] // bn_c)
        sc = 1
        while sc < (3 if bn_c == 64 else 2) and ctas_c * (sc + 1) <= SMS and p['ksteps'] // (sc + 1) >= 8:
            sc += 1
        kps_c = -(-p['ksteps'] // sc)
        csk = time_us(ctas_c * sc, kps_c * (16384 + bn_c * 128), kps_c * 2.0 * 128 * bn_c * 64, fixup_us=1.5 if sc > 1 else 0.0)
        csk = min(csk, now) if sc == 1 else csk
        # CTA pairs: BN = 256 if possible else 128 (needs cout_pad % 128 == 0); each CTA ingests A + half of B
        if p['cout_pad'] % 128 == 0:
            bn2 = 256 if p['cout_pad'] % 256 == 0 else 128
            ctas2 = (p['tiles'] + 1) // 2 * 2 * (p['cout_pad'] // bn2)
            cta2 = time_us(ctas2, p['ksteps'] * (16384 + bn2 * 64), p['ksteps'] * 2.0 * 128 * bn2 * 64)
        else:
            cta2 = now
        # haloed tile: 8x16 output tiles, per 64-channel block 36 KB of input + 9 weight tiles
        if k == 3 and s == 1:
            tiles_h = -(-Wo // 8) * -(-Ho // 16)
            bn_h = 128 if (p['cout_pad'] % 128 == 0 and tiles_h * (p['cout_pad'] // 128) * 2 > SMS) else 64
            ctas_h = tiles_h * (p['cout_pad'] // bn_h)
            blocks = cin // 64
            halo = time_us(ctas_h, blocks * (36864 + 9 * bn_h * 128), blocks * 9 * 2.0 * 128 * bn_h * 64)
        else:
            halo = now
        best = min(now, csk, cta2)
        tot['halo'] += min(now, halo) * cnt; tot['all'] += min(best, halo) * cnt
        tot['now'] += now * cnt; tot['csk'] += min(now, csk) * cnt; tot['cta2'] += min(now, cta2) * cnt; tot['both'] += best * cnt
        grid = f"({p['tiles']},{p['cout_pad'] // p['bn']},{p['splits']})x{p['bn']}"
        print(f'{name:24s} {grid:>14s} {now:7.1f} {csk:7.1f} {cta2:7.1f} {halo:7.1f}   x{cnt}')
    print(f"frame sum: now {tot['now']:.0f} us | with csk {tot['csk']:.0f} | with 2cta {tot['cta2']:.0f} | best of both {tot['both']:.0f} | with halo {tot['halo']:.0f} | best of all {tot['all']:.0f}"
          f"   (measured in round 1: 909 us graph-timed)")


if __name__ == '__main__':
    main()
