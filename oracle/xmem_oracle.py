"""CPU oracle for the XMem++ per-frame memory-attention path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (fp32, CPU) restatement of the reference algorithm
for the hot path named in BASELINE.json: encode_key -> affinity -> top-k softmax
-> readout -> decode, plus the memory bookkeeping around it.  It exists so the
CUDA implementation in `xmem2_b200/` can be checked on a box where
`/root/reference` does not exist.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product
package never does.

Pinning: the reference ships NO tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, generated in the
build container by `tests/golden/make_golden.py` (imports `/root/reference`,
loads the same hash-seeded parameters) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` re-checks the oracle against those fixtures.

Every function cites the reference lines it restates (paths relative to
/root/reference).  The arithmetic lives in PyTorch ATen in the reference as well
(SURVEY.md 8c), so this oracle uses the same primitive ops (conv2d, matmul,
topk, interpolate) but none of the reference's module structure.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# attention math  (model/memory_util.py)
# ----------------------------------------------------------------------------------------------
def similarity(mk: Tensor, ms: Optional[Tensor], qk: Tensor, qe: Optional[Tensor]) -> Tensor:
    """Anisotropic-L2 similarity, memory_util.py:7-39.
    mk [B,CK,N], ms [B,1,N] or None, qk [B,CK,Q], qe [B,CK,Q] or None -> [B,N,Q]."""
    ck = mk.shape[1]
    mk = mk.flatten(2); qk = qk.flatten(2)
    mkt = mk.transpose(1, 2)
    if qe is not None:
        qe = qe.flatten(2)
        a_sq = mkt.pow(2) @ qe                      # :24
        two_ab = 2 * (mkt @ (qk * qe))              # :25
        b_sq = (qe * qk.pow(2)).sum(1, keepdim=True)  # :26
        s = -a_sq + two_ab - b_sq                   # :27
    else:
        s = -mk.pow(2).sum(1).unsqueeze(2) + 2 * (mkt @ qk)   # :30-32
    if ms is not None:
        s = s * ms.flatten(1).unsqueeze(2) / math.sqrt(ck)    # :35
    else:
        s = s / math.sqrt(ck)
    return s


def softmax_topk(sim: Tensor, top_k: Optional[int], want_usage: bool = False):
    """memory_util.py:41-65.  top-k branch has NO max subtraction (:48-49)."""
    if top_k is not None:
        vals, idx = torch.topk(sim, k=top_k, dim=1)
        e = vals.exp()
        e = e / e.sum(dim=1, keepdim=True)
        aff = torch.zeros_like(sim).scatter_(1, idx, e)
    else:
        m = sim.max(dim=1, keepdim=True)[0]
        e = (sim - m).exp()
        aff = e / e.sum(dim=1, keepdim=True)
    if want_usage:
        return aff, aff.sum(dim=2)                  # :62-63
    return aff


def aggregate(prob: Tensor, dim: int, return_logits: bool = False):
    """Soft aggregation, model/aggregate.py:6-16."""
    bg = torch.prod(1 - prob, dim=dim, keepdim=True)
    p = torch.cat([bg, prob], dim).clamp(1e-7, 1 - 1e-7)
    logits = torch.log(p / (1 - p))
    out = F.softmax(logits, dim=dim)
    return (logits, out) if return_logits else out


def pad_to_16(x: Tensor):
    """util/tensor_util.py:47-62 with d=16."""
    h, w = x.shape[-2:]
    nh = h if h % 16 == 0 else h + 16 - h % 16
    nw = w if w % 16 == 0 else w + 16 - w % 16
    lh, lw = (nh - h) // 2, (nw - w) // 2
    pads = (lw, nw - w - lw, lh, nh - h - lh)
    return F.pad(x, pads), pads


def unpad(x: Tensor, pads):
    """util/tensor_util.py:64-77."""
    lw, uw, lh, uh = pads
    if lh + uh > 0:
        x = x[..., lh:x.shape[-2] - uh, :]
    if lw + uw > 0:
        x = x[..., lw:x.shape[-1] - uw]
    return x


# ----------------------------------------------------------------------------------------------
# network (model/network.py, model/modules.py, model/resnet.py, model/cbam.py, group_modules.py)
# ----------------------------------------------------------------------------------------------
class OracleNet:
    """Functional forward passes over a flat upstream-format state dict."""

    def __init__(self, state: Dict[str, Tensor], hidden_dim: int = 64, eps: float = 1e-5):
        self.w = {k: v.float() if v.is_floating_point() else v for k, v in state.items()}
        self.hidden_dim = hidden_dim
        self.eps = eps

    # -- primitives ---------------------------------------------------------------------------
    def conv(self, x, name, stride=1, pad=None):
        w = self.w[name + '.weight']
        b = self.w.get(name + '.bias')
        if pad is None:
            pad = w.shape[-1] // 2
        return F.conv2d(x, w, b, stride=stride, padding=pad)

    def bn(self, x, name):
        return F.batch_norm(x, self.w[name + '.running_mean'], self.w[name + '.running_var'],
                            self.w[name + '.weight'], self.w[name + '.bias'], False, 0.0, self.eps)

    def bottleneck(self, x, p, stride):          # resnet.py:77-114
        o = F.relu(self.bn(self.conv(x, p + '.conv1'), p + '.bn1'))
        o = F.relu(self.bn(self.conv(o, p + '.conv2', stride=stride), p + '.bn2'))
        o = self.bn(self.conv(o, p + '.conv3'), p + '.bn3')
        if (p + '.downsample.0.weight') in self.w:
            x = self.bn(self.conv(x, p + '.downsample.0', stride=stride), p + '.downsample.1')
        return F.relu(o + x)

    def basic(self, x, p, stride):               # resnet.py:46-74
        o = F.relu(self.bn(self.conv(x, p + '.conv1', stride=stride), p + '.bn1'))
        o = self.bn(self.conv(o, p + '.conv2'), p + '.bn2')
        if (p + '.downsample.0.weight') in self.w:
            x = self.bn(self.conv(x, p + '.downsample.0', stride=stride), p + '.downsample.1')
        return F.relu(o + x)

    def group_res(self, g, p):                   # group_modules.py:36-54, g is [B*n, C, H, W]
        o = self.conv(F.relu(g), p + '.conv1')
        o = self.conv(F.relu(o), p + '.conv2')
        if (p + '.downsample.weight') in self.w:
            g = self.conv(g, p + '.downsample')
        return o + g

    def cbam(self, x, p):                        # cbam.py:23-77
        def mlp(v):
            v = F.relu(F.linear(v, self.w[p + '.ChannelGate.mlp.1.weight'], self.w[p + '.ChannelGate.mlp.1.bias']))
            return F.linear(v, self.w[p + '.ChannelGate.mlp.3.weight'], self.w[p + '.ChannelGate.mlp.3.bias'])
        att = mlp(x.mean(dim=(2, 3))) + mlp(x.amax(dim=(2, 3)))
        x = x * torch.sigmoid(att)[:, :, None, None]
        comp = torch.cat([x.amax(dim=1, keepdim=True), x.mean(dim=1, keepdim=True)], 1)
        return x * torch.sigmoid(self.conv(comp, p + '.SpatialGate.spatial.conv'))

    def fusion(self, x, g, p):                   # modules.py:22-41; x [B,Cx,H,W], g [B,n,Cg,H,W]
        b, n = g.shape[:2]
        xg = torch.cat([x.unsqueeze(1).expand(-1, n, -1, -1, -1), g], 2).flatten(0, 1)
        g1 = self.group_res(xg, p + '.block1')
        r = self.cbam(g1, p + '.attention')
        g2 = self.group_res(g1 + r, p + '.block2')
        return g2.view(b, n, *g2.shape[1:])

    def gru(self, g, h, p):                      # modules.py:63-74, 88-99 (not a textbook GRU)
        b, n = g.shape[:2]
        hd = self.hidden_dim
        v = self.conv(torch.cat([g, h], 2).flatten(0, 1), p + '.transform').view(b, n, 3 * hd, *g.shape[-2:])
        f = torch.sigmoid(v[:, :, :hd]); u = torch.sigmoid(v[:, :, hd:2 * hd]); nv = torch.tanh(v[:, :, 2 * hd:])
        return f * h * (1 - u) + u * nv

    # -- public passes ------------------------------------------------------------------------
    def encode_key(self, frame, need_sk=True, need_ek=True):
        """network.py:40-70, modules.py:166-175 and 207-211.  frame [B,3,H,W]."""
        x = F.relu(self.bn(self.conv(frame, 'key_encoder.conv1', stride=2), 'key_encoder.bn1'))
        x = F.max_pool2d(x, 3, 2, 1)
        for i in range(3):
            x = self.bottleneck(x, f'key_encoder.res2.{i}', 1)
        f4 = x
        for i in range(4):
            x = self.bottleneck(x, f'key_encoder.layer2.{i}', 2 if i == 0 else 1)
        f8 = x
        for i in range(6):
            x = self.bottleneck(x, f'key_encoder.layer3.{i}', 2 if i == 0 else 1)
        f16 = x
        key = self.conv(f16, 'key_proj.key_proj')
        shrinkage = self.conv(f16, 'key_proj.d_proj') ** 2 + 1 if need_sk else None
        selection = torch.sigmoid(self.conv(f16, 'key_proj.e_proj')) if need_ek else None
        return key, shrinkage, selection, f16, f8, f4

    def encode_value(self, frame, f16, h16, masks, is_deep_update=True):
        """network.py:72-85, modules.py:124-150.  masks [B,n,H,W] (no background)."""
        n = masks.shape[1]
        if n != 1:
            others = torch.stack([masks[:, [j for j in range(n) if j != i]].sum(1) for i in range(n)], 1)
        else:
            others = torch.zeros_like(masks)
        g = torch.cat([frame.unsqueeze(1).expand(-1, n, -1, -1, -1), masks.unsqueeze(2), others.unsqueeze(2)], 2)
        b = g.shape[0]
        g = g.flatten(0, 1)
        g = self.bn(self.conv(g, 'value_encoder.conv1', stride=2), 'value_encoder.bn1')
        g = F.relu(F.max_pool2d(g, 3, 2, 1))          # maxpool THEN relu, modules.py:137-138
        for li, stride in (('layer1', 1), ('layer2', 2), ('layer3', 2)):
            for i in range(2):
                g = self.basic(g, f'value_encoder.{li}.{i}', stride if i == 0 else 1)
        g = g.view(b, n, *g.shape[1:])
        g = self.fusion(f16, g, 'value_encoder.fuser')
        if is_deep_update and self.hidden_dim > 0:
            h16 = self.gru(g, h16, 'value_encoder.hidden_reinforce')
        return g, h16

    def _upsample_block(self, skip, up_g, p):    # modules.py:178-191
        b, n = up_g.shape[:2]
        s = self.conv(skip, p + '.skip_conv')
        g = F.interpolate(up_g.flatten(0, 1), scale_factor=2, mode='bilinear', align_corners=False)
        g = g + s.unsqueeze(1).expand(-1, n, -1, -1, -1).flatten(0, 1)
        g = self.group_res(g, p + '.out_conv')
        return g.view(b, n, *g.shape[1:])

    def segment(self, feats, readout, hidden, h_out=True, strip_bg=True):
        """network.py:107-120 and Decoder.forward modules.py:229-250."""
        f16, f8, f4 = feats
        b, n = readout.shape[:2]
        g_in = torch.cat([readout, hidden], 2) if self.hidden_dim > 0 else readout
        g16 = self.fusion(f16, g_in, 'decoder.fuser')
        g8 = self._upsample_block(f8, g16, 'decoder.up_16_8')
        g4 = self._upsample_block(f4, g8, 'decoder.up_8_4')
        logits = self.conv(F.relu(g4.flatten(0, 1)), 'decoder.pred')
        new_h = None
        if h_out and self.hidden_dim > 0:           # HiddenUpdater modules.py:57-74
            g4c = torch.cat([g4, logits.view(b, n, 1, *logits.shape[-2:])], 2)
            def c1(t, name):
                o = self.conv(t.flatten(0, 1), name)
                return o.view(b, n, *o.shape[1:])
            def area(t, r):
                o = F.interpolate(t.flatten(0, 1), scale_factor=r, mode='area')
                return o.view(b, n, *o.shape[1:])
            g = c1(g16, 'decoder.hidden_update.g16_conv') + c1(area(g8, 1 / 2), 'decoder.hidden_update.g8_conv') \
                + c1(area(g4c, 1 / 4), 'decoder.hidden_update.g4_conv')
            new_h = self.gru(g, hidden, 'decoder.hidden_update')
        logits = F.interpolate(logits, scale_factor=4, mode='bilinear', align_corners=False)
        logits = logits.view(b, n, *logits.shape[-2:])
        prob = torch.sigmoid(logits)
        logits, prob = aggregate(prob, dim=1, return_logits=True)
        if strip_bg:
            prob = prob[:, 1:]
        return new_h, logits, prob


# ----------------------------------------------------------------------------------------------
# memory stores (inference/kv_memory_store.py)
# ----------------------------------------------------------------------------------------------
class OracleStore:
    """Growable key/value bank; layout as kv_memory_store.py:4-239 (k [1,CK,N], v[g] [n_g,CV,N_g])."""

    def __init__(self, count_usage: bool):
        self.count_usage = count_usage
        self.k = self.s = self.e = None
        self.v: List[Tensor] = []
        self.groups: List[List[int]] = []
        self.all_objects: List[int] = []
        self.use = self.life = None

    @property
    def size(self):
        return 0 if self.k is None else self.k.shape[-1]

    @property
    def num_groups(self):
        return len(self.v)

    def engaged(self):
        return self.k is not None

    def add(self, key, value, shrinkage, selection, objects):   # :36-94
        n_new = key.shape[2]
        cnt = torch.zeros((1, 1, n_new), device=key.device); life = torch.zeros((1, 1, n_new), device=key.device) + 1e-7
        if self.k is None:
            self.k, self.s, self.e = key, shrinkage, selection
            if self.count_usage:
                self.use, self.life = cnt, life
        else:
            self.k = torch.cat([self.k, key], -1)
            if shrinkage is not None:
                self.s = torch.cat([self.s, shrinkage], -1)
            if selection is not None:
                self.e = torch.cat([self.e, selection], -1)
            if self.count_usage:
                self.use = torch.cat([self.use, cnt], -1); self.life = torch.cat([self.life, life], -1)
        if objects is not None:
            rest = [o - 1 for o in objects]
            for gi, grp in enumerate(self.groups):
                for o in grp:
                    rest.remove(o)
                self.v[gi] = torch.cat([self.v[gi], value[grp]], -1)
            if rest:
                self.v.append(value[rest]); self.groups.append(list(rest)); self.all_objects.extend(rest)
                assert sorted(self.all_objects) == self.all_objects
        else:
            for gi, gv in enumerate(value):
                if gv is None:
                    continue
                if gi < self.num_groups:
                    self.v[gi] = torch.cat([self.v[gi], gv], -1)
                else:
                    self.v.append(gv)

    def update_usage(self, usage):                  # :96-103
        if self.count_usage:
            self.use = self.use + usage.view_as(self.use)
            self.life = self.life + 1

    def usage(self):                                # :183-189
        return self.use / self.life

    def remove_obsolete(self, max_size: int):       # :160-181
        u = self.usage().flatten()
        vals, _ = torch.topk(u, k=self.size - max_size, largest=False, sorted=True)
        keep = u > vals[-1]
        self.k = self.k[:, :, keep]; self.s = self.s[:, :, keep]
        if self.e is not None:
            self.e = self.e[:, :, keep]
        if self.num_groups > 1:
            raise NotImplementedError('feature removal with multiple object groups')
        self.v = [v[:, :, keep] for v in self.v]
        self.use = self.use[:, :, keep]; self.life = self.life[:, :, keep]


# ----------------------------------------------------------------------------------------------
# memory manager (inference/memory_manager.py)
# ----------------------------------------------------------------------------------------------
class OracleMemory:
    def __init__(self, config):
        self.cfg = config
        self.hidden_dim = config['hidden_dim']
        self.top_k = config['top_k']
        self.enable_long_term = config['enable_long_term']
        self.count_long_usage = config['enable_long_term_count_usage']
        if self.enable_long_term:
            self.max_mt = config['max_mid_term_frames']; self.min_mt = config['min_mid_term_frames']
            self.n_proto = config['num_prototypes']; self.max_long = config['max_long_term_elements']
        self.temp = OracleStore(self.enable_long_term)
        self.perm = OracleStore(False)
        self.long = OracleStore(self.count_long_usage) if self.enable_long_term else None
        self.hidden = None
        self.HW = None

    # memory_manager.py:61-190
    def match(self, qk, qe, no_usage=False):
        h, w = qk.shape[-2:]
        qk = qk.flatten(2); qe = qe.flatten(2) if qe is not None else None
        G = max(self.temp.num_groups, self.perm.num_groups)
        T = self.temp.size
        use_long = self.enable_long_term and self.long.engaged()
        banks = ([self.long] if use_long else []) + [self.temp, self.perm]
        L = self.long.size if use_long else 0
        sim = similarity(torch.cat([b.k for b in banks], -1), torch.cat([b.s for b in banks], -1), qk, qe)
        parts = [sim[:, :L], sim[:, L:L + T], sim[:, L + T:]] if use_long else [sim[:, :T], sim[:, T:]]
        affs, vals = [], []
        usage = None
        for gi in range(G):
            cols, vv = [], []
            for b, p in zip(banks, parts):
                if b is self.long and gi >= b.num_groups:
                    continue
                ng = b.v[gi].shape[-1]
                cols.append(p[:, p.shape[1] - ng:])
                vv.append(b.v[gi])
            s_g = torch.cat(cols, 1)
            if gi == 0 and self.enable_long_term:
                a, usage = softmax_topk(s_g, self.top_k, want_usage=True)
            else:
                a = softmax_topk(s_g, self.top_k)
            affs.append(a); vals.append(torch.cat(vv, -1))
        if usage is not None and not no_usage:
            if use_long:
                # group 0 of long memory always spans all long columns (:93-95)
                self.temp.update_usage(usage[:, L:L + T].flatten())
                if self.count_long_usage:
                    self.long.update_usage(usage[:, :L].flatten())
            else:
                self.temp.update_usage(usage[:, :T].flatten())
        out = torch.cat([v @ a for v, a in zip(vals, affs)], 0)
        return out.view(out.shape[0], -1, h, w)

    # memory_manager.py:212-281
    def add(self, key, shrinkage, value, objects, selection=None, permanent=False, ignore=False):
        if self.HW is None:
            self.HW = key.shape[-2] * key.shape[-1]
        key = key.flatten(2); shrinkage = shrinkage.flatten(2); value = value[0].flatten(2)
        if selection is not None:
            selection = selection.flatten(2)
        if not ignore:
            (self.perm if permanent else self.temp).add(key, value, shrinkage, selection, objects)
        if (not self.temp.engaged()) or self.temp.num_groups != self.perm.num_groups:
            z = lambda t: t[..., 0:0]
            tgt = self.temp if self.perm.num_groups > self.temp.num_groups else self.perm
            tgt.add(z(key), z(value), z(shrinkage), z(selection), objects)
        if self.enable_long_term and self.temp.size >= self.max_mt * self.HW:
            if self.long.size >= self.max_long - self.n_proto:
                self.long.remove_obsolete(self.max_long - self.n_proto)
            self.compress()

    # memory_manager.py:316-347
    def compress(self):
        HW = self.HW
        m = self.min_mt * HW
        total = self.temp.size
        cand_v = []
        for gv in self.temp.v:
            ng = gv.shape[-1]
            if ng == total or ng > m:
                cand_v.append(gv[:, :, :ng - m])
            else:
                cand_v.append(None)
        ck = self.temp.k[:, :, :total - m]; cs = self.temp.s[:, :, :total - m]
        ce = self.temp.e[:, :, :total - m] if self.temp.e is not None else None
        cu = self.temp.usage()[:, :, :total - m]
        pk, pv, ps = self.consolidate(ck, cs, ce, cu, cand_v)
        # sieve_by_range(0, -m, min_size=m+HW) (kv_memory_store.py:125-158)
        t = self.temp
        t.k = t.k[:, :, total - m:]; t.s = t.s[:, :, total - m:]
        if t.e is not None:
            t.e = t.e[:, :, total - m:]
        t.use = t.use[:, :, total - m:]; t.life = t.life[:, :, total - m:]
        for gi in range(t.num_groups):
            if t.v[gi].shape[-1] >= m + HW:
                t.v[gi] = t.v[gi][:, :, t.v[gi].shape[-1] - m:]
        self.long.add(pk, pv, ps, None, None)

    # memory_manager.py:349-390
    def consolidate(self, ck, cs, ce, usage, cand_v):
        N = ck.shape[-1]
        _, idx = torch.topk(usage, k=self.n_proto, dim=-1, sorted=True)
        idx = idx.flatten()
        valid = [idx >= (N - gv.shape[2]) if gv is not None else None for gv in cand_v]
        pk = ck[:, :, idx]
        pe = ce[:, :, idx] if ce is not None else None
        sim = similarity(ck, cs, pk, pe)
        affs = [softmax_topk(sim[:, N - gv.shape[2]:, valid[gi]], None) if gv is not None else None
                for gi, gv in enumerate(cand_v)]
        affs = [a if a is None or a.shape[-1] > 0 else None for a in affs]
        pv = [gv @ affs[gi] if affs[gi] is not None else None for gi, gv in enumerate(cand_v)]
        ps = cs @ affs[0] if cs is not None else None
        return pk, pv, ps

    def ensure_hidden(self, n, sample_key):         # :283-296
        h, w = sample_key.shape[-2:]
        if self.hidden is None:
            self.hidden = torch.zeros((1, n, self.hidden_dim, h, w), device=sample_key.device)
        elif self.hidden.shape[1] != n:
            self.hidden = torch.cat([self.hidden, torch.zeros((1, n - self.hidden.shape[1], self.hidden_dim, h, w), device=sample_key.device)], 1)


# ----------------------------------------------------------------------------------------------
# per-frame state machine (inference/inference_core.py)
# ----------------------------------------------------------------------------------------------
class OracleCore:
    def __init__(self, net: OracleNet, config):
        self.net = net
        self.cfg = config
        self.mem_every = config['mem_every']
        self.deep_every = config['deep_update_every']
        self.enable_long_term = config['enable_long_term']
        self.deep_sync = self.deep_every < 0
        self.ti = -1
        self.last_mem_ti = 0
        if not self.deep_sync:
            self.last_deep_ti = -self.deep_every
        self.mem = OracleMemory(config)
        self.labels = None

    def set_all_labels(self, labels):
        self.labels = labels

    # inference_core.py:154-179
    def put_to_permanent_memory(self, image, mask):
        image, self.pad = pad_to_16(image); image = image.unsqueeze(0)
        key, shr, sel, f16, _, _ = self.net.encode_key(image)
        mask, _ = pad_to_16(mask)
        prob = aggregate(mask, dim=0)
        self.mem.ensure_hidden(len(self.labels), key)
        value, _ = self.net.encode_value(image, f16, self.mem.hidden, prob[1:].unsqueeze(0), is_deep_update=False)
        self.mem.add(key, shr, value, self.labels, selection=sel if self.enable_long_term else None, permanent=True)

    # inference_core.py:62-152
    def step(self, image, mask=None, valid_labels=None, end=False, manually_curated_masks=False,
             disable_memory_updates=False, do_not_add_mask_to_memory=False):
        self.ti += 1
        image, self.pad = pad_to_16(image); image = image.unsqueeze(0)
        if manually_curated_masks:
            is_mem = (mask is not None) and not end
        else:
            is_mem = ((self.ti - self.last_mem_ti >= self.mem_every) or (mask is not None)) and not end
        need_seg = (valid_labels is None) or (len(self.labels) != len(valid_labels))
        is_deep = ((self.deep_sync and is_mem) or
                   (not self.deep_sync and self.ti - self.last_deep_ti >= self.deep_every)) and not end
        is_normal = (not self.deep_sync or not is_deep) and not end
        key, shr, sel, f16, f8, f4 = self.net.encode_key(image, need_ek=(self.enable_long_term or need_seg))
        if disable_memory_updates:
            is_normal = is_deep = is_mem = False
            self.ti -= 1
        prob = prob_nobg = None
        if need_seg:
            ro = self.mem.match(key, sel, no_usage=disable_memory_updates).unsqueeze(0)
            hid, _, prob = self.net.segment((f16, f8, f4), ro, self.mem.hidden, h_out=is_normal, strip_bg=False)
            prob = prob[0]; prob_nobg = prob[1:]
            if is_normal:
                self.mem.hidden = hid
        if mask is not None:
            mask, _ = pad_to_16(mask)
            if prob_nobg is not None:
                region = mask.sum(0) > 0.5
                prob_nobg[:, region] = 0
                mask = mask.type_as(prob_nobg)
                if valid_labels is not None:
                    keep = [i for i in range(prob_nobg.shape[0]) if (i + 1) not in valid_labels]
                    mask[keep] = prob_nobg[keep]
            prob = aggregate(mask, dim=0)
            if not disable_memory_updates:
                self.mem.ensure_hidden(len(self.labels), key)
        if is_mem:
            value, hid = self.net.encode_value(image, f16, self.mem.hidden, prob[1:].unsqueeze(0), is_deep_update=is_deep)
            self.mem.add(key, shr, value, self.labels, selection=sel if self.enable_long_term else None,
                         ignore=do_not_add_mask_to_memory)
            self.last_mem_ti = self.ti
            if is_deep:
                self.mem.hidden = hid
                self.last_deep_ti = self.ti
        return unpad(prob, self.pad)


# ----------------------------------------------------------------------------------------------
# annotation-candidate selector  (inference/frame_selection/frame_selection.py:99-244) — SURVEY.md 8f row 3,
# the second consumer of the similarity of memory_util.py:7-39 (no softmax)
# ----------------------------------------------------------------------------------------------
def cycle_dissimilarity(key_a: Tensor, shr_a: Tensor, sel_a: Tensor, key_b: Tensor, shr_b: Tensor, sel_b: Tensor) -> Tensor:
    """frame_selection.py:213-221 for one ordered pair (A = already chosen frame, B = candidate).
    key [CK,h,w] (composite key), shr [1,h,w], sel [CK,h,w].  mean over [HW,HW] of relu(S_ab - S_ba) where
    S_ab[n,q] = similarity(memory = A's pixel n, query = B's pixel q weighted by B's selection) and
    S_ba[n,q] = similarity(memory = B's pixel n, query = A's pixel q weighted by A's selection)."""
    s_ab = similarity(key_a.unsqueeze(0), shr_a.unsqueeze(0), key_b.unsqueeze(0), sel_b.unsqueeze(0))
    s_ba = similarity(key_b.unsqueeze(0), shr_b.unsqueeze(0), key_a.unsqueeze(0), sel_a.unsqueeze(0))
    d = (s_ab - s_ba).float()
    return F.relu(d).sum() / d.numel()


def composite_keys_for_selection(keys: Tensor, masks: List[Tensor], previously_chosen: List[int], alpha: float,
                                 min_mask_presence_percent: float, epsilon: float):
    """frame_selection.py:156-189: per-frame validity (mask covers at least `min_mask_presence_percent` PERCENT of the
    frame, previously chosen frames are always valid) and the mask-weighted key  key*(alpha*any_object_mask + 1-alpha),
    mask resized to the key grid with nearest-neighbour sampling."""
    n = len(keys)
    h, w = keys[0].shape[-2:]
    valid, comp = [], []
    for i in range(n):
        m = masks[i] if masks[i].ndim == 3 else masks[i].unsqueeze(0)
        m_bin = m.max(dim=0).values
        ratio = (m_bin > epsilon).sum() / m_bin.numel() * 100
        if ratio < min_mask_presence_percent and i not in previously_chosen:
            valid.append(False); comp.append(None)
            continue
        small = F.interpolate(m.unsqueeze(0).float(), size=(h, w), mode='nearest')[0]
        ck = keys[i] * small.max(dim=0, keepdim=True).values
        ck = ck * alpha + keys[i] * (1 - alpha)
        valid.append(True); comp.append(ck.to(keys[i].dtype))
    return valid, comp


def select_next_candidates(keys: Tensor, shrinkages: Tensor, selections: Tensor, masks: List[Tensor], num_next_candidates: int,
                           previously_chosen_candidates=(0,), alpha: float = 0.5, min_mask_presence_percent: float = 0.25,
                           only_new_candidates: bool = True, epsilon: float = 0.5) -> List[int]:
    """frame_selection.py:99-244: greedily add the frame whose SMALLEST cycle dissimilarity to the already chosen
    frames is the largest.  keys [N,CK,h,w], shrinkages [N,1,h,w], selections [N,CK,h,w], masks: N tensors [C,H,W]."""
    n = len(keys)
    chosen = list(previously_chosen_candidates)
    valid, comp = composite_keys_for_selection(keys, masks, chosen, alpha, min_mask_presence_percent, epsilon)
    for _ in range(num_next_candidates):
        scores = []
        for j in range(n):
            if not valid[j]:
                scores.append(torch.tensor(0.0))
                continue
            scores.append(min(cycle_dissimilarity(comp[c], shrinkages[c], selections[c], comp[j], shrinkages[j], selections[j])
                              for c in chosen))
        chosen.append(int(torch.argmax(torch.stack([torch.as_tensor(s, dtype=torch.float32) for s in scores]))))
    return chosen[len(previously_chosen_candidates):] if only_new_candidates else chosen
