"""Per-frame state machine — drop-in for the reference `inference/inference_core.py:11-186`.

Same constructor, attributes (`memory, network, config, mem_every, deep_update_every, enable_long_term,
curr_ti, last_mem_ti, all_labels, pad`) and methods (`step, put_to_permanent_memory, clear_memory,
update_config, set_all_labels, encode_frame_key, remove_from_permanent_memory, permanent_memory_frames`).
The frame scheduling rules are those of `InferenceCore.step` (:62-152); all heavy work is delegated to
`XMem` (tcgen05 conv kernels) and `MemoryManager` (fused read kernel).
"""
from __future__ import annotations

import os
import time
import weakref

import torch

from .. import lib
from ..model.aggregate import aggregate
from ..model.network import XMem
from ..util.tensor_util import pad_amounts, pad_divide_by, unpad
from .memory_manager import MemoryManager


# Recorded frame graphs are shared by every InferenceCore built on the same network (the reference driver builds one
# core per video): they live in a dict attached to the network object (so they die with its weights), keyed by the
# graph signature; value = graph + its static buffers.  Arena recycling in KeyValueMemoryStore makes consecutive videos
# of the same shape hit the same signature.
_GRAPH_CACHE_MAX = 8


def _graph_cache(network) -> dict:
    cache = network.__dict__.get('_xm_graph_cache')
    if cache is None:
        cache = {}
        network.__dict__['_xm_graph_cache'] = cache
    return cache


class InferenceCore:
    def __init__(self, network: XMem, config):
        self.config = config
        self.network = network
        self.mem_every = config['mem_every']
        self.deep_update_every = config['deep_update_every']
        self.enable_long_term = config['enable_long_term']
        self.deep_update_sync = self.deep_update_every < 0     # deep update rides on memory frames
        self.clear_memory()
        self.all_labels = None
        # steady-state frames (no mask, not a memory frame) are replayed from a CUDA graph: ~90 kernel launches per
        # frame would otherwise be bound by host launch latency, not by the GPU.  `use_cuda_graph=False` disables it.
        # `t_shard=True` (extension): the memory of this ONE video is sharded over the ranks of
        # config['t_shard_group'] (default: the world); every rank runs the same frames.  Collectives -> no graph replay.
        self.use_cuda_graph = config.get('use_cuda_graph', True) and not config.get('t_shard', False)
        self._graphs = {}
        self._graph_warm = set()
        self._g_out = None
        # warm-up on the network's own device (the reference hard-codes cuda:0, inference_core.py:26)
        dev = next(network.parameters()).device
        if dev.type == 'cuda':
            self.network.encode_key(torch.zeros((1, 3, 480, 864), device=dev))

    def clear_memory(self, keep_permanent=False):
        self.curr_ti = -1
        self.last_mem_ti = 0
        if not self.deep_update_sync:
            self.last_deep_update_ti = -self.deep_update_every
        if self.config.get('t_shard', False):
            if keep_permanent:
                raise NotImplementedError('keep_permanent with a T-sharded memory')
            from .sharded_memory import ShardedMemoryManager
            self.memory = ShardedMemoryManager(self.config, group=self.config.get('t_shard_group'))
        else:
            self.memory = self.memory.copy_perm_mem_only() if keep_permanent else MemoryManager(config=self.config)
        self._graphs = {}

    def update_config(self, config):
        self.mem_every = config['mem_every']
        self.deep_update_every = config['deep_update_every']
        self.enable_long_term = config['enable_long_term']
        self.deep_update_sync = self.deep_update_every < 0
        self.memory.update_config(config)

    def set_all_labels(self, all_labels):
        self.all_labels = all_labels

    def _prepare(self, image):
        image, self.pad = pad_divide_by(image, 16)
        return image.unsqueeze(0)

    def encode_frame_key(self, image):
        key, shrinkage, selection, _, _, _ = self.network.encode_key(self._prepare(image), need_ek=True, need_sk=True)
        return key, shrinkage, selection

    def _schedule(self, has_mask, end, manually_curated_masks):
        """frame flags of inference_core.py:75-87."""
        if manually_curated_masks:
            is_mem = has_mask and not end
        else:
            is_mem = ((self.curr_ti - self.last_mem_ti >= self.mem_every) or has_mask) and not end
        if self.deep_update_sync:
            is_deep = is_mem and not end
        else:
            is_deep = (self.curr_ti - self.last_deep_update_ti >= self.deep_update_every) and not end
        is_normal = (not self.deep_update_sync or not is_deep) and not end
        return is_mem, is_deep, is_normal

    def step(self, image, mask=None, valid_labels=None, end=False, manually_curated_masks=False,
             disable_memory_updates=False, do_not_add_mask_to_memory=False, return_key_and_stuff=False):
        """image 3xHxW (normalised), mask num_objects x H x W or None -> probabilities (num_objects+1) x H x W."""
        self.curr_ti += 1
        raw = image
        is_mem_frame, is_deep_update, is_normal_update = self._schedule(mask is not None, end, manually_curated_masks)
        need_segment = (valid_labels is None) or (len(self.all_labels) != len(valid_labels))

        if (self.use_cuda_graph and mask is None and need_segment and not end and not disable_memory_updates
                and not return_key_and_stuff and raw.is_cuda and raw.dim() == 3 and self.memory.get_hidden() is not None):
            # steady-state frames: the unpadded frame is copied straight into the interior of the graph's (zero-bordered) input
            # buffer -- no F.pad, no intermediate copy
            if is_normal_update and not is_mem_frame:
                prob = self._graph_step(raw, mem_frame=False)
                if prob is not None:
                    return unpad(prob, self.pad).clone()      # the graph's output buffer is rewritten by the next replay
            elif is_mem_frame and is_deep_update and self.deep_update_sync and not do_not_add_mask_to_memory:
                prob = self._graph_step(raw, mem_frame=True)
                if prob is not None:
                    # the recorded graph produced key/shrinkage/value/selection and the deep-updated hidden state;
                    # the arena append has a moving offset and stays eager (inference_core.py:136-145)
                    self.memory.add_memory(self._g_out['key'], self._g_out['shrinkage'], self._g_out['value'], self.all_labels,
                                           selection=self._g_out['selection'] if self.enable_long_term else None, ignore=False)
                    self.last_mem_ti = self.curr_ti
                    self.last_deep_update_ti = self.curr_ti
                    return unpad(prob, self.pad).clone()

        image = self._prepare(raw)
        if (self.use_cuda_graph and mask is not None and not need_segment and is_mem_frame and not disable_memory_updates
                and not return_key_and_stuff and image.is_cuda and mask.shape[0] == len(self.all_labels)
                and (is_deep_update == self.deep_update_sync)):
            # fully annotated frame: encode_key -> aggregate(mask) -> encode_value, replayed from a recorded graph
            mask_p, _ = pad_divide_by(mask, 16)
            self.memory.create_hidden_state(len(self.all_labels), image[..., ::16, ::16])
            g = self._encode_graph(image, mask_p, deep=is_deep_update)
            if g is not None:
                self.memory.add_memory(g['key'], g['shrinkage'], g['value'], self.all_labels,
                                       selection=g['selection'] if self.enable_long_term else None, ignore=do_not_add_mask_to_memory)
                self.last_mem_ti = self.curr_ti
                if is_deep_update:
                    self.memory.set_hidden(g['hidden_out'].clone())
                    self.last_deep_update_ti = self.curr_ti
                return unpad(g['pred'], self.pad).clone()

        key, shrinkage, selection, f16, f8, f4 = self.network.encode_key(
            image, need_ek=(self.enable_long_term or need_segment), need_sk=True)

        if disable_memory_updates:
            is_normal_update = is_deep_update = is_mem_frame = False
            self.curr_ti -= 1

        pred_prob_with_bg = pred_prob_no_bg = None
        if need_segment:
            readout = self.memory.match_memory(key, selection, disable_usage_updates=disable_memory_updates).unsqueeze(0)
            hidden, _, prob = self.network.segment((f16, f8, f4), readout, self.memory.get_hidden(),
                                                   h_out=is_normal_update, strip_bg=False)
            pred_prob_with_bg = prob[0]
            pred_prob_no_bg = pred_prob_with_bg[1:]
            if is_normal_update:
                self.memory.set_hidden(hidden)

        if mask is not None:
            mask, _ = pad_divide_by(mask, 16)
            if pred_prob_no_bg is not None:
                # make the prediction consistent with the user mask (inference_core.py:117-127)
                pred_prob_no_bg[:, mask.sum(0) > 0.5] = 0
                mask = mask.type_as(pred_prob_no_bg)
                if valid_labels is not None:
                    unlabelled = [i for i in range(pred_prob_no_bg.shape[0]) if (i + 1) not in valid_labels]
                    if unlabelled:
                        mask = mask.clone()      # never write into the caller's tensor (pad_divide_by / type_as may alias it)
                        mask[unlabelled] = pred_prob_no_bg[unlabelled]
            pred_prob_with_bg = aggregate(mask, dim=0)
            if not disable_memory_updates:
                self.memory.create_hidden_state(len(self.all_labels), key)

        if is_mem_frame:
            value, hidden = self.network.encode_value(image, f16, self.memory.get_hidden(),
                                                      pred_prob_with_bg[1:].unsqueeze(0), is_deep_update=is_deep_update)
            self.memory.add_memory(key, shrinkage, value, self.all_labels,
                                   selection=selection if self.enable_long_term else None, ignore=do_not_add_mask_to_memory)
            self.last_mem_ti = self.curr_ti
            if is_deep_update:
                self.memory.set_hidden(hidden)
                self.last_deep_update_ti = self.curr_ti

        res = unpad(pred_prob_with_bg, self.pad)
        if return_key_and_stuff:
            return res, key, shrinkage, selection
        return res

    # ------------------------------------------------------------------ CUDA-graph replay of steady-state frames
    def _graph_step(self, raw, mem_frame):
        """One frame replayed from a recorded CUDA graph.
          mem_frame=False: encode_key -> match_memory -> segment(h_out=True) -> hidden update        (ordinary frame)
          mem_frame=True : encode_key -> match_memory -> segment(h_out=False) -> encode_value(deep)  (memory frame)
        The graph depends on the image shape and on the memory arenas' addresses/capacities and group structure
        (MemoryManager.layout_signature); bank SIZES live in a device-side plan that is refreshed (stream ordered)
        whenever a memory frame changed them.  Returns None when this frame must run eagerly (first frame with a new
        signature = warm-up of lazily initialised kernel state)."""
        mem = self.memory
        self.pad = pad_amounts(raw, 16)
        lw, uw, lh, uh = self.pad
        hr, wr = raw.shape[-2:]
        shape = (1, raw.shape[0], hr + lh + uh, wr + lw + uw)             # the padded frame the network sees
        sig = (shape, mem.layout_signature(), len(self.all_labels), bool(mem_frame))
        g = self._graphs.get(sig)
        if g is None:
            cache = _graph_cache(self.network)
            g = cache.get(sig)
            if g is None:
                if sig not in self._graph_warm:
                    self._graph_warm.add(sig)           # run this frame eagerly, record on the next one
                    return None
                g = self._capture(self._prepare(raw), mem_frame)
                if len(cache) >= _GRAPH_CACHE_MAX:
                    cache.pop(next(iter(cache)))
                cache[sig] = g
            self._graphs = {sig: g, **{k: v for k, v in self._graphs.items() if k[:3] == sig[:3]}}
        # the graph's hidden-state buffer is shared by every core using this graph: hand the previous user its own copy
        g_hidden = g['hidden']
        prev = g['owner'][0]() if g['owner'][0] is not None else None
        if prev is not None and prev is not self:
            ph = prev.memory.get_hidden()
            if ph is not None and ph.data_ptr() == g_hidden.data_ptr():
                prev.memory.set_hidden(ph.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3))
        g['owner'][0] = weakref.ref(self)
        # the recorded read kernels have the K1 workspace address (device-side plan + scratch) baked in: a core that
        # inherits a graph recorded by an earlier core (same recycled arenas -> same signature) adopts that workspace
        if mem._ws is not g['ws']:
            mem.adopt_workspace(g['ws'])
        hid = mem.get_hidden()
        if hid.data_ptr() != g_hidden.data_ptr():
            g_hidden.copy_(hid)
            mem.set_hidden(g_hidden)
        g['image'][0, :, lh:lh + hr, lw:lw + wr].copy_(raw)                 # borders stay zero (F.pad of the recording frame)
        h, w = shape[-2] // 16, shape[-1] // 16
        mem.upload_plan(h * w, raw.device)
        g['graph'].replay()
        lib.load().xm_add_launch_count(g['launches'])
        self._g_out = g
        return g['prob']

    def _capture(self, image, mem_frame):
        t0 = time.perf_counter()
        mem, net = self.memory, self.network
        dev = image.device
        n = len(self.all_labels)
        h, w = image.shape[-2] // 16, image.shape[-1] // 16
        g = {'image': image.clone()}
        # all graphs of one memory layout share ONE hidden-state buffer (ordinary and memory frames alternate) and
        # therefore one owner record
        src = next(iter(self._graphs.values()), None)
        if src is not None and src['hidden'].shape[1] == n:
            shared, g['owner'] = src['hidden'], src['owner']
        else:
            shared, g['owner'] = torch.zeros((1, n, h, w, mem.hidden_dim), device=dev).permute(0, 1, 4, 2, 3), [None]
        g['hidden'] = shared
        if mem.get_hidden().data_ptr() != shared.data_ptr():
            shared.copy_(mem.get_hidden())
            mem.set_hidden(shared)
        mem.upload_plan(h * w, dev)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        launches0 = lib.load().xm_launch_count()
        with torch.cuda.graph(graph):
            key, shrinkage, selection, f16, f8, f4 = net.encode_key(g['image'], need_ek=True, need_sk=True)
            readout = mem.match_memory(key, selection).unsqueeze(0)
            hidden, _, prob = net.segment((f16, f8, f4), readout, shared, h_out=not mem_frame, strip_bg=False)
            g['prob'] = prob[0]
            if mem_frame:
                value, hidden = net.encode_value(g['image'], f16, shared, prob[0][1:].unsqueeze(0), is_deep_update=True)
                g.update(key=key, shrinkage=shrinkage, selection=selection, value=value)
            shared.copy_(hidden)
        g['graph'] = graph
        g['ws'] = mem._ws                       # keeps the workspace the recorded kernels point into alive
        g['launches'] = int(lib.load().xm_launch_count() - launches0)     # recorded, not executed
        lib.load().xm_add_launch_count(-g['launches'])
        if os.environ.get('XMEM_TRACE'):
            torch.cuda.synchronize(dev)
            print(f'[xmem2_b200] recorded {"memory" if mem_frame else "ordinary"}-frame graph in {time.perf_counter() - t0:.3f}s', flush=True)
        return g

    def _encode_graph(self, image, mask_padded, deep):
        """encode_key -> aggregate(mask) -> encode_value for a fully annotated frame (inference_core.py:128-137,157-167),
        replayed from a recorded CUDA graph.  Independent of the memory banks, so one graph per (shape, #objects, deep)
        serves every video on this network.  deep=True also runs the HiddenReinforcer on the current hidden state.
        Returns the dict of static outputs, or None on the first (warm-up) call of a signature."""
        n = mask_padded.shape[0]
        sig = (tuple(image.shape), n, 'enc', bool(deep))
        cache = _graph_cache(self.network)
        g = cache.get(sig)
        if g is None:
            if sig not in self._graph_warm:
                self._graph_warm.add(sig)
                return None
            net, mem, dev = self.network, self.memory, image.device
            h, w = image.shape[-2] // 16, image.shape[-1] // 16
            g = {'image': image.clone(), 'mask': mask_padded.clone().float()}
            if deep:
                g['hidden_in'] = torch.zeros((1, n, h, w, mem.hidden_dim), device=dev).permute(0, 1, 4, 2, 3)
                g['hidden_out'] = torch.zeros((1, n, h, w, mem.hidden_dim), device=dev).permute(0, 1, 4, 2, 3)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            launches0 = lib.load().xm_launch_count()
            with torch.cuda.graph(graph):
                key, shrinkage, selection, f16, _, _ = net.encode_key(g['image'], need_ek=True, need_sk=True)
                pred = aggregate(g['mask'], dim=0)
                value, hidden = net.encode_value(g['image'], f16, g.get('hidden_in'), pred[1:].unsqueeze(0), is_deep_update=deep)
                if deep:
                    g['hidden_out'].copy_(hidden)
                g.update(key=key, shrinkage=shrinkage, selection=selection, value=value, pred=pred)
            g['graph'] = graph
            g['launches'] = int(lib.load().xm_launch_count() - launches0)
            lib.load().xm_add_launch_count(-g['launches'])
            if len(cache) >= _GRAPH_CACHE_MAX:
                cache.pop(next(iter(cache)))
            cache[sig] = g
        g['image'].copy_(image)
        g['mask'].copy_(mask_padded)
        if deep:
            g['hidden_in'].copy_(self.memory.get_hidden())
        g['graph'].replay()
        lib.load().xm_add_launch_count(g['launches'])
        return g

    def put_to_permanent_memory(self, image, mask, ti=None):
        """encode an annotated frame straight into permanent memory (inference_core.py:154-179)."""
        image = self._prepare(image)
        if self.use_cuda_graph and image.is_cuda and mask.shape[0] == len(self.all_labels):
            mask_p, _ = pad_divide_by(mask, 16)
            g = self._encode_graph(image, mask_p, deep=False)
            if g is not None:
                self.memory.create_hidden_state(len(self.all_labels), g['key'])
                sel = g['selection'] if self.enable_long_term else None
                is_update = self.memory.frame_already_saved(ti)
                if is_update:
                    self.memory.update_permanent_memory(ti, g['key'], g['shrinkage'], g['value'], selection=sel)
                else:
                    self.memory.add_memory(g['key'], g['shrinkage'], g['value'], self.all_labels, selection=sel, permanent=True, ti=ti)
                return is_update
        key, shrinkage, selection, f16, _, _ = self.network.encode_key(image, need_ek=True, need_sk=True)
        mask, _ = pad_divide_by(mask, 16)
        pred_prob_with_bg = aggregate(mask, dim=0)
        self.memory.create_hidden_state(len(self.all_labels), key)
        value, _ = self.network.encode_value(image, f16, self.memory.get_hidden(), pred_prob_with_bg[1:].unsqueeze(0),
                                             is_deep_update=False)
        sel = selection if self.enable_long_term else None
        is_update = self.memory.frame_already_saved(ti)
        if is_update:
            self.memory.update_permanent_memory(ti, key, shrinkage, value, selection=sel)
        else:
            self.memory.add_memory(key, shrinkage, value, self.all_labels, selection=sel, permanent=True, ti=ti)
        return is_update

    def remove_from_permanent_memory(self, frame_idx):
        self.memory.remove_from_permanent_memory(frame_idx)

    @property
    def permanent_memory_frames(self):
        return list(self.memory.frame_id_to_permanent_mem_idx.keys())
