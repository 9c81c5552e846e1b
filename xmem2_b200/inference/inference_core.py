"""Per-frame state machine — drop-in for the reference `inference/inference_core.py:11-186`.

Same constructor, attributes (`memory, network, config, mem_every, deep_update_every, enable_long_term,
curr_ti, last_mem_ti, all_labels, pad`) and methods (`step, put_to_permanent_memory, clear_memory,
update_config, set_all_labels, encode_frame_key, remove_from_permanent_memory, permanent_memory_frames`).
The frame scheduling rules are those of `InferenceCore.step` (:62-152); all heavy work is delegated to
`XMem` (tcgen05 conv kernels) and `MemoryManager` (fused read kernel).
"""
from __future__ import annotations

import torch

from ..model.aggregate import aggregate
from ..model.network import XMem
from ..util.tensor_util import pad_divide_by, unpad
from .memory_manager import MemoryManager


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
        # warm-up on the network's own device (the reference hard-codes cuda:0, inference_core.py:26)
        dev = next(network.parameters()).device
        if dev.type == 'cuda':
            self.network.encode_key(torch.zeros((1, 3, 480, 864), device=dev))

    def clear_memory(self, keep_permanent=False):
        self.curr_ti = -1
        self.last_mem_ti = 0
        if not self.deep_update_sync:
            self.last_deep_update_ti = -self.deep_update_every
        self.memory = self.memory.copy_perm_mem_only() if keep_permanent else MemoryManager(config=self.config)

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
        image = self._prepare(image)
        is_mem_frame, is_deep_update, is_normal_update = self._schedule(mask is not None, end, manually_curated_masks)
        need_segment = (valid_labels is None) or (len(self.all_labels) != len(valid_labels))

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

    def put_to_permanent_memory(self, image, mask, ti=None):
        """encode an annotated frame straight into permanent memory (inference_core.py:154-179)."""
        image = self._prepare(image)
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
