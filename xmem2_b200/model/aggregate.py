"""Soft aggregation of per-object probabilities (reference model/aggregate.py:6-16).  Used by
InferenceCore for user-provided masks; the decoder path fuses the same math into
xm_upsample4x_aggregate (csrc/eltwise.cu)."""
import torch
import torch.nn.functional as F


def aggregate(prob, dim, return_logits=False):
    bg = torch.prod(1 - prob, dim=dim, keepdim=True)
    new_prob = torch.cat([bg, prob], dim).clamp(1e-7, 1 - 1e-7)
    logits = torch.log(new_prob / (1 - new_prob))
    prob = F.softmax(logits, dim=dim)
    return (logits, prob) if return_logits else prob
