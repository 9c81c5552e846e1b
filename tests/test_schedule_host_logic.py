"""CPU test of the frame-scheduling rules of InferenceCore (reference inference/inference_core.py:75-87): which frames are
memory frames, which trigger a deep hidden update, which a normal one — for the synchronised (deep_update_every < 0) and
the free-running deep-update modes, with and without masks, `end` and `manually_curated_masks`."""
import itertools

from xmem2_b200.inference.inference_core import InferenceCore


def _expected(curr_ti, last_mem_ti, last_deep_ti, mem_every, deep_every, has_mask, end, curated):
    deep_sync = deep_every < 0
    if curated:
        is_mem = has_mask and not end
    else:
        is_mem = ((curr_ti - last_mem_ti >= mem_every) or has_mask) and not end
    is_deep = ((deep_sync and is_mem) or (not deep_sync and curr_ti - last_deep_ti >= deep_every)) and not end
    is_normal = (not deep_sync or not is_deep) and not end
    return is_mem, is_deep, is_normal


def test_schedule_matches_reference_rules():
    core = InferenceCore.__new__(InferenceCore)          # no network / device needed for the scheduling rules
    for mem_every, deep_every in ((10, -1), (5, 3), (1, -1), (3, 1)):
        core.mem_every, core.deep_update_every = mem_every, deep_every
        core.deep_update_sync = deep_every < 0
        for curr, last_mem, last_deep, has_mask, end, curated in itertools.product(
                range(0, 12), (0, 4), (-3, 2), (False, True), (False, True), (False, True)):
            core.curr_ti, core.last_mem_ti, core.last_deep_update_ti = curr, last_mem, last_deep
            assert core._schedule(has_mask, end, curated) == _expected(curr, last_mem, last_deep, mem_every, deep_every,
                                                                       has_mask, end, curated)


def test_pad_amounts_matches_pad_divide_by():
    # the steady-state graph path copies the unpadded frame into the interior of a zero-bordered buffer using pad_amounts();
    # it must describe exactly the padding pad_divide_by (reference util/tensor_util.py:21-35) applies
    import torch
    from xmem2_b200.util.tensor_util import pad_amounts, pad_divide_by, unpad
    for h, w in ((480, 854), (853, 480), (1080, 1920), (96, 128), (17, 33), (16, 16)):
        x = torch.arange(3 * h * w, dtype=torch.float32).reshape(3, h, w)
        padded, pads = pad_divide_by(x, 16)
        assert pads == pad_amounts(x, 16)
        lw, uw, lh, uh = pads
        assert padded.shape[-2] % 16 == 0 and padded.shape[-1] % 16 == 0
        buf = torch.zeros_like(padded)
        buf[:, lh:lh + h, lw:lw + w].copy_(x)
        assert torch.equal(buf, padded)
        assert torch.equal(unpad(padded, pads), x)
