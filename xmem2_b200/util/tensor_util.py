"""pad / unpad helpers with the reference's semantics (util/tensor_util.py:47-77)."""
import torch.nn.functional as F


def pad_amounts(in_img, d):
    """(left, right, top, bottom) padding that makes H and W multiples of d -- the pad_array of pad_divide_by (util/tensor_util.py:21-35)"""
    h, w = in_img.shape[-2:]
    new_h = h if h % d == 0 else h + d - h % d
    new_w = w if w % d == 0 else w + d - w % d
    lh, lw = (new_h - h) // 2, (new_w - w) // 2
    return (int(lw), int(new_w - w - lw), int(lh), int(new_h - h - lh))


def pad_divide_by(in_img, d):
    h, w = in_img.shape[-2:]
    new_h = h if h % d == 0 else h + d - h % d
    new_w = w if w % d == 0 else w + d - w % d
    lh, lw = (new_h - h) // 2, (new_w - w) // 2
    pad_array = (int(lw), int(new_w - w - lw), int(lh), int(new_h - h - lh))
    if new_h == h and new_w == w:
        return in_img, pad_array
    return F.pad(in_img, pad_array), pad_array


def unpad(img, pad):
    lw, uw, lh, uh = pad
    if img.dim() not in (3, 4):
        raise NotImplementedError
    if lh + uh > 0:
        img = img[..., lh:img.shape[-2] - uh, :]
    if lw + uw > 0:
        img = img[..., lw:img.shape[-1] - uw]
    return img
