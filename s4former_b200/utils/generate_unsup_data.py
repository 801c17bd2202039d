"""Batch-level strong augmentation -- mirror of ``mmseg/utils/generate_unsup_data.py``
(``generate_cutout_mask`` :7-26, ``generate_unsup_cutmix_data`` :400-453,
``generate_unsup_patchmix_data`` :737-819).

The HOST RNG call order is part of the contract (numpy ``randint`` x3 per image for CutMix,
``np.random.rand`` then ``torch.randperm`` per image for PatchShuffle); the pixel movement runs in
two gather kernels (``s4_cutmix``, ``s4_patchshuffle``) instead of per-image Python slice loops.
"""
import numpy as np
import torch

from .. import ops


def generate_cutout_box(img_size, ratio=2):
    """Same draws as ``generate_cutout_mask`` but returns the zero box (y0, y1, x0, x1)."""
    if isinstance(ratio, int):
        cutout_area = img_size[0] * img_size[1] / ratio
    else:
        raise NotImplementedError('tuple ratios are not used by the shipped configs')
    w = np.random.randint(img_size[1] / ratio + 1, img_size[1])
    h = np.round(cutout_area / w)
    x_start = np.random.randint(0, img_size[1] - w + 1)
    y_start = np.random.randint(0, img_size[0] - h + 1)
    return int(y_start), int(y_start + h), int(x_start), int(x_start + w)


def generate_unsup_cutmix_data(teacher_info, student_info, ratio=2, patchwise=False, patchsize=16 * 8,
                               boxes=None):
    """``boxes``: pre-drawn zero boxes ([B, 4] int32 (y0, y1, x0, x1), host list or resident device
    tensor; an all-zero row leaves that image untouched) -- the train step draws them up front, in
    the reference's RNG order, so that the device program does not depend on the host
    (``EncoderDecoder.draw_aug_params``).  None: drawn here, as the reference does."""
    if patchwise:
        raise NotImplementedError('patchwise CutMix is not used by the shipped configs')
    data = student_info['img']
    target = teacher_info['hard_seg_label']
    if tuple(target.shape[-2:]) != tuple(data.shape[-2:]):
        raise NotImplementedError('label/image size mismatch: nearest resize is off the hot path')
    b, _, im_h, im_w = data.shape
    if boxes is None:
        boxes = [generate_cutout_box([im_h, im_w], ratio=ratio) for _ in range(b)]
    new_data, new_target = ops.cutmix(data, target, boxes)
    student_info['img'] = new_data
    teacher_info['hard_seg_label'] = new_target
    return teacher_info, student_info


def draw_patchmix_perms(b, h, w, patchmix_ratio=0.5, patch_size=16, PatchMix_N=1):
    """The host RNG part of ``generate_unsup_patchmix_data`` (:737-819): per image one
    ``np.random.rand()`` and, below the ratio, one ``torch.randperm``.  Returns [b, num] int64."""
    size = patch_size * PatchMix_N
    assert h % size == 0 and w % size == 0
    num = (h // size) * (w // size)
    perms = []
    for i in range(b):
        if np.random.rand() < patchmix_ratio:
            perm = torch.arange(num)[torch.randperm(num)]
        else:
            perm = torch.arange(num)
        perms.append(perm)
    return torch.stack(perms)


def generate_unsup_patchmix_data(results, teacher_info=None, patchmix_ratio=0.5, patch_size=16,
                                 PatchMix_N=1, use_mask=False, ratio=2, perms=None, perms_dev=None):
    """``perms`` / ``perms_dev``: pre-drawn permutations (host [b, num] int64 and its resident
    device copy), see ``generate_unsup_cutmix_data``."""
    if use_mask:
        raise NotImplementedError('use_mask PatchMix is not used by the shipped configs')
    data = results['img']
    b, c, h, w = data.shape
    size = patch_size * PatchMix_N
    assert h % size == 0 and w % size == 0
    if perms is None:
        perms = draw_patchmix_perms(b, h, w, patchmix_ratio, patch_size, PatchMix_N)
    for i in range(b):
        results['img_metas'][i]['PatchMixIndex'] = perms[i]
        results['img_metas'][i]['PatchMix_N'] = PatchMix_N
    if perms_dev is not None:      # resident copy for the head's un-shuffle row map (no host round trip)
        results['img_metas'][0]['_s4_perms_dev'] = perms_dev
    results['img'] = ops.patchshuffle(data, perms_dev if perms_dev is not None else perms, size)
    if teacher_info is not None:
        return results, teacher_info
    return results
