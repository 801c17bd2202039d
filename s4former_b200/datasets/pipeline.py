"""GPU-side input pipeline of the labeled / weak / strong branches.

Mirror of what the reference's dataloader workers do AFTER the geometric part of the train pipeline
(``configs/setr/*_MT_w_ours.py:42-126``: ``Resize`` -> ``RandomCrop`` -> ``RandomFlip`` stay on the
host, they work on decoded images of arbitrary size):

    sup crop       -> PhotoMetricDistortion -> Normalize -> Pad -> DefaultFormatBundle   (tag 'sup')
    unlabeled crop -> MultiBranch: the SAME crop through two such chains with independent
                      distortion draws                       (tags 'unsup_student', 'unsup_teacher')

``mmseg/datasets/pipelines/transforms.py:1165-1272`` (PhotoMetricDistortion), ``:572-604`` (Normalize
= ``mmcv.imnormalize``), ``:484-565`` (Pad), ``formatting.py:202-225`` (DefaultFormatBundle).

The host keeps what must stay on the host -- the RNG: ``draw_pmd_params`` consumes ``numpy.random``
exactly as ``PhotoMetricDistortion.__call__`` does (coin, then the magnitude only if the coin says
so) -- and uploads the uint8 crops (1 byte per pixel and channel; the reference's loader ships 4
bytes per pixel, channel AND branch).  ``s4_branch_pipeline`` then produces the flattened, tagged
float32 batch the segmentor's ``forward_train`` takes, in one launch.  No CPU fallback."""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from ..ops import _p, _st


def draw_pmd_params(brightness_delta=32, contrast_range=(0.5, 1.5), saturation_range=(0.5, 1.5), hue_delta=18,
                    rand_colorjitter_prob=2):
    """One PhotoMetricDistortion's random draws, in the reference's order (transforms.py:1203-1270).
    Returns (do_brightness, beta, mode, do_contrast, alpha_c, do_saturation, alpha_s, do_hue, hue_delta)."""
    random = np.random
    do_b = int(1 - random.randint(rand_colorjitter_prob))
    beta = float(random.uniform(-brightness_delta, brightness_delta)) if do_b else 0.0
    mode = int(random.randint(2))
    do_c = alpha_c = None
    if mode == 1:
        do_c = int(1 - random.randint(rand_colorjitter_prob))
        alpha_c = float(random.uniform(contrast_range[0], contrast_range[1])) if do_c else 1.0
    do_s = int(1 - random.randint(rand_colorjitter_prob))
    alpha_s = float(random.uniform(saturation_range[0], saturation_range[1])) if do_s else 1.0
    do_h = int(1 - random.randint(rand_colorjitter_prob))
    dh = int(random.randint(-hue_delta, hue_delta)) if do_h else 0
    if mode == 0:
        do_c = int(1 - random.randint(rand_colorjitter_prob))
        alpha_c = float(random.uniform(contrast_range[0], contrast_range[1])) if do_c else 1.0
    return (do_b, beta, mode, do_c, alpha_c, do_s, alpha_s, do_h, dh)


_PMD_DTYPE = np.dtype([('do_b', '<i4'), ('beta', '<f4'), ('mode', '<i4'), ('do_c', '<i4'), ('alpha_c', '<f4'),
                       ('do_s', '<i4'), ('alpha_s', '<f4'), ('do_h', '<i4'), ('dh', '<i4')])


class BranchPipeline:
    """``crop_size`` = the Pad target (512, 512); ``mean`` / ``std`` / ``to_rgb`` = ``img_norm_cfg``."""

    def __init__(self, crop_size, mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), to_rgb=True,
                 seg_pad_val=255, device='cuda', pmd_kwargs=None):
        assert L.load().s4_pmd_params_size() == _PMD_DTYPE.itemsize
        self.crop_size = (int(crop_size[0]), int(crop_size[1]))
        self.device = torch.device(device)
        self.to_rgb, self.seg_pad_val = bool(to_rgb), int(seg_pad_val)
        self.pmd_kwargs = dict(pmd_kwargs or {})
        mean32 = np.asarray(mean, dtype=np.float32)
        std32 = np.asarray(std, dtype=np.float32)
        # mmcv.imnormalize_: stdinv = 1 / np.float64(std); cv2 applies it to the float32 image in float32
        self.mean = torch.from_numpy(mean32.copy()).to(self.device)
        self.stdinv = torch.from_numpy(1.0 / std32.astype(np.float64)).to(self.device)          # float64
        self._stage = None

    def draw(self, n_sup, n_unsup):
        """Distortion draws of one batch in the dataloader's order: labeled samples first, then per
        unlabeled sample its student (strong) and teacher (weak) branch (MultiBranch iterates the
        branches in keyword order, ..._MT_w_ours.py:122-125)."""
        return [draw_pmd_params(**self.pmd_kwargs) for _ in range(n_sup + 2 * n_unsup)]

    def __call__(self, sup_crops, sup_labels, unsup_crops, unsup_labels=None, params=None, metas=None,
                 want_u8=False):
        """``*_crops``: lists of uint8 [h, w, 3] BGR arrays / tensors (h, w <= crop_size), ``*_labels``: uint8
        [h, w] (unlabeled ones may be None -> all ``seg_pad_val``).  Returns (img [n,3,H,W] f32, gt [n,1,H,W]
        i64, img_metas) in the flattened order sup.., (student_i, teacher_i)..; with ``want_u8`` also the
        distorted uint8 images [n, H, W, 3]."""
        ns, nu = len(sup_crops), len(unsup_crops)
        nb = ns + 2 * nu
        if params is None:
            params = self.draw(ns, nu)
        assert len(params) == nb
        crops = list(sup_crops) + list(unsup_crops)
        labels = list(sup_labels) + (list(unsup_labels) if unsup_labels is not None else [None] * nu)
        PH, PW = self.crop_size
        dev = self.device
        d_crops, d_labels, hw = [], [], []
        for c, lb in zip(crops, labels):
            c = torch.as_tensor(np.ascontiguousarray(c) if isinstance(c, np.ndarray) else c)
            assert c.dtype == torch.uint8 and c.dim() == 3 and c.shape[2] == 3 and c.shape[0] <= PH and c.shape[1] <= PW
            d_crops.append(c.contiguous().to(dev, non_blocking=True))
            hw += [int(c.shape[0]), int(c.shape[1])]
            if lb is None:
                d_labels.append(None)
            else:
                lb = torch.as_tensor(np.ascontiguousarray(lb) if isinstance(lb, np.ndarray) else lb)
                assert lb.dtype == torch.uint8 and tuple(lb.shape) == tuple(c.shape[:2])
                d_labels.append(lb.contiguous().to(dev, non_blocking=True))
        crop_of = list(range(ns))
        for i in range(nu):
            crop_of += [ns + i, ns + i]
        ptrs = torch.tensor([t.data_ptr() for t in d_crops] + [0 if t is None else t.data_ptr() for t in d_labels],
                            dtype=torch.int64).to(dev, non_blocking=True)
        meta_i = torch.tensor(hw + crop_of, dtype=torch.int32).to(dev, non_blocking=True)
        pm = np.zeros(nb, dtype=_PMD_DTYPE)
        for i, p in enumerate(params):
            pm[i] = tuple(p)
        pm_d = torch.from_numpy(pm.view(np.uint8).copy()).to(dev, non_blocking=True)
        img = torch.empty((nb, 3, PH, PW), dtype=torch.float32, device=dev)
        gt = torch.empty((nb, 1, PH, PW), dtype=torch.int64, device=dev)
        u8 = torch.empty((nb, PH, PW, 3), dtype=torch.uint8, device=dev) if want_u8 else None
        nc = len(d_crops)
        L.call('s4_branch_pipeline', ptrs.data_ptr(), ptrs.data_ptr() + 8 * nc, meta_i.data_ptr(),
               meta_i.data_ptr() + 4 * 2 * nc, _p(pm_d), _p(self.mean), _p(self.stdinv), int(self.to_rgb),
               self.seg_pad_val, _p(img), _p(gt), _p(u8), nb, PH, PW, _st())
        self._keep = (d_crops, d_labels, ptrs, meta_i, pm_d)      # alive until the launch has run
        tags = ['sup'] * ns
        for _ in range(nu):
            tags += ['unsup_student', 'unsup_teacher']
        out_metas = []
        for j, t in enumerate(tags):
            ci = crop_of[j]
            base = dict(metas[ci]) if metas is not None else dict(filename=f'crop_{ci}.jpg')
            h, w = hw[2 * ci], hw[2 * ci + 1]
            base.update(tag=t, img_shape=(h, w, 3), pad_shape=(PH, PW, 3),
                        img_norm_cfg=dict(mean=self.mean.cpu().numpy(), std=(1.0 / self.stdinv.cpu().numpy()).astype(np.float32),
                                          to_rgb=self.to_rgb))
            base.setdefault('ori_shape', (h, w, 3))
            base.setdefault('scale_factor', 1.0)
            base.setdefault('flip', False)
            base.setdefault('flip_direction', None)
            out_metas.append(base)
        return (img, gt, out_metas, u8) if want_u8 else (img, gt, out_metas)
