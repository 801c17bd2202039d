"""Synthetic VOC-shaped workload of SURVEY.md section 8(d): the tagged, flattened batch the
reference's ``collate(flatten=True)`` (mmseg/datasets/builder.py:295-302) hands to
``forward_train`` -- ``n_sup`` labeled crops, then per unlabeled sample a (student, teacher) pair.
Used by bench.py for both arms (no dataset or network access is needed)."""
import torch


def make_batch(n_sup, n_unsup, size, num_classes, seed=1999, cell=32, border=0.05):
    """img ~ N(0,1) [n,3,size,size] f32; gt [n,1,size,size] i64 piecewise constant on a
    ``cell``-px grid with a ``border`` band of 255 (ignore); metas with unique filenames/tags."""
    gen = torch.Generator().manual_seed(seed)
    n = n_sup + 2 * n_unsup
    img = torch.randn(n, 3, size, size, generator=gen)
    cells = max(size // cell, 1)
    coarse = torch.randint(0, num_classes, (n, 1, cells, cells), generator=gen)
    gt = coarse.repeat_interleave(size // cells, 2).repeat_interleave(size // cells, 3)
    if gt.shape[-1] != size:   # size not a multiple of the cell count: pad with ignore
        full = torch.full((n, 1, size, size), 255, dtype=torch.int64)
        full[:, :, :gt.shape[2], :gt.shape[3]] = gt
        gt = full
    bw = max(int(round(size * border)), 1)
    gt[:, :, :bw] = 255
    gt[:, :, -bw:] = 255
    gt[:, :, :, :bw] = 255
    gt[:, :, :, -bw:] = 255
    tags = ['sup'] * n_sup
    names = [f'sup_{i}.jpg' for i in range(n_sup)]
    for i in range(n_unsup):
        tags += ['unsup_student', 'unsup_teacher']
        names += [f'unsup_{i}.jpg'] * 2
    metas = [dict(filename=f, tag=t, ori_shape=(size, size, 3), img_shape=(size, size, 3),
                  pad_shape=(size, size, 3), scale_factor=1.0, flip=False, flip_direction=None)
             for f, t in zip(names, tags)]
    return img, gt.contiguous(), metas


def fresh_metas(metas):
    """``forward_train`` writes PatchMixIndex / PatchMix_N into the meta dicts in place
    (generate_unsup_data.py:805-812): every iteration needs its own copies, as the dataloader
    provides in the reference."""
    return [dict(m) for m in metas]
