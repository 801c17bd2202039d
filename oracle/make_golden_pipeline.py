"""Golden vectors for the GPU-side input pipeline, produced by the UNMODIFIED reference transforms
(``mmseg/datasets/pipelines/{transforms,formatting,compose}.py`` loaded through a small mmcv image
shim: ``mmcv.bgr2hsv / hsv2bgr / imnormalize / impad`` are the thin OpenCV wrappers mmcv defines):

    python -m oracle.make_golden_pipeline

For a seeded uint8 crop (smooth gradients + noise, 40 x 64 (rows a multiple of OpenCV's SIMD width, see tests/test_pipeline_gpu.py), pad target 64 x 64) and
12 numpy seeds: ``MultiBranch(unsup_student=strong, unsup_teacher=weak)`` of the shipped config's branch
pipelines (``PhotoMetricDistortion -> Normalize -> Pad -> DefaultFormatBundle``) -> the two float32
outputs and label maps; plus the reference ``PhotoMetricDistortion`` alone (uint8).  The tests replay the
same seeds through ``s4former_b200.datasets.draw_pmd_params`` (-> identical draws, or every output
differs) and the oracle / the CUDA kernel.  TEST INFRASTRUCTURE ONLY."""
import importlib.util
import os
import sys
import types

import cv2
import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
REF = os.environ.get('S4_REFERENCE_ROOT', '/root/reference')
NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
PAD = (64, 64)


def seeded_crop(seed=3, h=40, w=64):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 2 + yy) % 256, (yy * 3) % 256, (xx + 2 * yy) % 256], -1)
    img = np.clip(base + rng.randint(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)
    lab = rng.randint(0, 21, (h // 8, w // 8)).repeat(8, 0).repeat(8, 1).astype(np.uint8)
    return img, lab


def load_reference_pipelines():
    class DataContainer:
        def __init__(self, data, stack=False, **kw):
            self._data, self.stack = data, stack

        @property
        def data(self):
            return self._data

    class Registry:
        def __init__(self, name):
            self.name, self.module_dict = name, {}

        def register_module(self, name=None, force=False, module=None):
            def deco(cls):
                self.module_dict[name or cls.__name__] = cls
                return cls
            return deco(module) if module is not None else deco

        def get(self, k):
            return self.module_dict.get(k)

    def imnormalize(img, mean, std, to_rgb=True):
        img = img.copy().astype(np.float32)
        mean = np.float64(mean.reshape(1, -1))
        stdinv = 1 / np.float64(std.reshape(1, -1))
        if to_rgb:
            cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
        cv2.subtract(img, mean, img)
        cv2.multiply(img, stdinv, img)
        return img

    def impad(img, *, shape=None, padding=None, pad_val=0, padding_mode='constant'):
        width = max(shape[1] - img.shape[1], 0)
        height = max(shape[0] - img.shape[0], 0)
        return cv2.copyMakeBorder(img, 0, height, 0, width, cv2.BORDER_CONSTANT, value=pad_val)

    def build_from_cfg(cfg, registry, default_args=None):
        cfg = dict(cfg)
        return registry.get(cfg.pop('type'))(**cfg)

    mmcv = types.ModuleType('mmcv')
    mmcv.bgr2hsv = lambda img: cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    mmcv.hsv2bgr = lambda img: cv2.cvtColor(img, cv2.COLOR_HSV2BGR)
    mmcv.imnormalize, mmcv.impad = imnormalize, impad
    mmcv.is_str = lambda x: isinstance(x, str)
    utils = types.ModuleType('mmcv.utils')
    utils.deprecated_api_warning = lambda *a, **k: (lambda f: f)
    utils.is_tuple_of = lambda seq, t: isinstance(seq, tuple) and all(isinstance(x, t) for x in seq)
    utils.build_from_cfg = build_from_cfg
    utils.Registry = Registry
    parallel = types.ModuleType('mmcv.parallel')
    parallel.DataContainer = DataContainer
    saved = {k: sys.modules.get(k) for k in ('mmcv', 'mmcv.utils', 'mmcv.parallel')}
    sys.modules.update({'mmcv': mmcv, 'mmcv.utils': utils, 'mmcv.parallel': parallel})
    for name in ('mmseg_pl', 'mmseg_pl.datasets', 'mmseg_pl.datasets.pipelines'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    builder = types.ModuleType('mmseg_pl.datasets.builder')
    builder.PIPELINES = Registry('pipeline')
    sys.modules['mmseg_pl.datasets.builder'] = builder
    mods = {}
    for f in ('transforms', 'formatting', 'compose'):
        spec = importlib.util.spec_from_file_location(f'mmseg_pl.datasets.pipelines.{f}',
                                                      os.path.join(REF, f'mmseg/datasets/pipelines/{f}.py'))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[f] = m
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
    return mods


def main():
    mods = load_reference_pipelines()
    T, Fm, Cp = mods['transforms'], mods['formatting'], mods['compose']
    branch = [dict(type='PhotoMetricDistortion'), dict(type='Normalize', **NORM),
              dict(type='Pad', size=PAD, pad_val=0, seg_pad_val=255), dict(type='DefaultFormatBundle')]
    mb = Cp.MultiBranch(unsup_student=list(branch), unsup_teacher=list(branch))
    img, lab = seeded_crop()
    cases = []
    for seed in range(12):
        np.random.seed(seed)
        res = dict(img=img.copy(), gt_semantic_seg=lab.copy(), seg_fields=['gt_semantic_seg'],
                   img_shape=img.shape, ori_shape=img.shape)
        out = mb(res)
        np.random.seed(seed)
        pmd_only = T.PhotoMetricDistortion()(dict(img=img.copy()))['img']
        cases.append(dict(seed=seed, student=out[0]['img'].data.clone(), teacher=out[1]['img'].data.clone(),
                          gt=out[0]['gt_semantic_seg'].data.clone().to(torch.uint8),
                          pmd_u8=torch.from_numpy(pmd_only.copy())))
    path = os.path.join(OUT, 'pipeline.pt')
    torch.save(dict(crop_seed=3, pad=PAD, norm=NORM, cases=cases), path)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    sys.exit(main())
