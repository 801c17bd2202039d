"""Evidence script for DESIGN.md section 7 (TEST INFRASTRUCTURE ONLY; needs /root/reference).

Loads the reference's OWN ``mit.py`` / ``segformer_head.py`` / ``encoder_decoder.py`` (unmodified, under
the mmcv shim of ``oracle/ref_harness``), builds a tiny MiT + SegformerHead ``EncoderDecoder`` with the
semi-supervised settings of the shipped SETR configs (``..._MT.py`` and ``..._MT_w_ours.py``; the tree
holds no semi-supervised SegFormer config, only ``configs/segformer/..._CPS_sup.py``), and runs
``forward_train`` on a synthetic labeled + unlabeled batch on the CPU with the CutMix coin forced to
heads.  Prints, per variant, either the loss keys or the exception the reference raises.

    python oracle/check_segformer_path.py > profiles/r02_reference_segformer_semi_check.txt
"""
import copy
import os
import sys
import traceback
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle.ref_harness import load_reference as LR  # noqa: E402
from oracle import s4former_oracle as O  # noqa: E402


def load_segformer(ns):
    # the few extra mmcv names mit.py / segformer_head.py import (added here, not in the shim the golden
    # fixtures were generated under)
    import types
    import torch.nn as nn
    from oracle.ref_harness import mmcv_shim as shim
    sys.modules['mmcv.cnn'].Conv2d = nn.Conv2d
    drop = types.ModuleType('mmcv.cnn.bricks.drop')
    drop.build_dropout = shim.build_dropout
    sys.modules['mmcv.cnn.bricks.drop'] = drop

    def trunc_normal_init(module, mean=0., std=1., a=-2., b=2., bias=0.):
        if getattr(module, 'weight', None) is not None:
            shim.trunc_normal_(module.weight, mean, std, a, b)
        if getattr(module, 'bias', None) is not None:
            nn.init.constant_(module.bias, bias)
    sys.modules['mmcv.cnn.utils.weight_init'].trunc_normal_init = trunc_normal_init
    # mit.py:150 asks mmseg for the mmcv version (>= 1.3.17 selects the current forward, not the legacy one)
    sys.modules['mmseg'].digit_version = lambda v, length=4: tuple(int(x) for x in v.split('.')[:length])
    sys.modules['mmseg'].mmcv_version = (1, 6, 0)
    mutils = sys.modules['mmseg.models.utils']
    sc = LR._load('mmseg.models.utils.shape_convert', 'mmseg/models/utils/shape_convert.py')
    mutils.nchw_to_nlc, mutils.nlc_to_nchw = sc.nchw_to_nlc, sc.nlc_to_nchw
    mit = LR._load('mmseg.models.backbones.mit', 'mmseg/models/backbones/mit.py')
    sh = LR._load('mmseg.models.decode_heads.segformer_head', 'mmseg/models/decode_heads/segformer_head.py')
    return mit, sh


def cfg(variant, img=128, classes=5):
    norm_cfg = dict(type='BN', requires_grad=True)
    bb = dict(type='MixVisionTransformer', in_channels=3, embed_dims=16, num_stages=4, num_layers=[1, 1, 1, 1],
              num_heads=[1, 2, 4, 8], patch_sizes=[7, 3, 3, 3], sr_ratios=[8, 4, 2, 1], out_indices=(0, 1, 2, 3),
              mlp_ratio=2, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0)
    dh = dict(type='SegformerHead', in_channels=[16, 32, 64, 128], in_index=[0, 1, 2, 3], channels=32,
              dropout_ratio=0.0, num_classes=classes, norm_cfg=norm_cfg, align_corners=False,
              loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0))
    model = dict(type='EncoderDecoder', pretrained=None, backbone=bb, decode_head=dh, test_cfg=dict(mode='whole'),
                 backbone_ema=copy.deepcopy(bb), decode_head_ema=copy.deepcopy(dh), ema=True, ema_momentum=0.999,
                 unsup_weight=1.0, unsup_confidence=0.1)     # max softmax >= 1/C = 0.2: every pixel is pseudo-labelled
    if variant == 'mt':
        model.update(use_CutMix=True)
    elif variant == 'ours':
        model.update(attn_mask_seperate_head=True, attn_mask_weight=5, adaptive_attn_mask=True,
                     use_PatchShuffle_w_Cutmix=True, PatchMix_N=4, negative_class_ranking=True,
                     negative_class_ranking_mode='unsup_only')
    return model


def main():
    ns = LR.load()
    load_segformer(ns)
    print('reference tree:', LR.REF)
    print('semi-supervised SegFormer configs in the tree:',
          [f for f in os.listdir(os.path.join(LR.REF, 'configs', 'segformer'))])
    for variant in ('mt', 'ours'):
        for coin, coin_name in ((0.0, 'CutMix coin = heads (np.random.uniform(0, 1) -> 0.0)'),
                                (0.99, 'CutMix coin = tails (np.random.uniform(0, 1) -> 0.99)')):
            print(f'\n=== variant {variant!r}: {coin_name}')
            torch.manual_seed(0)
            np.random.seed(0)
            try:
                m = ns.builder.build_segmentor(cfg(variant))
                m.train()
                img, gt, metas = O.synthetic_batch(2, 2, 128, 5, seed=3, grid=16)
                # the strong-augmentation coin is np.random.uniform(0, 1) < strong_aug_prob
                # (encoder_decoder.py:604, :634); np.random.rand drives the cut-out box / shuffle draws
                real_rand, real_uniform = np.random.rand, np.random.uniform
                np.random.uniform = lambda lo=0.0, hi=1.0, size=None: (
                    coin if (size is None and lo == 0 and hi == 1) else real_uniform(lo, hi, size))
                try:
                    losses = m.forward_train(img, metas, gt_semantic_seg=gt, iter=10)
                finally:
                    np.random.rand, np.random.uniform = real_rand, real_uniform
                print('ran; loss keys:', sorted(losses.keys()))
                for k, v in sorted(losses.items()):
                    if torch.is_tensor(v):
                        print(f'   {k}: {float(v):.6f}')
            except Exception as e:
                tb = traceback.extract_tb(e.__traceback__)
                where = [f'{os.path.relpath(fr.filename, LR.REF)}:{fr.lineno} {fr.name}' for fr in tb
                         if fr.filename.startswith(LR.REF)]
                print(f'RAISED {type(e).__name__}: {str(e)[:300]}')
                print('   reference frames:', ' <- '.join(reversed(where[-4:])))


if __name__ == '__main__':
    main()
