"""The three shipped SETR-PUP DeiT-Base model dicts, parameterised by crop size / class count.

Mirrors ``configs/setr/setr_deit-base_pup_bs_8_512x512_80k_pascal_1over16_split_classic_*``
of the reference (``_sup.py``, ``_MT.py``, ``_MT_w_ours.py`` -- model section, lines 139-256 of
the ``_MT_w_ours`` file) so benchmarks and tests can build the exact model without mmcv's
``Config.fromfile``.  When mmcv is available the reference's own config files build the same
modules through ``register_into_mmseg()``.
"""
import copy

OPTIMIZER = dict(type='SGD', lr=0.001, momentum=0.9, weight_decay=0.0,
                 paramwise_cfg=dict(custom_keys={'head': dict(lr_mult=10.)}))
LR_CONFIG = dict(policy='poly', power=0.9, min_lr=1e-4, by_epoch=False)   # schedule_80k_pascal_1over8.py:2-5
MAX_ITERS = 80000


def setr_pup_deit_base(variant='ours', img_size=512, num_classes=21, norm='SyncBN', patchmix_n=8,
                       embed_dims=768, num_heads=12, num_layers=12, out_indices=(4, 7, 9, 11),
                       channels=256):
    """variant: 'sup' (beta=0), 'mt' (Mean Teacher + CutMix as shipped), 'ours' (S4Former full)."""
    norm_cfg = dict(type=norm, requires_grad=True)
    backbone = dict(type='VisionTransformer', img_size=(img_size, img_size), patch_size=16, in_channels=3,
                    norm_cfg=dict(type='LN', eps=1e-6, requires_grad=True), with_cls_token=True,
                    interpolate_mode='bilinear', drop_rate=0., embed_dims=embed_dims, num_heads=num_heads,
                    num_layers=num_layers, out_indices=tuple(out_indices))
    decode_head = dict(type='SETRUPHead', align_corners=False, num_convs=4, in_channels=embed_dims,
                       num_classes=num_classes, channels=channels, in_index=3, dropout_ratio=0,
                       norm_cfg=norm_cfg, up_scale=2, kernel_size=3,
                       loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0))
    auxiliary_head = [dict(type='SETRUPHead', in_channels=embed_dims, channels=channels, in_index=i,
                           num_classes=num_classes, dropout_ratio=0, norm_cfg=norm_cfg, num_convs=2,
                           up_scale=4, kernel_size=3, align_corners=False,
                           loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.4))
                      for i in range(4)]
    model = dict(type='EncoderDecoder', pretrained=None, backbone=backbone,
                 backbone_ema=copy.deepcopy(backbone), auxiliary_head=auxiliary_head,
                 decode_head=decode_head, decode_head_ema=copy.deepcopy(decode_head), ema=True,
                 ema_momentum=0.999, unsup_confidence=0.95, test_cfg=dict(mode='whole'))
    if variant == 'sup':
        model.update(unsup_weight=0.0)
    elif variant == 'mt':
        model.update(unsup_weight=1.0, use_CutMix=True)
    elif variant == 'ours':
        model.update(unsup_weight=1.0, attn_mask_seperate_head=True, attn_mask_weight=5,
                     adaptive_attn_mask=True, use_PatchShuffle_w_Cutmix=True, PatchMix_N=patchmix_n,
                     negative_class_ranking=True, negative_class_ranking_mode='unsup_only')
    else:
        raise KeyError(variant)
    return model


def step_flops(variant='ours', n_sup=8, n_unsup=8, img_size=512, num_classes=21, embed_dims=768,
               num_layers=12, channels=256, mt_intended=False):
    """Algorithmic FLOPs (2*MAC of the dense contractions; backward = 2x forward) of one train
    step, SURVEY.md section 8(d).  Returns FLOPs (not GFLOPs)."""
    g = img_size // 16
    L = g * g + 1
    D = embed_dims
    patch = 2.0 * g * g * D * (3 * 16 * 16)
    layer = 2.0 * L * D * 3 * D + 2 * 2.0 * L * L * D + 2.0 * L * D * D + 2 * 2.0 * L * D * 4 * D
    backbone = patch + num_layers * layer
    C = channels

    def conv(hw, cin):
        return 2.0 * hw * hw * C * 9 * cin

    main = conv(g, D) + conv(2 * g, C) + conv(4 * g, C) + conv(8 * g, C) + 2.0 * (16 * g) ** 2 * C * num_classes
    aux = conv(g, D) + conv(4 * g, C) + 2.0 * (16 * g) ** 2 * C * num_classes
    sup = 3 * (backbone + main + 4 * aux)
    teacher = backbone + main
    student = 3 * (backbone + main)
    if variant == 'sup':
        return n_sup * sup
    if variant == 'mt':
        # as shipped: teacher + a loss-less student backbone forward (hazard 4)
        extra = teacher + (student if mt_intended else backbone)
        return n_sup * sup + n_unsup * extra
    return n_sup * sup + n_unsup * (teacher + 2 * student)
