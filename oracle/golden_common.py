"""Seeded tiny configs / inputs shared by oracle/make_golden.py and tests/ (TEST
INFRASTRUCTURE ONLY).  Weights and inputs are regenerated from seeds instead of being
stored; golden files carry checksums of what the reference actually saw."""
import copy

import torch

TINY = dict(img=128, patch=16, dims=128, heads=2, layers=4, out_indices=(0, 1, 2, 3),
            channels=32, classes=5, patchmix_n=2)

GRAD_KEYS = (
    'backbone.cls_token', 'backbone.pos_embed', 'backbone.patch_embed.projection.bias',
    'backbone.layers.0.ln1.weight', 'backbone.layers.0.attn.attn.in_proj_bias',
    'backbone.layers.1.attn.attn.out_proj.weight', 'backbone.layers.3.ffn.layers.0.0.bias',
    'backbone.layers.3.ffn.layers.1.weight', 'decode_head.norm.weight',
    'decode_head.up_convs.0.0.conv.weight', 'decode_head.up_convs.3.0.bn.weight',
    'decode_head.conv_seg.weight', 'decode_head.conv_seg.bias',
    'auxiliary_head.0.up_convs.1.0.conv.weight', 'auxiliary_head.2.conv_seg.bias',
)
EMA_KEYS = (
    'backbone_ema.cls_token', 'backbone_ema.layers.2.ffn.layers.1.weight',
    'decode_head_ema.up_convs.1.0.bn.running_var', 'decode_head_ema.up_convs.1.0.bn.running_mean',
    'decode_head_ema.conv_seg.weight', 'decode_head_ema.up_convs.2.0.bn.num_batches_tracked',
)


def tiny_cfg(variant='ours'):
    """The three shipped configs (configs/setr/*_{sup,MT,MT_w_ours}.py) shrunk to TINY."""
    t = TINY
    norm_cfg = dict(type='SyncBN', requires_grad=True)
    bb = dict(type='VisionTransformer', img_size=(t['img'], t['img']), patch_size=t['patch'],
              in_channels=3, norm_cfg=dict(type='LN', eps=1e-6, requires_grad=True),
              with_cls_token=True, interpolate_mode='bilinear', drop_rate=0.,
              embed_dims=t['dims'], num_heads=t['heads'], num_layers=t['layers'],
              out_indices=t['out_indices'])
    dh = dict(type='SETRUPHead', align_corners=False, num_convs=3, in_channels=t['dims'],
              num_classes=t['classes'], channels=t['channels'], in_index=3, dropout_ratio=0,
              norm_cfg=norm_cfg, up_scale=2, kernel_size=3,
              loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0))
    # 128px / 16 = 8 tokens; 3 x2 up-convs give 64px; the reference config uses 4 (x16).
    dh['num_convs'] = 4
    aux = [dict(type='SETRUPHead', in_channels=t['dims'], channels=t['channels'], in_index=i,
                num_classes=t['classes'], dropout_ratio=0, norm_cfg=norm_cfg, num_convs=2,
                up_scale=4, kernel_size=3, align_corners=False,
                loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=0.4))
           for i in range(4)]
    model = dict(type='EncoderDecoder', pretrained=None, backbone=bb, auxiliary_head=aux,
                 decode_head=dh, test_cfg=dict(mode='whole'))
    if variant == 'sup':      # ..._sup.py: beta = 0, EMA still on
        model.update(backbone_ema=copy.deepcopy(bb), decode_head_ema=copy.deepcopy(dh), ema=True,
                     ema_momentum=0.999, unsup_weight=0.0, unsup_confidence=0.95)
    elif variant == 'mt':     # ..._MT.py: Mean Teacher + CutMix, as shipped
        model.update(backbone_ema=copy.deepcopy(bb), decode_head_ema=copy.deepcopy(dh), ema=True,
                     ema_momentum=0.999, unsup_weight=1.0, unsup_confidence=0.95,
                     use_CutMix=True)
    elif variant == 'ours':   # ..._MT_w_ours.py
        model.update(backbone_ema=copy.deepcopy(bb), decode_head_ema=copy.deepcopy(dh), ema=True,
                     ema_momentum=0.999, unsup_weight=1.0, unsup_confidence=0.95,
                     attn_mask_seperate_head=True, attn_mask_weight=5, adaptive_attn_mask=True,
                     use_PatchShuffle_w_Cutmix=True, PatchMix_N=t['patchmix_n'],
                     negative_class_ranking=True, negative_class_ranking_mode='unsup_only')
    else:
        raise KeyError(variant)
    return model


def seeded_state_dict(template, seed=5, ema_cls_std=6.0):
    """Deterministic weights for every key of ``template`` (shapes/dtypes kept).
    conv_seg of the EMA head is scaled up (``ema_cls_std``) so that a useful fraction of pixels
    clears the 0.95 confidence threshold."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(template.keys()):
        v = template[k]
        if not v.dtype.is_floating_point:
            out[k] = torch.zeros_like(v)
            continue
        if k.endswith('running_var'):
            t = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif k.endswith('running_mean'):
            t = torch.randn(v.shape, generator=g) * 0.1
        elif 'bn.weight' in k or 'ln1.weight' in k or 'ln2.weight' in k or 'norm.weight' in k:
            t = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('bias'):
            t = 0.02 * torch.randn(v.shape, generator=g)
        elif 'conv_seg.weight' in k:
            t = torch.randn(v.shape, generator=g) * (ema_cls_std if 'ema' in k else 0.1)
        elif v.dim() >= 2:
            fan_in = v[0].numel()
            t = torch.randn(v.shape, generator=g) * (1.0 / fan_in ** 0.5)
        else:
            t = 0.02 * torch.randn(v.shape, generator=g)
        out[k] = t.to(v.dtype)
    return out


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def tiny_segformer_cfg(variant='ours'):
    """SegFormer / MiT variant (BASELINE config 5) shrunk to TINY: the
    backbone / head of configs/segformer/segformer_mit-b4_..._CPS_sup.py with the semi-supervised
    settings of the SETR configs ('to be synthesised', SURVEY.md hazard 7).  PatchMix_N = 4 keeps the
    stage-4 un-shuffle block (PatchMix_N / 2 tokens) integral on a 128-pixel crop."""
    t = TINY
    norm_cfg = dict(type='SyncBN', requires_grad=True)
    bb = dict(type='MixVisionTransformer', in_channels=3, embed_dims=16, num_stages=4, num_layers=[1, 2, 1, 2],
              num_heads=[1, 2, 4, 8], patch_sizes=[7, 3, 3, 3], sr_ratios=[8, 4, 2, 1], out_indices=(0, 1, 2, 3),
              mlp_ratio=2, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0)
    dh = dict(type='SegformerHead', in_channels=[16, 32, 64, 128], in_index=[0, 1, 2, 3], channels=32,
              dropout_ratio=0.0, num_classes=t['classes'], norm_cfg=norm_cfg, align_corners=False,
              loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0))
    model = dict(type='EncoderDecoder', pretrained=None, backbone=bb, decode_head=dh, test_cfg=dict(mode='whole'),
                 backbone_ema=copy.deepcopy(bb), decode_head_ema=copy.deepcopy(dh), ema=True, ema_momentum=0.999,
                 unsup_confidence=0.95)
    if variant == 'sup':
        model.update(unsup_weight=0.0)
    elif variant == 'ours':
        model.update(unsup_weight=1.0, attn_mask_seperate_head=True, attn_mask_weight=5, adaptive_attn_mask=True,
                     use_PatchShuffle_w_Cutmix=True, PatchMix_N=4, negative_class_ranking=True,
                     negative_class_ranking_mode='unsup_only')
    else:
        raise KeyError(variant)
    return model


def tiny_batch(variant='ours', seed=1999):
    from oracle.s4former_oracle import synthetic_batch
    n_unsup = 0 if variant == 'sup' else 2
    return synthetic_batch(2, n_unsup, TINY['img'], TINY['classes'], seed=seed, grid=16)


# ----------------------------------------------------------------------------------------
# full-size fixtures (BASELINE shapes): oracle/make_golden_full.py, tests/test_full_parity_gpu.py
# ----------------------------------------------------------------------------------------
FULL = dict(
    full512=dict(size=512, classes=21, n_sup=2, n_unsup=2, wseed=5, ema_cls_std=2.0, seed=1999),
    full768=dict(size=768, classes=19, n_sup=1, n_unsup=1, wseed=6, ema_cls_std=2.0, seed=2024),
)


def full_cfg(shape='full512', variant='ours', norm='SyncBN'):
    """The shipped ``_MT_w_ours`` model dict at a BASELINE shape (DeiT-B, SETR-PUP)."""
    from s4former_b200 import configs
    f = FULL[shape]
    return configs.setr_pup_deit_base(variant, f['size'], f['classes'], norm=norm)


def full_batch(shape='full512'):
    from oracle.s4former_oracle import synthetic_batch
    f = FULL[shape]
    return synthetic_batch(f['n_sup'], f['n_unsup'], f['size'], f['classes'], seed=f['seed'], grid=32)


def strided_sample(t, n):
    """<= n elements of ``t`` at evenly spaced flat positions (the whole tensor if it is smaller)."""
    flat = t.detach().reshape(-1)
    if flat.numel() <= n:
        return flat.clone().cpu()
    idx = (torch.arange(n, dtype=torch.int64) * flat.numel()) // n
    return flat[idx.to(flat.device)].clone().cpu()


def full_grad_keys(names):
    """Parameters whose gradient samples are stored: the tiny-fixture keys mapped to DeiT-B depth
    plus one tensor of every kind in every part of the model."""
    want = [
        'backbone.cls_token', 'backbone.pos_embed', 'backbone.patch_embed.projection.weight',
        'backbone.patch_embed.projection.bias',
        'backbone.layers.0.ln1.weight', 'backbone.layers.0.ln1.bias',
        'backbone.layers.0.attn.attn.in_proj_weight', 'backbone.layers.0.attn.attn.in_proj_bias',
        'backbone.layers.0.attn.attn.out_proj.weight', 'backbone.layers.0.ffn.layers.0.0.weight',
        'backbone.layers.5.attn.attn.in_proj_weight', 'backbone.layers.5.attn.attn.out_proj.bias',
        'backbone.layers.5.ffn.layers.0.0.bias', 'backbone.layers.5.ffn.layers.1.weight',
        'backbone.layers.5.ln2.weight',
        'backbone.layers.11.attn.attn.in_proj_bias', 'backbone.layers.11.attn.attn.out_proj.weight',
        'backbone.layers.11.ffn.layers.0.0.weight', 'backbone.layers.11.ffn.layers.1.bias',
        'decode_head.norm.weight', 'decode_head.norm.bias',
        'decode_head.up_convs.0.0.conv.weight', 'decode_head.up_convs.0.0.bn.weight',
        'decode_head.up_convs.1.0.conv.weight', 'decode_head.up_convs.2.0.bn.bias',
        'decode_head.up_convs.3.0.conv.weight', 'decode_head.up_convs.3.0.bn.weight',
        'decode_head.conv_seg.weight', 'decode_head.conv_seg.bias',
        'auxiliary_head.0.norm.weight', 'auxiliary_head.0.up_convs.1.0.conv.weight',
        'auxiliary_head.1.up_convs.0.0.conv.weight', 'auxiliary_head.2.conv_seg.bias',
        'auxiliary_head.3.up_convs.1.0.bn.weight', 'auxiliary_head.3.conv_seg.weight',
    ]
    return [k for k in want if k in names]
