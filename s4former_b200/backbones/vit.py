"""DeiT/ViT backbone with S4Former's patch-adaptive self-attention -- B200-native mirror of
``mmseg/models/backbones/vit.py`` (reference file:line cited per method).

Same constructor kwargs, ``forward`` signature, ``init_weights`` and ``state_dict`` keys
(``patch_embed.projection.*``, ``cls_token``, ``pos_embed``, ``layers.{i}.ln1/ln2.*``,
``layers.{i}.attn.attn.in_proj_{weight,bias}``, ``...attn.attn.out_proj.*``,
``layers.{i}.ffn.layers.0.0.*``, ``layers.{i}.ffn.layers.1.*``), but the arithmetic runs in
the CUDA library: tokens stay a flat ``[B*L, D]`` matrix for the whole encoder and each layer
is one autograd node (``ops.EncoderLayerFn``).
"""
import math

import torch
import torch.nn as nn

from .. import ops
from ..builder import BACKBONES


class _MHAParams(nn.Module):
    """Parameter holder named like torch ``nn.MultiheadAttention`` (mmcv wraps one as ``.attn``)."""

    def __init__(self, embed_dims, bias=True):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dims, embed_dims))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dims)) if bias else None
        self.out_proj = nn.Linear(embed_dims, embed_dims)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.)


class _MultiheadAttention(nn.Module):
    def __init__(self, embed_dims, num_heads, bias=True):
        super().__init__()
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.attn = _MHAParams(embed_dims, bias)


class _FFN(nn.Module):
    def __init__(self, embed_dims, feedforward_channels):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.GELU(), nn.Dropout(0.)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(0.))


class TransformerEncoderLayer(nn.Module):
    """vit.py:24-127.  ``forward(x2d, B, L, u0, gate, w)`` on the flat token matrix."""

    def __init__(self, embed_dims, num_heads, feedforward_channels, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., num_fcs=2, qkv_bias=True, act_cfg=dict(type='GELU'),
                 norm_cfg=dict(type='LN'), batch_first=True, attn_cfg=dict(), ffn_cfg=dict(),
                 with_cp=False):
        super().__init__()
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError('dropout / drop-path are 0 in every S4Former config')
        if num_fcs != 2 or act_cfg.get('type') != 'GELU' or norm_cfg.get('type') != 'LN':
            raise NotImplementedError('only the 2-layer GELU FFN with LayerNorm is on the hot path')
        if with_cp:
            raise NotImplementedError('with_cp is unusable in the reference as well (vit.py:123-124)')
        eps = norm_cfg.get('eps', 1e-5)
        self.ln1 = nn.LayerNorm(embed_dims, eps=eps)
        self.attn = _MultiheadAttention(embed_dims, num_heads, qkv_bias)
        self.ln2 = nn.LayerNorm(embed_dims, eps=eps)
        self.ffn = _FFN(embed_dims, feedforward_channels)
        self.num_heads = num_heads

    def forward(self, x2d, B, L, u0=None, gate=None, w=0.0):
        return ops.EncoderLayerFn.apply(x2d, self, B, L, u0, gate, float(w), self.ln1.weight)


class _PatchEmbed(nn.Module):
    def __init__(self, in_channels, embed_dims, patch_size):
        super().__init__()
        self.projection = nn.Conv2d(in_channels, embed_dims, patch_size, patch_size)


@BACKBONES.register_module()
class VisionTransformer(nn.Module):
    """vit.py:187-577."""

    def __init__(self, img_size=224, patch_size=16, in_channels=3, embed_dims=768, num_layers=12,
                 num_heads=12, mlp_ratio=4, out_indices=-1, qkv_bias=True, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., with_cls_token=True, output_cls_token=False,
                 norm_cfg=dict(type='LN'), act_cfg=dict(type='GELU'), patch_norm=False,
                 final_norm=False, interpolate_mode='bicubic', num_fcs=2, norm_eval=False,
                 with_cp=False, pretrained=None, init_cfg=None, no_pos_embed=False,
                 feature_ps_indices=0, w_PatchRelativeAttention=False):
        super().__init__()
        if isinstance(img_size, int):
            img_size = (img_size, img_size)
        elif isinstance(img_size, tuple):
            if len(img_size) == 1:
                img_size = (img_size[0], img_size[0])
            assert len(img_size) == 2, \
                f'The size of image should have length 1 or 2, but got {len(img_size)}'
        if output_cls_token:
            assert with_cls_token is True, f'with_cls_token must be True if' \
                f'set output_cls_token to True, but got {with_cls_token}'
        assert not (init_cfg and pretrained), 'init_cfg and pretrained cannot be set at the same time'
        if isinstance(pretrained, str):
            init_cfg = dict(type='Pretrained', checkpoint=pretrained)
        elif pretrained is not None:
            raise TypeError('pretrained must be a str or None')
        for flag, name in ((not with_cls_token, 'with_cls_token=False'), (output_cls_token, 'output_cls_token'),
                           (patch_norm, 'patch_norm'), (final_norm, 'final_norm'), (no_pos_embed, 'no_pos_embed'),
                           (w_PatchRelativeAttention, 'w_PatchRelativeAttention')):
            if flag:
                raise NotImplementedError(f'{name} is not used by any S4Former config (out of the hot path)')
        if drop_path_rate:
            raise NotImplementedError('drop_path_rate != 0 (stochastic depth) is not used by any S4Former '
                                      'config; refusing to train silently without it')
        self.init_cfg = init_cfg
        self.img_size, self.patch_size = tuple(img_size), patch_size
        self.interpolate_mode, self.norm_eval, self.with_cp = interpolate_mode, norm_eval, with_cp
        self.pretrained, self.num_heads, self.embed_dims = pretrained, num_heads, embed_dims
        self.with_cls_token, self.output_cls_token = with_cls_token, output_cls_token
        self.no_pos_embed, self.feature_ps_indices = no_pos_embed, feature_ps_indices
        self.w_PatchRelativeAttention, self.final_norm = w_PatchRelativeAttention, final_norm
        self.patch_embed = _PatchEmbed(in_channels, embed_dims, patch_size)
        num_patches = (img_size[0] // patch_size) * (img_size[1] // patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dims))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dims))
        if isinstance(out_indices, int):
            if out_indices == -1:
                out_indices = num_layers - 1
            self.out_indices = [out_indices]
        elif isinstance(out_indices, (list, tuple)):
            self.out_indices = out_indices
        else:
            raise TypeError('out_indices must be type of int, list or tuple')
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(embed_dims=embed_dims, num_heads=num_heads,
                                    feedforward_channels=mlp_ratio * embed_dims,
                                    attn_drop_rate=attn_drop_rate, drop_rate=drop_rate,
                                    drop_path_rate=0., num_fcs=num_fcs, qkv_bias=qkv_bias,
                                    act_cfg=act_cfg, norm_cfg=norm_cfg, with_cp=with_cp, batch_first=True)
            for _ in range(num_layers)])
        self.multi_self_attn = [[], None]

    def init_weights(self):
        """vit.py:369-414."""
        if isinstance(self.init_cfg, dict) and self.init_cfg.get('type') == 'Pretrained':
            checkpoint = torch.load(self.init_cfg['checkpoint'], map_location='cpu')
            state_dict = checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint
            state_dict = dict(state_dict)
            if 'pos_embed' in state_dict and self.pos_embed.shape != state_dict['pos_embed'].shape:
                # vit.py:381-392: e.g. the DeiT-B/16 checkpoint (14 x 14 grid, 197 tokens) -> 32 x 32
                h, w = self.img_size
                pos_size = int(math.sqrt(state_dict['pos_embed'].shape[1] - 1))
                state_dict['pos_embed'] = self.resize_pos_embed(
                    state_dict['pos_embed'], (h // self.patch_size, w // self.patch_size),
                    (pos_size, pos_size), self.interpolate_mode, self.no_pos_embed)
            self.load_state_dict(state_dict, strict=False)
            return
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        for n, m in self.named_modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    if 'ffn' in n:
                        nn.init.normal_(m.bias, mean=0., std=1e-6)
                    else:
                        nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0.)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.)

    @staticmethod
    def resize_pos_embed(pos_embed, input_shpae, pos_shape, mode, no_pos_embed=False):
        """vit.py:447-477: cls row kept, the grid rows resized ([1, L, C] -> [1, 1 + h*w, C]).
        CPU tensors (checkpoint load) go through ``F.interpolate``; CUDA tensors through the
        library's bilinear resize."""
        assert pos_embed.ndim == 3, 'shape of pos_embed must be [B, L, C]'
        pos_h, pos_w = pos_shape
        cls_token_weight = pos_embed[:, 0:1]
        pos_embed_weight = pos_embed[:, (-1 * pos_h * pos_w):]
        pos_embed_weight = pos_embed_weight.reshape(1, pos_h, pos_w, pos_embed.shape[2]).permute(0, 3, 1, 2)
        if pos_embed_weight.is_cuda:
            if mode != 'bilinear':
                raise NotImplementedError(f'on-device pos_embed resize: bilinear only (got {mode})')
            pos_embed_weight = ops.resize_bilinear(pos_embed_weight.float().contiguous(), input_shpae)
        else:
            import torch.nn.functional as F
            pos_embed_weight = F.interpolate(pos_embed_weight, size=tuple(input_shpae), mode=mode,
                                             align_corners=False)
        pos_embed_weight = torch.flatten(pos_embed_weight, 2).transpose(1, 2)
        if no_pos_embed:
            pos_embed_weight = torch.zeros_like(pos_embed_weight)
        return torch.cat((cls_token_weight, pos_embed_weight.to(cls_token_weight.dtype)), dim=1)

    def _resized_pos_embed(self, gh, gw):
        pos_h, pos_w = self.img_size[0] // self.patch_size, self.img_size[1] // self.patch_size
        if self.pos_embed.shape[1] != pos_h * pos_w + 1:
            raise ValueError('Unexpected shape of pos_embed, got {}.'.format(self.pos_embed.shape))
        cache = self.__dict__.setdefault('_s4_pos_cache', {})
        tag = (gh, gw, self.pos_embed._version, getattr(self.pos_embed, '_s4_gen', 0), ops._global_gen[0],
               self.pos_embed.data_ptr())
        hit = cache.get('pos')
        if hit is None or hit[0] != tag:
            with torch.no_grad():
                val = self.resize_pos_embed(self.pos_embed.detach(), (gh, gw), (pos_h, pos_w),
                                            self.interpolate_mode).contiguous()
            cache['pos'] = hit = (tag, val)
        return hit[1]

    @staticmethod
    def pasa_bias_vectors(attn_mask, adaptive_attn_mask, topk_idx=None):
        """vit.py:519-535 in rank-1 form: returns (u0 [B,L], gate [B,L]) float32.

        bias[b,h,q,k] = w * gate[b,q] * u0[b,k];  gate is 0 for the floor(0.5*g*g) most confident
        patches (topk smallest unconfidence, +1 for the cls column), 1 elsewhere incl. the cls row.
        ``topk_idx`` lets parity tests inject the index set (tie order is device specific)."""
        b = attn_mask.size(0)
        flat = attn_mask.reshape(b, -1).float()
        u0 = torch.cat((torch.zeros(b, 1, device=flat.device), flat), -1).contiguous()
        gate = None
        if adaptive_attn_mask:
            if callable(topk_idx):      # parity tests: e.g. the reference's CPU torch.topk on this u
                topk_idx = topk_idx(flat)
            if topk_idx is None:
                topk_idx = torch.topk(flat, int(0.5 * flat.size(-1)), dim=-1, largest=False)[1]
            gate = torch.ones_like(u0)
            gate.scatter_(1, topk_idx.to(flat.device) + 1, 0.0)
        return u0, gate

    def forward(self, inputs, no_pos_embed=False, avg_pos_emd=False, duplicate_pos_emd=False,
                use_fdrop=False, attn_mask=None, attn_mask_weight=0.0, adaptive_attn_mask=False,
                topk_idx=None, attn_mask_rows=None):
        """vit.py:479-570.  Returns NCHW-shaped feature maps (channels-last views of the token
        matrix, no copy); each carries ``_s4_tokens = (x2d, B, L)`` for the fused head.

        ``attn_mask_rows=(b0, b1)``: ``attn_mask`` covers only images ``[b0, b1)`` of the batch
        (the segmentor runs its student passes as ONE batch; the other images get a zero bias
        row, which the attention kernels treat as "no bias")."""
        if no_pos_embed or avg_pos_emd or duplicate_pos_emd or use_fdrop:
            raise NotImplementedError('pos-embed ablations / fdrop are off in every shipped config')
        B = inputs.shape[0]
        P = self.patch_size
        gh, gw = math.ceil(inputs.shape[2] / P), math.ceil(inputs.shape[3] / P)
        L = gh * gw + 1
        pos_override = None
        if L != self.pos_embed.shape[1]:
            # vit.py:416-445 _pos_embeding: an input whose patch grid differs from the training grid
            # (whole-image / sliding-window inference) gets a bilinearly resized position embedding
            if torch.is_grad_enabled() and self.pos_embed.requires_grad:
                raise NotImplementedError('training through a resized pos_embed is not on the S4Former path '
                                          '(every shipped config trains at img_size); inference only')
            pos_override = self._resized_pos_embed(gh, gw)
        x = ops.PatchEmbedFn.apply(self.cls_token, self, inputs, pos_override)
        u0 = gate = None
        if attn_mask is not None:
            u0, gate = self.pasa_bias_vectors(attn_mask, adaptive_attn_mask, topk_idx)
            if attn_mask_rows is not None:
                b0, b1 = attn_mask_rows
                assert b1 - b0 == u0.shape[0] and 0 <= b0 and b1 <= B
                u0_all = torch.zeros((B, L), dtype=u0.dtype, device=u0.device)
                u0_all[b0:b1] = u0
                u0 = u0_all
                if gate is not None:
                    gate_all = torch.ones((B, L), dtype=gate.dtype, device=gate.device)
                    gate_all[b0:b1] = gate
                    gate = gate_all
        outs = []
        for i, layer in enumerate(self.layers):
            x = layer(x, B, L, u0, gate, attn_mask_weight if u0 is not None else 0.0)
            if i in self.out_indices:
                out = x.view(B, L, -1)[:, 1:].unflatten(1, (gh, gw)).permute(0, 3, 1, 2)
                out._s4_tokens = (x, B, L)
                out._s4_hw = (gh, gw)
                outs.append(out)
        self.multi_self_attn = [[], (gh, gw)]   # head-averaged weights are visualisation-only
        return tuple(outs)

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.LayerNorm):
                    m.eval()
        return self
