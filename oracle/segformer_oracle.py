"""CPU restatement of the reference's SegFormer / MiT variant of the hot path (SURVEY.md section
8(f) rank 2, BASELINE config 5).  TEST INFRASTRUCTURE ONLY, like ``s4former_oracle.py``: nothing
under ``s4former_b200/`` imports it.  There is NO CUDA path for this variant yet (DESIGN.md
section 7); this file and its golden fixtures (``oracle/make_golden_segformer.py`` ->
``tests/golden/segformer_*.pt``, generated from the reference's own ``mit.py`` /
``segformer_head.py`` / ``encoder_decoder.py``) are the oracle such a path has to match.

Every class keeps the reference's ``state_dict`` keys, so one seeded state dict loads into the
reference module, this oracle and (later) the product module.

  OracleMiT             mmseg/models/backbones/mit.py:20-89 (MixFFN), :92-196
                        (EfficientMultiheadAttention), :230-322 (TransformerEncoderLayer),
                        :376-495 (MixVisionTransformer incl. the patch-adaptive mask :464-475)
  OracleSegformerHead   mmseg/models/decode_heads/segformer_head.py:111-190 (+ decode_head.py
                        forward_get_logits / losses / _repatchmix_inputs :186-212)

The segmentor logic is ``s4former_oracle.OracleEncoderDecoder`` (it builds these classes for
``type='MixVisionTransformer'`` / ``'SegformerHead'`` and maps the teacher's quarter-resolution
confidence map to patches of 8, encoder_decoder.py:548-551).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .s4former_oracle import OracleSETRUPHead, _MHAParams, token_unshuffle


# ----------------------------------------------------------------------------------------
# patch-adaptive bias of the MiT variant (mit.py:464-475)
# ----------------------------------------------------------------------------------------
def mit_pasa_bias(u, weight, adaptive=True, topk_idx=None):
    """u [B, g, g] (patch unconfidence, NO cls token) -> additive bias [B, L, L], L = g*g.

    bias[b, q, k] = weight * (1 - u[b, k]) for every query row q, except that the rows listed in
    idx[b] are zeroed, where idx[b] = topk(u[b, 1:], floor(0.5 * (L - 1)), largest=False): the
    reference slices off the first PATCH (there is no cls token here) and then uses the indices of
    the slice as row numbers, so "the most confident half" is shifted by one patch (mit.py:470-472).
    Non-adaptive: bias = weight * u[b, k]."""
    b = u.shape[0]
    flat = u.reshape(b, -1).float()
    L = flat.shape[-1]
    if not adaptive:
        return weight * flat.unsqueeze(1).expand(b, L, L)
    if topk_idx is None:
        topk_idx = torch.topk(flat[:, 1:], int(0.5 * (L - 1)), dim=-1, largest=False)[1]
    gate = torch.ones_like(flat)
    gate[torch.arange(b).unsqueeze(1), topk_idx] = 0
    return weight * gate.unsqueeze(2) * (1.0 - flat).unsqueeze(1)


# ----------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------
class _OverlapPatchEmbed(nn.Module):
    """embed.py PatchEmbed as MiT builds it (mit.py:414-420): Conv2d(k, stride, padding=k//2) ->
    flatten -> LayerNorm; returns (tokens [B, h*w, D], (h, w))."""

    def __init__(self, cin, d, k, stride, eps):
        super().__init__()
        self.projection = nn.Conv2d(cin, d, k, stride=stride, padding=k // 2)
        self.norm = nn.LayerNorm(d, eps=eps)

    def forward(self, x):
        x = self.projection(x)
        hw = (x.shape[2], x.shape[3])
        return self.norm(x.flatten(2).transpose(1, 2)), hw


class _EffAttn(nn.Module):
    """mit.py:92-196: queries from every token, keys / values from the tokens after an sr x sr
    strided conv + LayerNorm (sr > 1); the additive mask is applied only when sr == 1 (:183-189)."""

    def __init__(self, d, heads, sr, eps):
        super().__init__()
        self.attn = _MHAParams(d)
        self.heads, self.sr_ratio = heads, sr
        if sr > 1:
            self.sr = nn.Conv2d(d, d, sr, stride=sr)
            self.norm = nn.LayerNorm(d, eps=eps)

    def forward(self, x, hw, identity, bias=None):
        b, l, d = x.shape
        kv = x
        if self.sr_ratio > 1:
            kv = self.sr(x.transpose(1, 2).reshape(b, d, hw[0], hw[1]))
            kv = self.norm(kv.flatten(2).transpose(1, 2))
        w, bi = self.attn.in_proj_weight, self.attn.in_proj_bias
        q = F.linear(x, w[:d], bi[:d])
        k = F.linear(kv, w[d:2 * d], bi[d:2 * d])
        v = F.linear(kv, w[2 * d:], bi[2 * d:])
        h, hd = self.heads, d // self.heads
        lk = kv.shape[1]
        q = q.reshape(b, l, h, hd).transpose(1, 2)
        k = k.reshape(b, lk, h, hd).transpose(1, 2)
        v = v.reshape(b, lk, h, hd).transpose(1, 2)
        s = (q / math.sqrt(hd)) @ k.transpose(-1, -2)
        if bias is not None and self.sr_ratio == 1:
            s = s + bias.unsqueeze(1)
        o = torch.softmax(s, dim=-1) @ v
        o = o.transpose(1, 2).reshape(b, l, d)
        return identity + self.attn.out_proj(o)


class _MixFFN(nn.Module):
    """mit.py:20-89: 1x1 conv -> depth-wise 3x3 conv -> GELU -> 1x1 conv on the (h, w) map, + identity.
    ``layers`` keeps the reference's Sequential indices (0 fc1, 1 pe_conv, 4 fc2)."""

    def __init__(self, d, hidden):
        super().__init__()
        self.layers = nn.Sequential(nn.Conv2d(d, hidden, 1), nn.Conv2d(hidden, hidden, 3, padding=1, groups=hidden),
                                    nn.GELU(), nn.Identity(), nn.Conv2d(hidden, d, 1), nn.Identity())

    def forward(self, x, hw, identity):
        b, l, d = x.shape
        y = self.layers(x.transpose(1, 2).reshape(b, d, hw[0], hw[1]))
        return identity + y.flatten(2).transpose(1, 2)


class _MiTLayer(nn.Module):
    """mit.py:293-313: x = attn(norm1(x)) + x; x = ffn(norm2(x)) + x."""

    def __init__(self, d, heads, hidden, sr, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(d, eps=eps)
        self.attn = _EffAttn(d, heads, sr, eps)
        self.norm2 = nn.LayerNorm(d, eps=eps)
        self.ffn = _MixFFN(d, hidden)

    def forward(self, x, hw, bias=None):
        x = self.attn(self.norm1(x), hw, identity=x, bias=bias)
        return self.ffn(self.norm2(x), hw, identity=x)


class OracleMiT(nn.Module):
    """MixVisionTransformer (mit.py:325-495), dropout / drop-path 0."""

    def __init__(self, in_channels=3, embed_dims=64, num_stages=4, num_layers=(3, 4, 6, 3),
                 num_heads=(1, 2, 4, 8), patch_sizes=(7, 3, 3, 3), strides=(4, 2, 2, 2),
                 sr_ratios=(8, 4, 2, 1), out_indices=(0, 1, 2, 3), mlp_ratio=4, qkv_bias=True,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., eps=1e-6, **_ignored):
        super().__init__()
        assert drop_rate == 0 and attn_drop_rate == 0 and drop_path_rate == 0 and qkv_bias
        self.out_indices = tuple(out_indices)
        self.layers = nn.ModuleList()
        cin = in_channels
        for i in range(num_stages):
            d = embed_dims * num_heads[i]
            pe = _OverlapPatchEmbed(cin, d, patch_sizes[i], strides[i], eps)
            blocks = nn.ModuleList([_MiTLayer(d, num_heads[i], mlp_ratio * d, sr_ratios[i], eps)
                                    for _ in range(num_layers[i])])
            self.layers.append(nn.ModuleList([pe, blocks, nn.LayerNorm(d, eps=eps)]))
            cin = d

    def init_weights(self):
        """mit.py:409-424."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02, a=-2., b=2.)
                nn.init.constant_(m.bias, 0.)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.)
            elif isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                nn.init.normal_(m.weight, 0, math.sqrt(2.0 / fan_out))
                nn.init.constant_(m.bias, 0.)

    def forward(self, x, attn_mask=None, attn_mask_weight=0.0, adaptive_attn_mask=False, topk_idx=None):
        bias = None
        if attn_mask is not None:
            bias = mit_pasa_bias(attn_mask, attn_mask_weight, adaptive_attn_mask, topk_idx)
        outs = []
        for i, (pe, blocks, norm) in enumerate(self.layers):
            x, hw = pe(x)
            for blk in blocks:
                x = blk(x, hw, bias)           # only the sr == 1 stages look at it
            x = norm(x)
            x = x.reshape(x.shape[0], hw[0], hw[1], -1).permute(0, 3, 1, 2).contiguous()
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)


class _ConvBNAct(nn.Module):
    """mmcv ConvModule(1x1, norm, ReLU): conv (no bias) -> bn -> relu."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class OracleSegformerHead(OracleSETRUPHead):
    """segformer_head.py:111-190.  ``forward_get_logits`` / ``losses`` / ``forward_train`` are the
    BaseDecodeHead ones restated in OracleSETRUPHead (the supervised loss upsamples the
    quarter-resolution logits to the label size, decode_head.py:325-330)."""

    def __init__(self, in_channels=(32, 64, 160, 256), in_index=(0, 1, 2, 3), channels=256, num_classes=19,
                 loss_weight=1.0, align_corners=False, ignore_index=255, dropout_ratio=0.0, **_ignored):
        nn.Module.__init__(self)
        assert dropout_ratio == 0
        self.in_index, self.num_classes = tuple(in_index), num_classes
        self.loss_weight, self.ignore_index, self.align_corners = loss_weight, ignore_index, align_corners
        self.convs = nn.ModuleList([_ConvBNAct(c, channels) for c in in_channels])
        self.fusion_conv = _ConvBNAct(channels * len(in_channels), channels)
        self.conv_seg = nn.Conv2d(channels, num_classes, 1)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d) and m is not self.conv_seg:
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
        nn.init.normal_(self.conv_seg.weight, 0, 0.01)
        nn.init.constant_(self.conv_seg.bias, 0)

    def forward(self, inputs, PatchMix_N=0, PatchMixIndex=None):
        xs = [inputs[i] for i in self.in_index]
        outs = []
        for idx, x in enumerate(xs):
            f = self.convs[idx](x)
            if PatchMix_N != 0:
                # the level's tokens are un-shuffled in blocks of PatchMix_N * 4 / 2^idx tokens: the same
                # 16 * PatchMix_N pixel blocks at this level's stride (segformer_head.py:167-170)
                n, c, h, w = f.shape
                t = token_unshuffle(f.reshape(n, c, h * w).permute(0, 2, 1), PatchMixIndex,
                                    int(PatchMix_N * (4 / (2 ** idx))))
                f = t.permute(0, 2, 1).reshape(n, c, h, w)
            outs.append(F.interpolate(f, size=xs[0].shape[2:], mode='bilinear', align_corners=self.align_corners))
        return self.conv_seg(self.fusion_conv(torch.cat(outs, dim=1)))
