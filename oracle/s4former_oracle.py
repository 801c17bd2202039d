"""CPU oracle for the S4Former semi-supervised train step.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file,
and only as the checker / the timed CPU baseline.  ``s4former_b200`` never imports it.

It is a plain-PyTorch (fp32, CPU) restatement of the reference's algorithm for the path
SURVEY.md section 8(a) lists, each function citing the reference file:line it follows
(paths relative to the reference root).  The mmcv bricks (absent from the reference
tree: mmcv-full >=1.4.4,<=1.6.0) are restated from their published definitions:
``MultiheadAttention`` = ``identity + out_proj(softmax(q k^T / sqrt(d) + mask) v)`` with
packed ``in_proj`` (torch ``nn.MultiheadAttention``), ``FFN`` = ``identity +
W2 gelu_erf(W1 x + b1) + b2``, ``ConvModule`` = conv(no bias) -> BN -> ReLU.

PINNING.  The reference ships no golden vectors for this path (its only known answers
are ``tests/test_models/test_losses/test_ce_loss.py:25-39,43-86`` -- both reproduced in
``tests/test_oracle.py``).  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF run in the build container: ``oracle/make_golden.py`` imports the
reference's own ``vit.py / setr_up_head.py / decode_head.py / cross_entropy_loss.py /
encoder_decoder.py / generate_unsup_data.py`` unmodified (through
``oracle/ref_harness``), runs them on seeded inputs and stores inputs + outputs in
``tests/golden/*.pt``; ``tests/test_oracle.py`` replays them through this file.

Modules expose the same ``state_dict`` keys as the reference so weights move freely
between reference, oracle and the CUDA implementation.
"""
import math
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# losses / pseudo labels  (per-pixel statements, Appendix A of SURVEY.md)
# ----------------------------------------------------------------------------------------
def cross_entropy_mean_all(logits, label, ignore_index=255, loss_weight=1.0):
    """mmseg CrossEntropyLoss with avg_non_ignore=False.

    cross_entropy_loss.py:45-61 (F.cross_entropy(reduction='none', ignore_index)) and
    losses/utils.py:65-69 (``loss.mean()`` over ALL N*H*W pixels, ignored ones count as 0),
    times ``loss_weight`` (cross_entropy_loss.py:268).
    """
    logp = torch.log_softmax(logits.float(), dim=1)
    valid = label != ignore_index
    safe = torch.where(valid, label, torch.zeros_like(label))
    nll = -logp.gather(1, safe.unsqueeze(1)).squeeze(1)
    nll = torch.where(valid, nll, torch.zeros_like(nll))
    return loss_weight * nll.sum() / label.numel()


def pseudo_label(logits_t, threshold=0.95):
    """encoder_decoder.py:890,899-901 and :541-542.

    p = softmax(z_t); (max_value, hard) = max_c p; conf = int64(max_value > thr);
    hard[conf == 0] = 255.  Returns (hard_with_ignore, conf, max_value).
    """
    p = torch.softmax(logits_t.float(), dim=1)
    max_value, hard = torch.max(p, dim=1)
    conf = (max_value > threshold) * 1
    hard = hard.clone()
    hard[conf == 0] = 255
    return hard, conf, max_value


def patch_unconfidence(conf, patch=16):
    """encoder_decoder.py:547-555: u[b,i,j] = mean over the 16x16 block of (1 - conf)."""
    b, h, w = conf.shape
    x = (1 - conf).view(b, h // patch, patch, h // patch, patch).permute(0, 1, 3, 2, 4)
    x = x.reshape(b, h // patch, h // patch, -1)
    return torch.sum(x, -1) / (patch * patch)


def ncr_unsup_only(logits_s, logits_t, hard):
    """Negative-class ranking, mode 'unsup_only' (encoder_decoder.py:936-954).

    For every pixel whose label is c (0..C-1; 255 never matches): softmax over the C-1
    logits != c for student and teacher, then torch PairwiseDistance(p=2, eps=1e-6)
    = || p_s - p_t + 1e-6 ||_2, summed; divided by B*H*W (all pixels).
    Restated per pixel instead of per class (sum order differs only in fp rounding).
    """
    b, c, h, w = logits_s.shape
    zs = logits_s.permute(0, 2, 3, 1).reshape(-1, c).float()
    zt = logits_t.permute(0, 2, 3, 1).reshape(-1, c).float()
    y = hard.reshape(-1)
    valid = (y >= 0) & (y < c)
    ysafe = torch.where(valid, y, torch.zeros_like(y))
    drop = F.one_hot(ysafe, c).bool()
    neg_inf = torch.finfo(torch.float32).min
    ps = torch.softmax(zs.masked_fill(drop, neg_inf), dim=1)
    pt = torch.softmax(zt.masked_fill(drop, neg_inf), dim=1)
    d = (ps - pt + 1e-6).masked_fill(drop, 0.0)
    r = torch.sqrt((d * d).sum(1))
    r = torch.where(valid, r, torch.zeros_like(r))
    return r.sum() / (b * h * w)


# ----------------------------------------------------------------------------------------
# augmentation (host RNG order is part of the contract: Appendix B-7)
# ----------------------------------------------------------------------------------------
def cutout_box(img_size, ratio=2):
    """generate_unsup_data.py:7-26 -- returns (y0, y1, x0, x1) of the zero box.

    RNG calls, in order: np.random.randint(W/ratio+1, W); randint(0, W-w+1);
    randint(0, H-h+1).
    """
    cutout_area = img_size[0] * img_size[1] / ratio
    w = np.random.randint(img_size[1] / ratio + 1, img_size[1])
    h = np.round(cutout_area / w)
    x_start = np.random.randint(0, img_size[1] - w + 1)
    y_start = np.random.randint(0, img_size[0] - h + 1)
    return int(y_start), int(y_start + h), int(x_start), int(x_start + w)


def cutmix(img, hard, boxes):
    """generate_unsup_data.py:400-453: img_i*M_i + img_{(i+1)%B}*(1-M_i), same on labels.
    Labels below the image resolution (SegFormer's quarter-resolution teacher) are nearest-resized
    up to the image, mixed there and nearest-resized back (:407-408, :449-450): label pixel (oy, ox)
    takes the neighbour's value iff image pixel (s*oy, s*ox) lies in the box."""
    b = img.shape[0]
    s = img.shape[-1] // hard.shape[-1]
    out_img, out_lab = img.clone(), hard.clone()
    for i in range(b):
        y0, y1, x0, x1 = boxes[i]
        j = (i + 1) % b
        out_img[i, :, y0:y1, x0:x1] = img[j, :, y0:y1, x0:x1]
        ly0, ly1, lx0, lx1 = -(-y0 // s), -(-y1 // s), -(-x0 // s), -(-x1 // s)
        out_lab[i, ly0:ly1, lx0:lx1] = hard[j, ly0:ly1, lx0:lx1]
    return out_img, out_lab


def draw_patchshuffle_perms(batch, nblocks, patchmix_ratio=0.5):
    """generate_unsup_data.py:737-819 RNG order: per image np.random.rand(), then
    torch.randperm(nblocks) only if the draw is < ratio (identity otherwise)."""
    perms = []
    for _ in range(batch):
        if np.random.rand() < patchmix_ratio:
            perms.append(torch.randperm(nblocks))
        else:
            perms.append(torch.arange(nblocks))
    return torch.stack(perms)


def patchshuffle(img, perms, block):
    """generate_unsup_data.py:786-802: out_block[p] = in_block[perm[p]], blocks row-major."""
    b, c, h, w = img.shape
    gw = w // block
    out = img.clone()
    for i in range(b):
        for p in range(perms.shape[1]):
            s = int(perms[i, p])
            py, px = divmod(p, gw)
            sy, sx = divmod(s, gw)
            out[i, :, py * block:(py + 1) * block, px * block:(px + 1) * block] = \
                img[i, :, sy * block:(sy + 1) * block, sx * block:(sx + 1) * block]
    return out


def token_unshuffle(tokens, perms, n):
    """decode_head.py:186-212: tokens [B, g*g, D] as (g/n)^2 blocks of n x n tokens;
    out_block[perm[p]] = in_block[p]."""
    b, l, d = tokens.shape
    g = int(math.isqrt(l))
    gb = g // n
    x = tokens.reshape(b, gb, n, gb, n, d)
    out = torch.empty_like(x)
    for i in range(b):
        for p in range(gb * gb):
            q = int(perms[i, p])
            py, px = divmod(p, gb)
            qy, qx = divmod(q, gb)
            out[i, qy, :, qx, :, :] = x[i, py, :, px, :, :]
    return out.reshape(b, l, d)


def ema_update(student, teacher, momentum):
    """encoder_decoder.py:1044-1066: t <- m t + (1-m) s for every parameter pair (zip of
    named_parameters()), and for buffers whose student name contains 'bn' and not
    'num_batches_tracked'."""
    with torch.no_grad():
        for (_, sp), (_, tp) in zip(student.named_parameters(), teacher.named_parameters()):
            tp.data.mul_(momentum).add_(sp.data, alpha=1 - momentum)
        for (sn, sb), (_, tb) in zip(student.named_buffers(), teacher.named_buffers()):
            if 'bn' in sn and 'num_batches_tracked' not in sn:
                tb.data.mul_(momentum).add_(sb.data, alpha=1 - momentum)


# ----------------------------------------------------------------------------------------
# PASA (patch-adaptive self-attention) bias, vit.py:519-535
# ----------------------------------------------------------------------------------------
def pasa_gate_u0(u, adaptive=True, topk_idx=None):
    """u [B,g,g] -> (u0 [B,L], gate [B,L]) with L = 1+g*g.

    u0 = [0, flatten(u)] (vit.py:521-522).  adaptive: idx = topk(u, floor(0.5*g*g),
    largest=False) + 1; rows idx of the bias are zeroed (vit.py:525-529) => gate=0 there.
    ``topk_idx`` lets a caller inject the index set (tie order is implementation defined,
    Appendix B-1)."""
    b = u.shape[0]
    flat = u.reshape(b, -1).float()
    u0 = torch.cat((torch.zeros(b, 1, dtype=flat.dtype, device=flat.device), flat), -1)
    gate = torch.ones_like(u0)
    if adaptive:
        if topk_idx is None:
            topk_idx = torch.topk(flat, int(0.5 * flat.shape[-1]), dim=-1, largest=False)[1]
        gate[torch.arange(b, device=gate.device).unsqueeze(1), topk_idx.to(gate.device) + 1] = 0
    return u0, gate


# ----------------------------------------------------------------------------------------
# modules (same state_dict keys as the reference)
# ----------------------------------------------------------------------------------------
class _MHAParams(nn.Module):
    """Parameter container named like torch nn.MultiheadAttention."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.)


class _Attn(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.attn = _MHAParams(d)
        self.heads = heads

    def forward(self, x, identity, u0=None, gate=None, weight=0.0):
        """mmcv MultiheadAttention(batch_first) called at vit.py:119; the additive float
        mask of vit.py:531-535 is the rank-1 form w*gate[b,q]*u0[b,k], equal for all heads."""
        b, l, d = x.shape
        h, hd = self.heads, d // self.heads
        qkv = F.linear(x, self.attn.in_proj_weight, self.attn.in_proj_bias)
        q, k, v = qkv.split(d, dim=-1)
        q = q.view(b, l, h, hd).transpose(1, 2) * (1.0 / math.sqrt(hd))
        k = k.view(b, l, h, hd).transpose(1, 2)
        v = v.view(b, l, h, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        if u0 is not None:
            s = s + (weight * gate.unsqueeze(-1) * u0.unsqueeze(1)).unsqueeze(1)
        p = torch.softmax(s, dim=-1)
        o = (p @ v).transpose(1, 2).reshape(b, l, d)
        return identity + self.attn.out_proj(o)


class _FFN(nn.Module):
    def __init__(self, d, hidden):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(d, hidden), nn.GELU(), nn.Dropout(0.)),
            nn.Linear(hidden, d), nn.Dropout(0.))

    def forward(self, x, identity):
        return identity + self.layers(x)


class _EncoderLayer(nn.Module):
    """vit.py:113-127 (pre-LN block)."""

    def __init__(self, d, heads, hidden, eps):
        super().__init__()
        self.ln1 = nn.LayerNorm(d, eps=eps)
        self.attn = _Attn(d, heads)
        self.ln2 = nn.LayerNorm(d, eps=eps)
        self.ffn = _FFN(d, hidden)

    def forward(self, x, u0, gate, weight):
        x = self.attn(self.ln1(x), identity=x, u0=u0, gate=gate, weight=weight)
        x = self.ffn(self.ln2(x), identity=x)
        return x


class _PatchEmbed(nn.Module):
    def __init__(self, in_ch, d, patch):
        super().__init__()
        self.projection = nn.Conv2d(in_ch, d, patch, patch)
        self.patch = patch

    def forward(self, x):
        """embed.py:183-204 with AdaptivePadding('corner') (embed.py:58-80)."""
        h, w = x.shape[-2:]
        ph = (-h) % self.patch
        pw = (-w) % self.patch
        if ph or pw:
            x = F.pad(x, [0, pw, 0, ph])
        x = self.projection(x)
        hw = (x.shape[2], x.shape[3])
        return x.flatten(2).transpose(1, 2), hw


class OracleViT(nn.Module):
    """vit.py:479-570, default branch (pos_embed added, cls token, no final norm)."""

    def __init__(self, img_size=(512, 512), patch_size=16, in_channels=3, embed_dims=768,
                 num_layers=12, num_heads=12, mlp_ratio=4, out_indices=(4, 7, 9, 11),
                 eps=1e-6, **_ignored):
        super().__init__()
        if isinstance(img_size, int):
            img_size = (img_size, img_size)
        self.img_size, self.patch_size, self.num_heads = img_size, patch_size, num_heads
        self.out_indices = list(out_indices)
        self.patch_embed = _PatchEmbed(in_channels, embed_dims, patch_size)
        n = (img_size[0] // patch_size) * (img_size[1] // patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dims))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dims))
        self.layers = nn.ModuleList(
            [_EncoderLayer(embed_dims, num_heads, mlp_ratio * embed_dims, eps)
             for _ in range(num_layers)])

    def init_weights(self):
        """vit.py:396-414 random-init branch."""
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
                nn.init.constant_(m.bias, 0.)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0.)

    def forward(self, inputs, attn_mask=None, attn_mask_weight=0.0, adaptive_attn_mask=False,
                topk_idx=None):
        b = inputs.shape[0]
        x, hw = self.patch_embed(inputs)
        x = torch.cat((self.cls_token.expand(b, -1, -1), x), dim=1)
        assert x.shape[1] == self.pos_embed.shape[1], 'pos_embed resize not on the hot path'
        x = x + self.pos_embed
        u0 = gate = None
        if attn_mask is not None:
            u0, gate = pasa_gate_u0(attn_mask, adaptive_attn_mask, topk_idx)
        outs = []
        for i, layer in enumerate(self.layers):
            x = layer(x, u0, gate, attn_mask_weight)
            if i in self.out_indices:
                out = x[:, 1:]
                outs.append(out.reshape(b, hw[0], hw[1], -1).permute(0, 3, 1, 2).contiguous())
        return tuple(outs)


class _ConvBN(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=(k - 1) // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class OracleSETRUPHead(nn.Module):
    """setr_up_head.py:28-111 + decode_head.py (forward / forward_train / losses)."""

    def __init__(self, in_channels=768, channels=256, num_classes=21, num_convs=4, up_scale=2,
                 kernel_size=3, in_index=3, loss_weight=1.0, align_corners=False,
                 ignore_index=255, **_ignored):
        super().__init__()
        self.in_index, self.up_scale, self.num_classes = in_index, up_scale, num_classes
        self.loss_weight, self.ignore_index, self.align_corners = loss_weight, ignore_index, align_corners
        self.norm = nn.LayerNorm(in_channels, eps=1e-6)
        self.up_convs = nn.ModuleList()
        cin = in_channels
        for _ in range(num_convs):
            self.up_convs.append(nn.Sequential(_ConvBN(cin, channels, kernel_size)))
            cin = channels
        self.conv_seg = nn.Conv2d(channels, num_classes, 1)

    def init_weights(self):
        nn.init.constant_(self.norm.weight, 1.0)
        nn.init.constant_(self.norm.bias, 0.)
        for uc in self.up_convs:
            nn.init.kaiming_normal_(uc[0].conv.weight, mode='fan_out', nonlinearity='relu')
        nn.init.normal_(self.conv_seg.weight, 0, 0.01)
        nn.init.constant_(self.conv_seg.bias, 0)

    def forward(self, inputs, PatchMix_N=0, PatchMixIndex=None):
        x = inputs[self.in_index]
        n, c, h, w = x.shape
        x = x.reshape(n, c, h * w).transpose(2, 1).contiguous()
        if PatchMix_N != 0:
            x = token_unshuffle(x, PatchMixIndex, PatchMix_N)
        x = self.norm(x)
        x = x.transpose(1, 2).reshape(n, c, h, w).contiguous()
        for uc in self.up_convs:
            x = uc(x)
            size = [int(t * float(self.up_scale)) for t in x.shape[-2:]]   # ops/wrappers.py:46-51
            x = F.interpolate(x, size, None, 'bilinear', self.align_corners)
        return self.conv_seg(x)

    def forward_get_logits(self, inputs, img_metas):
        """decode_head.py:261-271."""
        if 'PatchMix_N' not in img_metas[0]:
            return self.forward(inputs)
        idx = torch.stack([torch.as_tensor(m['PatchMixIndex']) for m in img_metas])
        return self.forward(inputs, PatchMix_N=img_metas[-1]['PatchMix_N'], PatchMixIndex=idx)

    def losses(self, seg_logit, seg_label):
        """decode_head.py:318-355 (resize to label size is the identity on this path)."""
        if seg_logit.shape[2:] != seg_label.shape[2:]:
            seg_logit = F.interpolate(seg_logit, seg_label.shape[2:], None, 'bilinear',
                                      self.align_corners)
        return {'loss_ce': cross_entropy_mean_all(seg_logit, seg_label.squeeze(1),
                                                  self.ignore_index, self.loss_weight)}

    def forward_train(self, inputs, img_metas, gt):
        return self.losses(self.forward_get_logits(inputs, img_metas), gt)


class OracleEncoderDecoder(nn.Module):
    """encoder_decoder.py:386-514 (forward_train) and :516-687 (foward_unsup_train) for the
    flags the shipped configs use: ema, unsup_weight, unsup_confidence,
    attn_mask_seperate_head, attn_mask_weight, adaptive_attn_mask, use_CutMix,
    use_PatchShuffle_w_Cutmix, PatchMix_N, negative_class_ranking('unsup_only'),
    fdrop_loss_weight, strong_aug_prob, cutout_area, patchmix_ratio, patchsize."""

    def __init__(self, backbone, decode_head, auxiliary_head=None, ema=False,
                 ema_momentum=0.999, unsup_weight=2.0, unsup_confidence=0.75,
                 strong_aug_prob=0.5, cutout_area=2, use_CutMix=False, PatchMix_N=8,
                 patchmix_ratio=0.5, patchsize=16, use_PatchShuffle_w_Cutmix=False,
                 adaptive_attn_mask=False, attn_mask_weight=50, attn_mask_seperate_head=False,
                 negative_class_ranking=False, negative_class_ranking_mode='sup_only',
                 fdrop_loss_weight=0.5, **_ignored):
        super().__init__()

        def mk_head(cfg):
            cfg = dict(cfg)
            typ = cfg.pop('type', 'SETRUPHead')
            lw = cfg.pop('loss_decode', {}).get('loss_weight', 1.0)
            if typ == 'SegformerHead':          # oracle only, no CUDA path yet
                from .segformer_oracle import OracleSegformerHead
                return OracleSegformerHead(loss_weight=lw, **cfg)
            return OracleSETRUPHead(loss_weight=lw, **cfg)

        def mk_bb(cfg):
            cfg = dict(cfg)
            typ = cfg.pop('type', 'VisionTransformer')
            cfg.pop('norm_cfg', None)
            if typ == 'MixVisionTransformer':
                from .segformer_oracle import OracleMiT
                return OracleMiT(**cfg)
            return OracleViT(**cfg)

        self.backbone = mk_bb(backbone)
        self.decode_head = mk_head(decode_head)
        self.auxiliary_head = nn.ModuleList([mk_head(c) for c in (auxiliary_head or [])])
        self.ema = ema
        if ema:
            self.backbone_ema = mk_bb(backbone)
            self.decode_head_ema = mk_head(decode_head)
            for p in list(self.backbone_ema.parameters()) + list(self.decode_head_ema.parameters()):
                p.detach_()
        self.momentum = ema_momentum
        self.unsup_weight, self.unsup_confidence = unsup_weight, unsup_confidence
        self.strong_aug_prob, self.cutout_area, self.use_CutMix = strong_aug_prob, cutout_area, use_CutMix
        self.PatchMix_N, self.patchmix_ratio, self.patchsize = PatchMix_N, patchmix_ratio, patchsize
        self.use_PatchShuffle_w_Cutmix = use_PatchShuffle_w_Cutmix
        self.adaptive_attn_mask, self.attn_mask_weight = adaptive_attn_mask, attn_mask_weight
        self.attn_mask_seperate_head = attn_mask_seperate_head
        self.negative_class_ranking = negative_class_ranking
        self.negative_class_ranking_mode = negative_class_ranking_mode
        self.fdrop_loss_weight = fdrop_loss_weight

    def init_weights(self):
        for m in [self.backbone, self.decode_head, *self.auxiliary_head] + \
                 ([self.backbone_ema, self.decode_head_ema] if self.ema else []):
            m.init_weights()

    # -- encoder_decoder.py:906-954
    def compute_pseudo_loss(self, feat, metas, teacher):
        out = {}
        z_s = self.decode_head.forward_get_logits(feat, metas)
        out['loss_seg_unsup'] = cross_entropy_mean_all(z_s, teacher['hard_seg_label'], 255)
        out['mask_ratio'] = teacher['conf_mask'].sum().float() / teacher['conf_mask'].numel()
        if self.negative_class_ranking and self.negative_class_ranking_mode in ('unsup_only', 'both'):
            out['loss_ncr_unsup'] = ncr_unsup_only(z_s, teacher['seg_logits'],
                                                   teacher['hard_seg_label'])
        return out

    def forward_train(self, img, img_metas, gt_semantic_seg, topk_idx=None, record=None,
                      teacher_override=None):
        """Returns the reference's loss dict.  ``record`` (a dict) receives intermediates.
        ``teacher_override`` = dict(seg_logits, hard_seg_label, conf_mask): parity tests pin the
        teacher outputs (e.g. to the ones the CUDA path produced) so that the student-side
        arithmetic is compared given IDENTICAL pseudo labels."""
        tags = [m['tag'] for m in img_metas]
        groups = {}
        for t in dict.fromkeys(tags):
            sel = [i for i, tt in enumerate(tags) if tt == t]
            groups[t] = dict(img=img[sel], gt=gt_semantic_seg[sel],
                             metas=[img_metas[i] for i in sel])
        losses = {}
        if self.ema:   # :416-423 -- EMA happens BEFORE any forward
            ema_update(self.backbone, self.backbone_ema, self.momentum)
            ema_update(self.decode_head, self.decode_head_ema, self.momentum)
        if 'sup' in groups:   # :426-441
            g = groups['sup']
            feats = self.backbone(g['img'])
            dec = self.decode_head.forward_train(feats, g['metas'], g['gt'])
            for i, aux in enumerate(self.auxiliary_head):
                losses[f'aux_{i}.loss_ce'] = aux.forward_train(feats, g['metas'], g['gt'])['loss_ce']
            losses['decode.loss_ce'] = dec['loss_ce']
        if 'unsup_student' in groups and self.unsup_weight != 0:   # :488-512
            un = self.forward_unsup_train(groups['unsup_teacher'], groups['unsup_student'],
                                          topk_idx=topk_idx, record=record,
                                          teacher_override=teacher_override)
            for k in un:
                if 'loss' in k:   # structual_utils.py:132-154
                    un[k] = un[k] * self.unsup_weight
            losses.update(un)
        return losses

    def forward_unsup_train(self, teacher_data, student_data, topk_idx=None, record=None,
                            teacher_override=None):
        loss_unsup = {}
        tnames = [m['filename'] for m in teacher_data['metas']]
        snames = [m['filename'] for m in student_data['metas']]
        tidx = [tnames.index(n) for n in snames]
        with torch.no_grad():   # :523-539, teacher in eval mode
            self.backbone_ema.eval()
            self.decode_head_ema.eval()
            timg = teacher_data['img'][tidx]
            feat_t = self.backbone_ema(timg)
            z_t = self.decode_head_ema.forward(feat_t)
            hard, conf, _ = pseudo_label(z_t, self.unsup_confidence)
            self.backbone_ema.train()
            self.decode_head_ema.train()
        if teacher_override is not None:
            z_t = teacher_override['seg_logits']
            hard = teacher_override['hard_seg_label'].clone()
            conf = teacher_override['conf_mask']
        teacher = dict(seg_logits=z_t, hard_seg_label=hard, conf_mask=conf)
        simg = student_data['img']
        smetas = student_data['metas']
        if record is not None:
            record.update(teacher_logits=z_t, hard0=hard.clone(), conf=conf)
        # :548-551 -- a confidence map at the image resolution is "VIT style" (patches of
        # self.patchsize); SegFormer's quarter-resolution map uses patches of 8
        apatch = self.patchsize if conf.shape[-1] == simg.shape[-1] else 8
        if self.attn_mask_seperate_head:   # :547-567
            u = patch_unconfidence(conf, apatch)
            feat = self.backbone(simg, attn_mask=u, attn_mask_weight=self.attn_mask_weight,
                                 adaptive_attn_mask=self.adaptive_attn_mask, topk_idx=topk_idx)
            loss_unsup['loss_seg_unsup_attn_mask'] = \
                self.compute_pseudo_loss(feat, smetas, teacher)['loss_seg_unsup'] * 0.5
        if self.use_CutMix:   # :604-607
            if np.random.uniform(0, 1) < self.strong_aug_prob:
                boxes = [cutout_box(simg.shape[2:], self.cutout_area) for _ in range(simg.shape[0])]
                simg, teacher['hard_seg_label'] = cutmix(simg, teacher['hard_seg_label'], boxes)
        if self.use_PatchShuffle_w_Cutmix:   # :633-638
            if np.random.uniform(0, 1) < self.strong_aug_prob:
                boxes = [cutout_box(simg.shape[2:], self.cutout_area) for _ in range(simg.shape[0])]
                simg, teacher['hard_seg_label'] = cutmix(simg, teacher['hard_seg_label'], boxes)
            block = self.patchsize * self.PatchMix_N
            nblocks = (simg.shape[2] // block) * (simg.shape[3] // block)
            perms = draw_patchshuffle_perms(simg.shape[0], nblocks, self.patchmix_ratio)
            simg = patchshuffle(simg, perms, block)
            for i, m in enumerate(smetas):
                m['PatchMixIndex'] = perms[i]
                m['PatchMix_N'] = self.PatchMix_N
        if record is not None:
            record.update(student_img_mixed=simg, hard_mixed=teacher['hard_seg_label'])
        if not self.attn_mask_seperate_head:   # :650-670 (MT as shipped: loss-less pass)
            u = patch_unconfidence(conf, apatch)
            feat = self.backbone(simg, attn_mask=u, attn_mask_weight=self.attn_mask_weight,
                                 adaptive_attn_mask=self.adaptive_attn_mask, topk_idx=topk_idx)
        else:   # :671-677
            feat = self.backbone(simg)
        if self.attn_mask_seperate_head:   # :681-685 (use_fdrop is never set on this path)
            ls = self.compute_pseudo_loss(feat, smetas, teacher)
            if self.negative_class_ranking:
                loss_unsup['loss_ncr_unsup'] = ls['loss_ncr_unsup'] * 0.5
            loss_unsup['loss_seg_unsup'] = ls['loss_seg_unsup'] * self.fdrop_loss_weight
        return loss_unsup


def parse_losses(losses):
    """base.py:230-253: loss = sum of entries whose key contains 'loss'."""
    return sum(v for k, v in losses.items() if 'loss' in k)


# ----------------------------------------------------------------------------------------
# synthetic workload -- shared by tests and bench
# ----------------------------------------------------------------------------------------
def synthetic_batch(n_sup, n_unsup, size, num_classes, seed=1999, grid=32):
    """img ~ N(0,1); gt piecewise-constant on a ``grid``-px grid over U{0..C-1} with a 5%
    border of 255; tagged metas in the collate(flatten=True) order sup.., then per unsup
    sample (student, teacher)."""
    g = torch.Generator().manual_seed(seed)
    n = n_sup + 2 * n_unsup
    img = torch.randn(n, 3, size, size, generator=g)
    cells = max(size // grid, 1)
    lab = torch.randint(0, num_classes, (n, 1, cells, cells), generator=g)
    gt = F.interpolate(lab.float(), size=(size, size), mode='nearest').long()
    bw = max(int(round(size * 0.05)), 1)
    gt[:, :, :bw] = 255
    gt[:, :, -bw:] = 255
    gt[:, :, :, :bw] = 255
    gt[:, :, :, -bw:] = 255
    metas = []
    for i in range(n_sup):
        metas.append(dict(filename=f'sup_{i}.jpg', tag='sup'))
    for i in range(n_unsup):
        metas.append(dict(filename=f'unsup_{i}.jpg', tag='unsup_student'))
        metas.append(dict(filename=f'unsup_{i}.jpg', tag='unsup_teacher'))
    for m in metas:
        m.update(ori_shape=(size, size, 3), img_shape=(size, size, 3), pad_shape=(size, size, 3),
                 scale_factor=1.0, flip=False, flip_direction=None)
    return img, gt, metas


def seed_host_rng(seed=1999):
    """apis/train.py:51-67 set_random_seed."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
