"""SETR-PUP / Naive head -- B200-native mirror of
``mmseg/models/decode_heads/setr_up_head.py:10-111`` (same kwargs and ``state_dict`` keys:
``norm.*``, ``up_convs.{j}.0.conv.weight``, ``up_convs.{j}.0.bn.*``, ``conv_seg.*``).

Execution plan (NHWC, compute dtype):
  tokens --[row-gathered LayerNorm = feature tap + PatchMix un-shuffle + LN]--> [B,g,g,C]
  stage j < n-1 : conv3x3 -> BN stats -> fused BN+ReLU+bilinear(x s)
  last stage    : conv3x3 -> BN stats -> fused BN+ReLU+conv_seg(1x1) -> bilinear(x s) -> NCHW fp32
"""
import math

import torch
import torch.nn as nn

from .. import ops
from ..builder import HEADS
from .decode_head import BaseDecodeHead


class _ConvModule(nn.Module):
    """Parameter holder named like mmcv ``ConvModule`` (conv without bias -> bn -> ReLU)."""

    def __init__(self, cin, cout, k, norm_cfg):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=(k - 1) // 2, bias=False)
        cfg = dict(norm_cfg)
        typ = cfg.pop('type')
        cfg.pop('requires_grad', None)
        if typ not in ('BN', 'SyncBN'):
            raise NotImplementedError(f'norm {typ}: only BN / SyncBN ConvModules are on the hot path')
        self.sync = typ == 'SyncBN'
        self.bn = nn.BatchNorm2d(cout, **cfg)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')


class _Upsample(nn.Module):
    def __init__(self, scale_factor, mode, align_corners):
        super().__init__()
        self.scale_factor, self.mode, self.align_corners = float(scale_factor), mode, align_corners


@HEADS.register_module()
class SETRUPHead(BaseDecodeHead):
    def __init__(self, norm_layer=dict(type='LN', eps=1e-6, requires_grad=True), num_convs=1,
                 up_scale=4, kernel_size=3, use_addition_up_scale=False,
                 init_cfg=[dict(type='Constant', val=1.0, bias=0, layer='LayerNorm'),
                           dict(type='Normal', std=0.01, override=dict(name='conv_seg'))],
                 **kwargs):
        assert kernel_size in [1, 3], 'kernel_size must be 1 or 3.'
        super().__init__(init_cfg=init_cfg, **kwargs)
        assert isinstance(self.in_channels, int)
        if kernel_size != 3 or use_addition_up_scale or self.align_corners:
            raise NotImplementedError('SETR-PUP path: kernel_size=3, align_corners=False, no extra upscale')
        if self.norm_cfg is None or self.act_cfg is None or self.act_cfg.get('type') != 'ReLU':
            raise NotImplementedError('ConvModule must be conv -> BN/SyncBN -> ReLU')
        if int(up_scale) != up_scale or num_convs < 1:
            raise NotImplementedError('integer up_scale and num_convs >= 1 required')
        self.norm = nn.LayerNorm(self.in_channels, eps=norm_layer.get('eps', 1e-5))
        self.up_scale = int(up_scale)
        self.up_convs = nn.ModuleList()
        cin = self.in_channels
        for _ in range(num_convs):
            self.up_convs.append(nn.Sequential(
                _ConvModule(cin, self.channels, kernel_size, self.norm_cfg),
                _Upsample(up_scale, 'bilinear', self.align_corners)))
            cin = self.channels
        self._row_maps = {}

    def init_weights(self):
        nn.init.constant_(self.norm.weight, 1.0)
        nn.init.constant_(self.norm.bias, 0.)
        nn.init.normal_(self.conv_seg.weight, mean=0, std=0.01)
        nn.init.constant_(self.conv_seg.bias, 0)
        for uc in self.up_convs:
            nn.init.kaiming_normal_(uc[0].conv.weight, a=0, mode='fan_out', nonlinearity='relu')
            nn.init.constant_(uc[0].bn.weight, 1.)
            nn.init.constant_(uc[0].bn.bias, 0.)

    # -- row map: dst token row -> src row of the backbone's [B*L, D] matrix ------------------
    def _row_map(self, B, g, has_cls, device, PatchMix_N, PatchMixIndex, b0=0, gw=None):
        """Feature tap (+1 skips the cls row, vit.py:556-562) and, when PatchMix_N != 0, the
        inverse block permutation of decode_head.py:186-212: out_block[perm[p]] = in_block[p].
        ``b0``: index of this group's first image in the backbone's token matrix."""
        ntok = g * (g if gw is None else gw)           # (g, gw): rows x columns of the token grid
        Ls = ntok + (1 if has_cls else 0)
        off = (1 if has_cls else 0) + b0 * Ls
        if PatchMix_N == 0:
            key = (B, g, gw, has_cls, str(device), b0)
            m = self._row_maps.get(key)
            if m is None:
                pos = torch.arange(ntok, dtype=torch.int64)
                m = (torch.arange(B, dtype=torch.int64).view(B, 1) * Ls + off + pos.view(1, -1))
                m = m.reshape(-1).to(torch.int32).to(device)
                self._row_maps[key] = m
            return m
        if gw is not None and gw != g:
            raise NotImplementedError('PatchMix un-shuffle needs a square token grid (training crops are square)')
        n = int(PatchMix_N)
        gb = g // n
        perm = torch.as_tensor(PatchMixIndex)
        # the index arithmetic runs where the permutation lives: on the host for the reference's
        # per-meta tensors, on the device (a dozen tiny launches, no host round trip -> CUDA-graph
        # capturable) for the resident copy of a pre-drawn step
        pdev = perm.device if perm.is_cuda else torch.device('cpu')
        perm = perm.to(torch.int64).reshape(B, gb * gb)
        key = ('grid', B, g, n, has_cls, b0, str(pdev))
        grid = self._row_maps.get(key)
        if grid is None:
            ty, tx = torch.meshgrid(torch.arange(g), torch.arange(g), indexing='ij')
            qb = ((ty // n) * gb + (tx // n)).reshape(-1)                   # destination block of each token
            grid = (qb.to(pdev), (ty % n).reshape(1, -1).to(pdev), (tx % n).reshape(1, -1).to(pdev),
                    torch.arange(gb * gb, dtype=torch.int64, device=pdev).expand(B, -1).contiguous(),
                    (torch.arange(B, dtype=torch.int64, device=pdev).view(B, 1) * Ls + off))
            self._row_maps[key] = grid
        qb, ry, rx, ar, base = grid
        inv = torch.empty_like(perm)
        inv.scatter_(1, perm, ar)
        src_blk = inv[:, qb]                                 # [B, g*g] source block
        sy = (src_blk // gb) * n + ry
        sx = (src_blk % gb) * n + rx
        m = base + sy * g + sx
        return m.reshape(-1).to(torch.int32).to(device, non_blocking=True)

    def _group_info(self):
        if self.up_convs[0][0].sync and self.training:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                peer = None
                if dist.get_backend() == 'nccl':
                    from ..parallel import PeerAllReduce
                    peer = PeerAllReduce.get(None)
                return dict(world=dist.get_world_size(), group=None, peer=peer)
        return None

    def forward(self, x, PatchMix_N=0, PatchMixIndex=None, return_last_feat=False):
        """setr_up_head.py:92-111."""
        if return_last_feat:
            raise NotImplementedError('return_last_feat is a visualisation path')
        x = self._transform_inputs(x)
        tok = getattr(x, '_s4_tokens', None)
        b0 = 0
        if tok is not None:
            x2d, B, L = tok[:3]
            b0 = tok[3] if len(tok) > 3 else 0
            gh, gw = getattr(x, '_s4_hw', None) or (int(math.isqrt(L - 1)),) * 2
            assert gh * gw == L - 1
            has_cls = True
        else:   # a foreign NCHW tensor: flatten to tokens (copy) in the compute dtype
            B, Cc, gh, gw = x.shape
            x2d = ops.cast(x.permute(0, 2, 3, 1).reshape(B * gh * gw, Cc).contiguous(), ops.compute_dtype())
            has_cls = False
        row_map = self._row_map(B, gh, has_cls, x2d.device, PatchMix_N, PatchMixIndex, b0, gw=gw)
        Ltok = gh * gw + 1
        y = ops.HeadLNFn.apply(x2d, self, row_map, B, Ltok, self.norm.weight)
        H, W = gh, gw
        s = self.up_scale
        gi = self._group_info()
        n = len(self.up_convs)
        for j, uc in enumerate(self.up_convs):
            if j < n - 1:
                y = ops.ConvBNReLUUpFn.apply(y, uc[0], B, H, W, s, self.training, gi)
                H, W = H * s, W * s
            else:
                y = ops.ConvBNReLUClsUpFn.apply(y, uc[0], self.conv_seg, B, H, W, s, self.training, gi)
        return y

    def cls_seg(self, feat):
        raise NotImplementedError('cls_seg is fused into the last up-conv stage of forward()')
