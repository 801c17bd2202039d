"""``BaseDecodeHead`` -- mirror of ``mmseg/models/decode_heads/decode_head.py`` (methods the
train step calls: ``forward_train`` :225-259, ``forward_get_logits`` :261-271,
``forward_test`` :295-309, ``cls_seg`` :311-316, ``losses`` :318-355)."""
from abc import ABCMeta, abstractmethod

import torch
import torch.nn as nn

from ..builder import build_loss


class BaseDecodeHead(nn.Module, metaclass=ABCMeta):
    def __init__(self, in_channels, channels, *, num_classes, dropout_ratio=0.1, conv_cfg=None,
                 norm_cfg=None, act_cfg=dict(type='ReLU'), in_index=-1, input_transform=None,
                 loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                 ignore_index=255, sampler=None, align_corners=False, class_re_weight=False,
                 init_cfg=dict(type='Normal', std=0.01, override=dict(name='conv_seg')),
                 get_mean_feat=False, decoder_params=None):
        super().__init__()
        if input_transform is not None:
            raise NotImplementedError('input_transform is None in the SETR configs')
        assert isinstance(in_channels, int) and isinstance(in_index, int)
        if sampler is not None:
            raise NotImplementedError('sampler=None on the S4Former path')
        if dropout_ratio > 0:
            raise NotImplementedError('dropout_ratio=0 in every SETR-PUP config')
        self.init_cfg = init_cfg
        self.in_channels, self.input_transform, self.in_index = in_channels, input_transform, in_index
        self.channels, self.num_classes, self.dropout_ratio = channels, num_classes, dropout_ratio
        self.conv_cfg, self.norm_cfg, self.act_cfg = conv_cfg, norm_cfg, act_cfg
        self.ignore_index, self.align_corners = ignore_index, align_corners
        self.get_mean_feat = get_mean_feat
        if isinstance(loss_decode, dict):
            self.loss_decode = build_loss(loss_decode)
        elif isinstance(loss_decode, (list, tuple)):
            self.loss_decode = nn.ModuleList([build_loss(l) for l in loss_decode])
        else:
            raise TypeError(f'loss_decode must be a dict or sequence of dict, but got {type(loss_decode)}')
        self.sampler = None
        self.conv_seg = nn.Conv2d(channels, num_classes, kernel_size=1)
        self.dropout = None
        self.fp16_enabled = False

    def extra_repr(self):
        return f'input_transform={self.input_transform}, ignore_index={self.ignore_index}, ' \
               f'align_corners={self.align_corners}'

    def _transform_inputs(self, inputs):
        return inputs[self.in_index]

    @abstractmethod
    def forward(self, inputs):
        pass

    @staticmethod
    def _patchmix_index(img_metas):
        dev = img_metas[0].get('_s4_perms_dev')
        if dev is not None:           # resident copy of the same permutations (no host round trip)
            return dev, img_metas[-1]['PatchMix_N']
        idx = torch.stack([torch.as_tensor(m['PatchMixIndex']) for m in img_metas])
        return idx, img_metas[-1]['PatchMix_N']

    def forward_train(self, inputs, img_metas, gt_semantic_seg, train_cfg):
        if 'PatchMix_N' in img_metas[0]:
            idx, n = self._patchmix_index(img_metas)
            seg_logits = self.forward(inputs, PatchMix_N=n, PatchMixIndex=idx)
        else:
            seg_logits = self.forward(inputs)
        return self.losses(seg_logits, gt_semantic_seg)

    def forward_get_logits(self, inputs, train_cfg, img_metas=None):
        if 'PatchMix_N' not in img_metas[0]:
            return self.forward(inputs)
        idx, n = self._patchmix_index(img_metas)
        return self.forward(inputs, PatchMix_N=n, PatchMixIndex=idx)

    def forward_test(self, inputs, img_metas, test_cfg, return_last_feat=False):
        return self.forward(inputs, return_last_feat=return_last_feat)

    def losses(self, seg_logit, seg_label):
        """decode_head.py:318-355.  The bilinear resize to the label size is the identity on the
        train path (logits are produced at crop resolution); any other size is rejected loudly."""
        loss = dict()
        if tuple(seg_logit.shape[2:]) != tuple(seg_label.shape[2:]):
            raise NotImplementedError('logit/label size mismatch: resize in losses() is off the hot path')
        seg_label = seg_label.squeeze(1)
        losses_decode = self.loss_decode if isinstance(self.loss_decode, nn.ModuleList) else [self.loss_decode]
        for loss_decode in losses_decode:
            v = loss_decode(seg_logit, seg_label, weight=None, ignore_index=self.ignore_index)
            if loss_decode.loss_name not in loss:
                loss[loss_decode.loss_name] = v
            else:
                loss[loss_decode.loss_name] += v
        return loss
