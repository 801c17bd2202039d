"""``CrossEntropyLoss`` -- mirror of ``mmseg/models/losses/cross_entropy_loss.py:196-297`` for
the branch the S4Former configs use (``use_sigmoid=False``, no class weights, reduction
'mean', ``avg_non_ignore=False``): ``loss_weight * mean_{ALL pixels}(CE(ignore_index))``
(cross_entropy_loss.py:45-61, losses/utils.py:65-69), computed by ``s4_ce_ncr``."""
import warnings

import torch.nn as nn

from .. import ops
from ..builder import LOSSES


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None,
                 loss_weight=1.0, loss_name='loss_ce', avg_non_ignore=False):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        if use_sigmoid or use_mask or class_weight is not None or avg_non_ignore or reduction != 'mean':
            raise NotImplementedError(
                'only softmax CE with reduction="mean", avg_non_ignore=False is on the S4Former path')
        self.use_sigmoid, self.use_mask, self.reduction = use_sigmoid, use_mask, reduction
        self.loss_weight, self.class_weight, self.avg_non_ignore = loss_weight, class_weight, avg_non_ignore
        if not self.avg_non_ignore and self.reduction == 'mean':
            warnings.warn('Default ``avg_non_ignore`` is False, if you would like to ignore the certain '
                          'label and average loss over non-ignore labels, which is the same with PyTorch '
                          'official cross_entropy, set ``avg_non_ignore=True``.')
        self._loss_name = loss_name

    def extra_repr(self):
        return f'avg_non_ignore={self.avg_non_ignore}'

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None,
                ignore_index=-100, **kwargs):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        if weight is not None or avg_factor is not None or reduction_override not in (None, 'mean'):
            raise NotImplementedError('pixel weights / avg_factor / reduction overrides are not used '
                                      'on the S4Former path (sampler=None)')
        return ops.cross_entropy(cls_score.float(), label, self.loss_weight, ignore_index)

    @property
    def loss_name(self):
        return self._loss_name
