"""``BaseSegmentor`` -- mirror of ``mmseg/models/segmentors/base.py`` for the train path:
``train_step`` (:155-206), ``val_step``, ``forward`` and ``_parse_losses`` (:230-274).

Differences, all on purpose and numerically neutral:
 * the per-iteration debug dump (``np.save`` of the batch + JSON of the metas, base.py:181-195)
   is not performed;
 * ``_parse_losses`` packs all log variables into ONE tensor: one all-reduce and one
   device->host copy instead of one all-reduce + ``.item()`` per key (N3 in SURVEY.md).
"""
from abc import ABCMeta, abstractmethod
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn


_pending_key_checks = []      # deferred cross-rank key-count checks of _parse_losses(sync=False)


class BaseSegmentor(nn.Module, metaclass=ABCMeta):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        self.fp16_enabled = False

    @property
    def with_neck(self):
        return hasattr(self, 'neck') and self.neck is not None

    @property
    def with_auxiliary_head(self):
        return hasattr(self, 'auxiliary_head') and self.auxiliary_head is not None

    @property
    def with_decode_head(self):
        return hasattr(self, 'decode_head') and self.decode_head is not None

    @abstractmethod
    def extract_feat(self, imgs):
        pass

    @abstractmethod
    def forward_train(self, imgs, img_metas, **kwargs):
        pass

    def forward_test(self, imgs, img_metas, **kwargs):
        """base.py:65-99.  ``img_metas`` entries may be mmcv ``DataContainer``s (``._data[0]``, what the
        shipped code assumes) or plain lists of dicts."""
        for var, name in [(imgs, 'imgs'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError(f'{name} must be a list, but got {type(var)}')
        num_augs = len(imgs)
        if num_augs != len(img_metas):
            raise ValueError(f'num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})')
        metas = [m._data[0] if hasattr(m, '_data') else m for m in img_metas]
        for img_meta in metas:
            for key in ('ori_shape', 'img_shape', 'pad_shape'):
                vals = [_[key] for _ in img_meta]
                assert all(v == vals[0] for v in vals)
        if num_augs == 1:
            return self.simple_test(imgs[0], metas[0], **kwargs)
        return self.aug_test(imgs, metas, **kwargs)

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def train_step(self, data_batch, optimizer, **kwargs):
        data_batch = dict(data_batch)
        data_batch['iter'] = kwargs.get('iter', 0)
        losses = self(**data_batch)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data_batch['img_metas']))

    def val_step(self, data_batch, optimizer=None, **kwargs):
        losses = self(**data_batch)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars={k + '_val': v for k, v in log_vars.items()},
                    num_samples=len(data_batch['img_metas']))

    @staticmethod
    def _parse_losses(losses, sync=True):
        log_vars = OrderedDict()
        def _mean(t):                      # the path's losses are already scalars: no launch for them
            return t if t.dim() == 0 else t.mean()
        for loss_name, loss_value in losses.items():
            if isinstance(loss_value, torch.Tensor):
                log_vars[loss_name] = _mean(loss_value)
            elif isinstance(loss_value, list):
                log_vars[loss_name] = sum(_mean(_loss) for _loss in loss_value)
            else:
                raise TypeError(f'{loss_name} is not a tensor or list of tensors')
        loss = sum(_value for _key, _value in log_vars.items() if 'loss' in _key)
        log_vars['loss'] = loss
        names = list(log_vars.keys())
        packed = torch.stack([v.detach().float().reshape(()) for v in log_vars.values()])
        if dist.is_available() and dist.is_initialized():
            n = packed.new_full((1,), float(len(names)))     # (a fill, not a host->device copy:
            packed = torch.cat([packed, n])                  #  the step may be under graph capture)
            dist.all_reduce(packed)
            ws = dist.get_world_size()
            capturing = packed.is_cuda and torch.cuda.is_current_stream_capturing()
            # the reference's cross-rank key-count assertion (base.py:263).  Reading the count is a
            # device->host sync: with sync=False it is checked one call LATE (the value is long
            # since final by then), so the host keeps enqueueing instead of stalling mid-step
            check = (packed[-1], len(names) * ws, ','.join(names))
            if capturing:       # no host read inside a capture; a replayed step has static keys
                todo = []
                _pending_key_checks[:] = []
            else:
                todo = [check] if sync else _pending_key_checks[:]
                if not sync:
                    _pending_key_checks[:] = [check]
            for cnt, want, nm in todo:
                assert int(round(float(cnt))) == want, \
                    'loss log variables are different across GPUs!\n' + nm
            packed = packed[:-1] / ws
        if sync:
            vals = packed.tolist()        # ONE device->host copy
            out = OrderedDict(zip(names, vals))
        else:
            out = OrderedDict(zip(names, packed.unbind(0)))
        return loss, out
