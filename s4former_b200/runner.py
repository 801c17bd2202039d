"""One training iteration as the reference's runner drives it:
``model.train_step`` (mmseg/models/segmentors/base.py:155-206) -> ``loss.backward()`` -> gradient
all-reduce -> ``optimizer.step()`` (mmcv ``OptimizerHook.after_train_iter`` + DDP reducer) with
the poly LR schedule.  ``TrainStep`` is the call a user makes per iteration; everything it does
on the device runs in the CUDA library.
"""
import torch
import torch.distributed as dist

from . import configs
from .optim import FusedSGD
from .parallel import GradReducer


class TrainStep:
    def __init__(self, model, optimizer_cfg=None, lr_cfg=None, max_iters=configs.MAX_ITERS):
        ocfg = dict(configs.OPTIMIZER if optimizer_cfg is None else optimizer_cfg)
        lcfg = dict(configs.LR_CONFIG if lr_cfg is None else lr_cfg)
        assert ocfg.get('type', 'SGD') == 'SGD' and lcfg.get('policy', 'poly') == 'poly'
        self.model = model
        self.optimizer = FusedSGD(
            model.named_parameters(), lr=ocfg.get('lr', 1e-3), momentum=ocfg.get('momentum', 0.9),
            weight_decay=ocfg.get('weight_decay', 0.0),
            custom_keys=(ocfg.get('paramwise_cfg') or {}).get('custom_keys'),
            max_iters=max_iters, power=lcfg.get('power', 0.9), min_lr=lcfg.get('min_lr', 1e-4))
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.reducer = GradReducer(model, self.optimizer.grads) if self.world > 1 else None
        self.device = next(model.parameters()).device

    def __call__(self, img, img_metas, gt_semantic_seg, it, sync=True):
        """Device-resident batch -> one optimisation step.  Returns ``(loss, log_vars)``; with
        ``sync=False`` the log variables stay device tensors (no host synchronisation)."""
        self.optimizer.zero_grad()
        losses = self.model(img, img_metas, return_loss=True, gt_semantic_seg=gt_semantic_seg, iter=it)
        loss, log_vars = self.model._parse_losses(losses, sync=False)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finalize()
        self.optimizer.step(it)
        if sync:
            packed = torch.stack(list(log_vars.values())).tolist()     # one device->host copy
            log_vars = type(log_vars)(zip(log_vars.keys(), packed))
        return loss, log_vars

    def step_from_host(self, img_host, img_metas, gt_host, it):
        """End-to-end iteration: pinned host batch -> device, step, log variables back on the host
        (what the runner's dataloader scatter + ``log_vars`` ``.item()`` do in the reference)."""
        img = img_host.to(self.device, non_blocking=True)
        gt = gt_host.to(self.device, non_blocking=True)
        return self(img, img_metas, gt, it, sync=True)
