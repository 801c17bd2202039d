"""One training iteration as the reference's runner drives it:
``model.train_step`` (mmseg/models/segmentors/base.py:155-206) -> ``loss.backward()`` -> gradient
all-reduce -> ``optimizer.step()`` (mmcv ``OptimizerHook.after_train_iter`` + DDP reducer) with
the poly LR schedule.  ``TrainStep`` is the call a user makes per iteration; everything it does
on the device runs in the CUDA library.
"""
import torch
import torch.distributed as dist

from . import configs, ops
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
        ops.reset_arena()
        self.stale_pending = ops.reset_pending()      # 0 in a healthy step (see ops._pending_inc)
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

    # ---- end-to-end iteration (host batch in, host log variables out) ---------------------------
    # Two persistent device staging slots (no allocator traffic in the loop): the copy of batch
    # i+1 runs on a side stream while step i computes; a slot is rewritten only after the step
    # that read it has finished on the compute stream.
    def _slot(self, k, img_host, gt_host):
        slots = self.__dict__.setdefault('_slots', [None, None])
        sl = slots[k]
        if sl is None or sl['img'].shape != img_host.shape or sl['gt'].shape != gt_host.shape:
            sl = dict(img=torch.empty(img_host.shape, dtype=img_host.dtype, device=self.device),
                      gt=torch.empty(gt_host.shape, dtype=gt_host.dtype, device=self.device),
                      free=None, ready=None)
            slots[k] = sl
        return sl

    def _start_copy(self, k, img_host, gt_host):
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        s = self._copy_stream
        sl = self._slot(k, img_host, gt_host)
        if sl['free'] is not None:
            s.wait_event(sl['free'])
        else:
            s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            sl['img'].copy_(img_host, non_blocking=True)
            sl['gt'].copy_(gt_host, non_blocking=True)
            sl['ready'] = torch.cuda.Event()
            sl['ready'].record(s)
        return sl

    def prefetch(self, img_host, gt_host):
        """Start the host->device copy of a coming batch on a side stream, so that it overlaps the
        step in flight (the dataloader's pinned batch is known one iteration ahead).  The
        ``step_from_host`` call on the same host tensors consumes it.  With two staging slots at
        most one batch besides the one being consumed can be in flight; a further call is a no-op
        (that step then copies for itself)."""
        pending = self.__dict__.setdefault('_pending', {})
        key = (img_host.data_ptr(), gt_host.data_ptr())
        if key in pending:
            return
        used = set(pending.values())
        free = [k for k in (0, 1) if k not in used]
        if not free:
            return
        # (the free slot may still be read by the step in flight: the copy waits for that step's
        # completion event on the side stream and then overlaps the NEXT step's compute)
        self._start_copy(free[0], img_host, gt_host)
        pending[key] = free[0]

    def step_from_host(self, img_host, img_metas, gt_host, it, deferred=False):
        """End-to-end iteration: pinned host batch -> device, step, log variables back on the host
        (what the runner's dataloader scatter + ``log_vars`` ``.item()`` do in the reference).

        ``deferred=True`` returns ``(loss, pending)`` where ``pending()`` yields the host log
        variables: the device->host copy is queued behind the step and read later, so the host
        can enqueue the next iteration instead of idling the GPU at every step boundary."""
        pending = self.__dict__.setdefault('_pending', {})
        k = pending.pop((img_host.data_ptr(), gt_host.data_ptr()), None)
        if k is None:
            used = set(pending.values())
            free = [j for j in (0, 1) if j not in used]
            if not free:                       # both slots hold prefetched batches: drop one
                dropped = next(iter(pending))
                free = [pending.pop(dropped)]
            idle = [j for j in free if j != getattr(self, '_busy_slot', None)]
            k = (idle or free)[0]
            self._start_copy(k, img_host, gt_host)
        sl = self._slot(k, img_host, gt_host)
        self._busy_slot = k
        cur = torch.cuda.current_stream()
        cur.wait_event(sl['ready'])
        loss, log_vars = self(sl['img'], img_metas, sl['gt'], it, sync=False)
        keys = list(log_vars.keys())
        # (the pinned-host allocator caches its blocks: no cudaHostAlloc after the first steps)
        host = torch.empty(len(keys), dtype=torch.float32, pin_memory=True)
        host.copy_(torch.stack([v.detach().float().reshape(()) for v in log_vars.values()]), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        sl['free'] = ev            # the slot is reusable once this step has run

        def pending_logs():
            ev.synchronize()
            return type(log_vars)(zip(keys, host.tolist()))
        if deferred:
            return loss, pending_logs
        return loss, pending_logs()
