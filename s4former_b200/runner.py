"""One training iteration as the reference's runner drives it:
``model.train_step`` (mmseg/models/segmentors/base.py:155-206) -> ``loss.backward()`` -> gradient
all-reduce -> ``optimizer.step()`` (mmcv ``OptimizerHook.after_train_iter`` + DDP reducer) with
the poly LR schedule.  ``TrainStep`` is the call a user makes per iteration; everything it does
on the device runs in the CUDA library.
"""
from collections import OrderedDict

import torch
import torch.distributed as dist

from . import configs, ops
from .optim import FusedSGD
from .parallel import GradReducer


class TrainStep:
    """``cuda_graph=True``: after ``graph_warmup`` eager iterations the WHOLE step (EMA, teacher,
    student passes, losses, backward, gradient all-reduce, SGD) is captured in one CUDA graph and
    replayed; per-step host work shrinks to the augmentation RNG draws, one small parameter upload
    (``ops.StepParams``) and one graph launch.  The eager path runs the very same device program
    (same code, same parameter buffer), so a replay is bit-identical to the eager step it replaces
    up to the order of fp32 atomics.  A batch whose shape / tag layout differs from the captured
    one falls back to the eager path."""

    def __init__(self, model, optimizer_cfg=None, lr_cfg=None, max_iters=configs.MAX_ITERS, cuda_graph=False,
                 graph_warmup=3, fused_ema=True, bucket_mb=25):
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
        self.reducer = GradReducer(model, self.optimizer.grads, bucket_bytes=int(bucket_mb * 1024 * 1024)) \
            if self.world > 1 else None
        self.device = next(model.parameters()).device
        self.params = ops.StepParams(self.device, nbytes=max(1 << 16, 8 * len(self.optimizer.params) + (1 << 14)))
        self.optimizer.attach_step_params(self.params)
        # f1: the EMA-teacher update rides in the SGD sweep (one pass over grad / momentum / weight
        # / teacher); the model then only EMAs its BatchNorm buffers at the start of a step
        self.fused_ema = bool(fused_ema) and bool(getattr(model, 'ema', False)) and hasattr(model, 'ema_pairs')
        if self.fused_ema:
            self.optimizer.attach_ema(model.ema_pairs())
        self.cuda_graph, self.graph_warmup = bool(cuda_graph), int(graph_warmup)
        self._graph = None
        self._graphs = {}
        self.max_graphs = 4          # distinct (input buffers, batch layout) signatures kept as graphs
        self._calls = 0
        self.replays = 0

    # ---- host side of a step: RNG draws (reference order) + learning rates -> one H2D copy ------
    def _prepare(self, img, img_metas, it):
        tags = [m['tag'] for m in img_metas]
        n_unsup = sum(1 for t in tags if t == 'unsup_student')
        draw = getattr(self.model, 'draw_aug_params', None)
        aug = draw(n_unsup, int(img.shape[2]), int(img.shape[3])) if draw is not None else dict(cutmix=[], perms=None)
        sp = self.params
        sp.begin()
        self.optimizer.write_lrs(it)
        staged = dict(cutmix=[sp.set(f'cutmix{i}', b) for i, b in enumerate(aug['cutmix'])], next_cutmix=0,
                      perms=aug['perms'], perms_dev=None)
        if aug['perms'] is not None:
            staged['perms_dev'] = sp.set('perms', aug['perms'])
        sp.commit()
        return staged

    def _device_step(self, img, img_metas, gt_semantic_seg, it, staged):
        """The device program of one step (eager, or under CUDA-graph capture)."""
        self.optimizer.zero_grad()
        ops.reset_arena()
        self.stale_pending = ops.reset_pending()      # 0 in a healthy step (see ops._pending_inc)
        self.model._step_aug = staged
        try:
            losses = self.model(img, img_metas, return_loss=True, gt_semantic_seg=gt_semantic_seg, iter=it)
        finally:
            self.model._step_aug = None
        loss, log_vars = self.model._parse_losses(losses, sync=False)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finalize()
        self.optimizer.step(it)
        if self.fused_ema:
            self.model._ema_fused_primed = True     # the next forward_train skips the parameter EMA
        # (detached: the step is complete, nothing differentiates through the returned loss; a
        # caller holding the previous step's autograd graph alive broke the next capture)
        return loss.detach(), log_vars

    def _signature(self, img, img_metas, gt):
        names = [m['filename'] for m in img_metas]
        return (img.data_ptr(), gt.data_ptr(), tuple(img.shape), tuple(gt.shape), img.dtype,
                tuple(m['tag'] for m in img_metas), tuple(names.index(n) for n in names), self.model.training,
                ops.compute_dtype())

    def _capture(self, img, img_metas, gt, it):
        """Capture the step reading DIRECTLY from the caller's device buffers (the graph is keyed by
        their addresses; the end-to-end path alternates two staging slots -> two graphs).  A copy
        into private static inputs would cost a 126 MB device-to-device memcpy per step, and the
        copy engines run that at a fraction of the HBM rate (measured +0.5 ms per step)."""
        staged = self._prepare(img, img_metas, it)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        steps0 = self.optimizer.steps
        from . import _lib
        l0 = _lib.load().s4_launch_count()
        # No garbage collection while capturing: a collection may free pinned staging buffers /
        # events of an older TrainStep, and the host allocator's event queries invalidate a
        # capture in progress (cudaErrorStreamCaptureInvalidated).
        import gc
        gc.collect()
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            # 'thread_local': backward runs on autograd's device thread, whose allocations may have
            # to grow the graph's private pool (cudaMalloc) -- legal, but a capture in the default
            # 'global' mode is invalidated by such a call from any other thread
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                loss, log_vars = self._device_step(img, [dict(m) for m in img_metas], gt, it, staged)
                packed = torch.stack([v.detach().float().reshape(()) for v in log_vars.values()])
        finally:
            if gc_was_enabled:
                gc.enable()
        self.graph_kernel_launches = int(_lib.load().s4_launch_count() - l0)   # library kernel nodes per replay
        self.optimizer.steps = steps0          # capture launched nothing; the replay that follows is the step
        rec = dict(graph=graph, out=(loss, list(log_vars.keys()), packed), keep=(img, gt), staged=staged)
        self._graphs[self._signature(img, img_metas, gt)] = rec
        self._graph = graph                    # (most recent capture; None = no graph yet)
        return rec

    def _replay(self, rec, img_metas, it, sync, prepared=None, img=None):
        staged = prepared if prepared is not None else self._prepare(img, img_metas, it)
        if staged['perms'] is not None:      # the reference writes the permutation into the metas
            st = [m for m in img_metas if m['tag'] == 'unsup_student']
            for m, p in zip(st, staged['perms']):
                m['PatchMixIndex'] = p
                m['PatchMix_N'] = self.model.PatchMix_N
        rec['graph'].replay()
        self.replays += 1
        self.optimizer.steps += 1
        ops.bump_all_generations()             # weights changed behind the host-side caches
        loss, keys, packed = rec['out']
        if sync:
            return loss, OrderedDict(zip(keys, packed.tolist()))
        return loss, OrderedDict(zip(keys, packed.unbind(0)))

    def __call__(self, img, img_metas, gt_semantic_seg, it, sync=True):
        """Device-resident batch -> one optimisation step.  Returns ``(loss, log_vars)``; with
        ``sync=False`` the log variables stay device tensors (no host synchronisation)."""
        self._calls += 1
        if self.cuda_graph:
            if self._graph is None:
                self._graphs = {}
            sig = self._signature(img, img_metas, gt_semantic_seg)
            rec = self._graphs.get(sig) if self._graph is not None else None
            if rec is not None:
                return self._replay(rec, img_metas, it, sync, img=img)
            if self._calls > self.graph_warmup and len(getattr(self, '_graphs', {})) < self.max_graphs:
                rec = self._capture(img, img_metas, gt_semantic_seg, it)
                # the capture consumed this step's RNG draws and staged them: replay with those
                return self._replay(rec, img_metas, it, sync, prepared=rec['staged'])
        staged = self._prepare(img, img_metas, it)
        loss, log_vars = self._device_step(img, img_metas, gt_semantic_seg, it, staged)
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
