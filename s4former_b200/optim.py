"""Flat parameter/gradient storage, the fused SGD-momentum step and the poly LR schedule
(the step either side of the hot path).

Reference behaviour reproduced: ``torch.optim.SGD(lr=1e-3, momentum=0.9, weight_decay=0)`` with
mmcv's ``DefaultOptimizerConstructor`` ``custom_keys={'head': dict(lr_mult=10.)}``
(configs/setr/*_MT_w_ours.py:259-262) and the ``poly`` policy (power 0.9, min_lr 1e-4,
by_epoch=False; configs/_base_/schedules/schedule_80k_pascal_1over8.py:2-5).
"""
import torch

from . import ops


class FlatGrads:
    """All gradients of ``params`` as views of ONE contiguous float32 buffer: zeroing is one
    memset, the data-parallel reducer all-reduces slices of it, the optimizer reads it in one
    multi-tensor launch."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets = {}
        off = 0
        for p in self.params:
            self.offsets[id(p)] = (off, p.numel())
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()
        for p in self.params:            # a foreign zero_grad(set_to_none=True) would detach the views
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + self.offsets[id(p)][0] * 4:
                off, n = self.offsets[id(p)]
                p.grad = self.flat[off:off + n].view_as(p)


def poly_lr(base_lr, it, max_iters, power=0.9, min_lr=1e-4):
    """mmcv PolyLrUpdaterHook: (base - min) * (1 - it/max)^power + min."""
    coeff = (1 - it / max_iters) ** power
    return (base_lr - min_lr) * coeff + min_lr


class FusedSGD:
    """SGD with momentum over named parameters in one multi-tensor kernel launch."""

    def __init__(self, named_params, lr=1e-3, momentum=0.9, weight_decay=0.0, custom_keys=None,
                 max_iters=80000, power=0.9, min_lr=1e-4):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.base_lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.max_iters, self.power, self.min_lr = max_iters, power, min_lr
        custom_keys = custom_keys or {}
        self.lr_mult = []
        for n in self.names:
            mult = 1.0
            for key in sorted(custom_keys, key=len, reverse=True):   # mmcv: longest key first
                if key in n:
                    mult = custom_keys[key].get('lr_mult', 1.0)
                    break
            self.lr_mult.append(mult)
        self.grads = FlatGrads(self.params)
        self.bufs = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.steps = 0
        self._table = None
        self._shadow_key = None

    def zero_grad(self):
        self.grads.zero()

    def current_lrs(self, it=None):
        it = self.steps if it is None else it
        # mmcv applies the schedule to each group's own initial lr (base*mult), min_lr is absolute
        return [poly_lr(self.base_lr * m, it, self.max_iters, self.power, self.min_lr) for m in self.lr_mult]

    def attach_step_params(self, sp):
        """Read the per-step learning rates from ``sp`` (ops.StepParams, field 'lrs') instead of
        copying a fresh host table inside ``step``: the step's device program then has no host
        dependency of its own (CUDA-graph capture)."""
        self._sp = sp
        self._sp_view = sp.view('lrs', (len(self.params),), torch.float32)
        self._table = None

    def write_lrs(self, it=None):
        """Stage this step's learning rates (call between ``sp.begin()`` and ``sp.commit()``)."""
        self._sp.set('lrs', torch.tensor(self.current_lrs(it), dtype=torch.float32))

    def attach_ema(self, pairs):
        """``pairs``: {id(student_param): (teacher_param, ema_momentum)}.  From now on ``step``
        also applies teacher <- m teacher + (1-m) student for those parameters in the SAME sweep
        (the reference does it at the start of the next iteration on exactly these values:
        encoder_decoder.py:416-423, 1044-1066)."""
        self._ema_pairs = dict(pairs)
        self._table = None

    def step(self, it=None):
        sp_view = getattr(self, '_sp_view', None)
        pairs = getattr(self, '_ema_pairs', None)
        shadows = ops.shadow_list(self.params)      # bf16 weight copies the GEMMs read
        ema = [pairs.get(id(p), (None, 0.0))[0] for p in self.params] if pairs else None
        ema_sh = ops.shadow_list([e for e in ema]) if pairs else []
        skey = tuple(0 if s is None else s.data_ptr() for s in list(shadows) + list(ema_sh))
        if self._table is None or self._shadow_key != skey:
            dev = self.params[0].device
            lists = [[p.data for p in self.params], [p.grad for p in self.params], self.bufs, shadows]
            if pairs:
                lists += [[None if e is None else e.data for e in ema], ema_sh]
            self._table = ops.TensorTable(lists, dev, lrs=sp_view if sp_view is not None else self.current_lrs(it))
            self._table.targets = self.params
            if pairs:
                self._table.ema_m = torch.tensor([pairs.get(id(p), (None, 0.0))[1] for p in self.params],
                                                 dtype=torch.float32).to(dev)
                self._ema_targets = [e for e in ema if e is not None]
            self._shadow_key = skey
        fn = ops.sgd_ema_step if pairs else ops.sgd_step
        fn(self._table, self.momentum, self.weight_decay, first_step=(self.steps == 0),
           lrs=None if sp_view is not None else self.current_lrs(it))
        for p in self.params:
            ops.bump_generation(p)                  # repacked conv weights are rebuilt lazily
        ops.mark_shadows_fresh(self.params)         # ... the bf16 shadows were refreshed in the same pass
        if pairs:
            ops.mark_shadows_fresh(self._ema_targets)
        self.steps += 1
