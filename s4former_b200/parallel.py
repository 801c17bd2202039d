"""Data-parallel gradient reduction for the S4Former step.

One process per GPU; the model is replicated; each rank sees its own (sup, unsup) images.
The only data-path collective is the gradient average.  Gradients live in one flat float32
buffer (``optim.FlatGrads``) that is cut into buckets in the order the LAST backward pass
(the supervised pass, forwarded first, hence differentiated last) finishes them: heads,
encoder layers 11..0, patch embedding.  A module signals "my gradients are final" through
``_s4_grad_ready_hook`` once all of its pending backward passes (three student passes share
the weights) have run; when every parameter of a bucket is final the bucket is all-reduced
asynchronously on NCCL's stream while backward continues.
"""
import torch
import torch.distributed as dist


class PeerAllReduce:
    """One-shot all-reduce (sum) of small fp32 vectors over NVLink peer memory: SyncBatchNorm's
    per-layer statistics (``s4_peer_allreduce_f32``, csrc/peer.cu).  Collective constructor (every
    rank of ``group``, outside a stream capture); ``PeerAllReduce.get`` returns None when symmetric
    memory cannot be set up on ANY rank, and the callers fall back to ``dist.all_reduce``."""

    _cache = {}

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        n = int(lib.s4_peer_allreduce_buffer_bytes()) // 4
        self.max_elems = int(lib.s4_peer_allreduce_max_elems())
        self.buf = symm.empty(n, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.ptrs = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=device)
        self.seq = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)          # every buffer is zeroed before anybody's first call

    def all_reduce(self, t):
        from . import _lib, ops
        _lib.call('s4_peer_allreduce_f32', t.data_ptr(), t.numel(), self.ptrs.data_ptr(), self.rank, self.world,
                  self.seq.data_ptr(), ops._st())
        return t

    def usable(self, t):
        return (t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= self.max_elems
                and t.device == self.buf.device)

    @classmethod
    def get(cls, group=None):
        """The group's reducer, created on first use (collective).  S4_PEER_SYNCBN=0 disables it."""
        import os
        key = id(group) if group is not None else 0
        if key in cls._cache:
            return cls._cache[key]
        obj = None
        ok = 0
        if (os.environ.get('S4_PEER_SYNCBN', '1') != '0' and dist.get_backend(group) == 'nccl'
                and not torch.cuda.is_current_stream_capturing()):
            try:
                obj = cls(group, torch.device('cuda', torch.cuda.current_device()))
                ok = 1
            except Exception as e:       # no symmetric memory on this box / build: every rank falls back
                import warnings
                warnings.warn(f'PeerAllReduce unavailable ({type(e).__name__}: {e}); SyncBN statistics use NCCL')
                obj = None
            flag = torch.tensor([ok], dtype=torch.int32, device='cuda')
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                obj = None
        elif torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            return None                  # not cached: try again outside the capture
        cls._cache[key] = obj
        return obj


class GradReducer:
    def __init__(self, model, flat_grads, bucket_bytes=25 * 1024 * 1024, group=None):
        self.model, self.fg, self.group = model, flat_grads, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if self.world > 1 else None
        if self.world > 1 and self.backend == 'nccl':
            # NCCL's all-reduce CTAs hold SMs while backward runs: a persistent GEMM grid with a static
            # tile share per CTA then takes ~2x (the CTAs that are not resident start a whole share
            # late).  The dynamic tile scheduler lets the resident CTAs take those tiles instead.
            import os
            if os.environ.get('S4_TC_SCHED') is None:
                from . import _lib
                _lib.load().s4_set_tc_sched(1)
        # buckets over the flat buffer, walking parameters in reverse registration order
        order = list(reversed(self.fg.params))
        self.buckets = []          # [lo, hi) element ranges of the flat buffer
        self.param_bucket = {}
        lo = hi = None
        cur = []
        size = 0
        for p in order:
            off, n = self.fg.offsets[id(p)]
            cur.append(p)
            lo = off if lo is None else min(lo, off)
            hi = off + n if hi is None else max(hi, off + n)
            size += n * 4
            if size >= bucket_bytes:
                self._close(cur, lo, hi)
                cur, lo, hi, size = [], None, None, 0
        if cur:
            self._close(cur, lo, hi)
        self._owner = {}
        for mod in model.modules():
            own = [p for p in mod.parameters(recurse=False) if id(p) in self.fg.offsets]
            for p in own:
                self._owner[id(p)] = mod
        self.reset()
        self._install_hooks()

    def _close(self, params, lo, hi):
        idx = len(self.buckets)
        self.buckets.append((lo, hi, len(params)))
        for p in params:
            self.param_bucket[id(p)] = idx

    def _install_hooks(self):
        """Modules that run as a single autograd node and know when they are done."""
        for mod in self.model.modules():
            if hasattr(mod, '_s4_pending') or mod.__class__.__name__ in (
                    'TransformerEncoderLayer', 'VisionTransformer', 'SETRUPHead'):
                mod._s4_grad_ready_hook = self.module_ready

    def reset(self):
        self.ready = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.works = []

    def module_ready(self, mod):
        """All parameters directly or indirectly owned by ``mod`` have final gradients."""
        if self.world == 1:
            return
        seen = getattr(mod, '_s4_owned_params', None)
        if seen is None:
            if mod.__class__.__name__ == 'VisionTransformer':   # patch embed + cls/pos only
                seen = [mod.cls_token, mod.pos_embed] + list(mod.patch_embed.parameters())
            else:
                seen = list(mod.parameters())
            seen = [p for p in seen if id(p) in self.param_bucket]
            mod._s4_owned_params = seen
        for p in seen:
            b = self.param_bucket[id(p)]
            self.ready[b] += 1
            if self.ready[b] == self.buckets[b][2] and not self.launched[b]:
                self._launch(b)

    def _launch(self, b):
        lo, hi, _ = self.buckets[b]
        self.launched[b] = True
        buf = self.fg.flat[lo:hi]
        if self.backend == 'nccl':
            self.works.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.works.append((w, buf))

    def finalize(self):
        """Reduce whatever has not been launched, wait for everything, and re-arm."""
        if self.world > 1:
            late = [b for b in range(len(self.buckets)) if not self.launched[b]]
            self.late_buckets = len(late)
            if late and not getattr(self, '_warned_late', False):
                import warnings
                warnings.warn(f'GradReducer: {len(late)} of {len(self.buckets)} gradient buckets were not '
                              'signalled ready during backward and are all-reduced without overlap '
                              '(a student pass that fed no loss?)')
                self._warned_late = True
            for b in late:
                self._launch(b)
            for w in self.works:
                if isinstance(w, tuple):
                    w[0].wait()
                    w[1].div_(self.world)
                else:
                    w.wait()
        self.reset()


class S4DistributedDataParallel(torch.nn.Module):
    """Drop-in for mmcv's ``MMDistributedDataParallel`` as built by ``mmseg.utils.util_distribution.
    build_ddp`` (util_distribution.py:39-66, called from mmseg/apis/train.py:129-138 with
    ``device_ids=[LOCAL_RANK], broadcast_buffers=False, find_unused_parameters=...``).

    ``torch.nn.parallel.DistributedDataParallel`` cannot drive this model: the parameters are not
    autograd inputs of the fused layer nodes (weight gradients are accumulated in place by the
    wgrad kernels), so its per-parameter gradient hooks never fire.  This wrapper gives the
    runner the same contract through the library's own reducer:

      * construction broadcasts rank 0's parameters and buffers (DDP's initial sync);
      * ``train_step`` / ``val_step`` / ``forward`` delegate to the wrapped segmentor (batch tensors
        are moved to this rank's device, which is what ``MMDistributedDataParallel.scatter`` does
        for the single-device case);
      * every gradient lives in one flat buffer; buckets are all-reduced (average) as the modules
        finish their backward, and whatever is left is reduced when ``loss.backward()`` ends (an
        autograd-engine callback queued from a hook on the loss) -- so when mmcv's
        ``OptimizerHook.after_train_iter`` reaches ``optimizer.step()`` the ``.grad`` tensors hold
        the cross-rank mean, exactly as after DDP's backward.

    The runner's ``optimizer.zero_grad()`` (set_to_none or not) is honoured: gradients are re-zeroed
    here at the start of every ``train_step`` and ``ops.grad_buffer`` re-attaches a dropped
    ``.grad`` to its slice of the flat buffer."""

    def __init__(self, module, device_ids=None, dim=0, broadcast_buffers=False, find_unused_parameters=False,
                 bucket_cap_mb=25, **kwargs):
        super().__init__()
        from .optim import FlatGrads
        self.module = module
        self.dim = dim
        self.device_ids = device_ids
        self.broadcast_buffers = broadcast_buffers
        self.device = next(module.parameters()).device
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if self.world > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=0)
        self.grads = FlatGrads([p for p in module.parameters() if p.requires_grad])
        for p in self.grads.params:
            p._s4_flat_view = p.grad
        self.reducer = GradReducer(module, self.grads, bucket_bytes=int(bucket_cap_mb) * 1024 * 1024) \
            if self.world > 1 else None

    def _to_device(self, obj):
        if torch.is_tensor(obj):
            return obj.to(self.device, non_blocking=True) if obj.device != self.device else obj
        if hasattr(obj, '_data'):            # mmcv DataContainer: .data[0] is this rank's part
            data = obj._data
            return self._to_device(data[0] if isinstance(data, (list, tuple)) and len(data) == 1 else data)
        if isinstance(obj, dict):
            return {k: self._to_device(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._to_device(v) for v in obj)
        return obj

    def _arm(self, outputs):
        """Reduce the gradients when the backward pass of ``outputs['loss']`` completes."""
        loss = outputs.get('loss') if isinstance(outputs, dict) else None
        if self.reducer is None or loss is None or not loss.requires_grad:
            return outputs

        def on_backward_start(grad):
            torch.autograd.Variable._execution_engine.queue_callback(self.reducer.finalize)
            return grad
        loss.register_hook(on_backward_start)
        return outputs

    def train_step(self, *inputs, **kwargs):
        from . import ops
        self.grads.zero()
        ops.reset_arena()
        ops.reset_pending()
        if self.reducer is not None:
            self.reducer.reset()
        inputs, kwargs = self._to_device(inputs), self._to_device(kwargs)
        return self._arm(self.module.train_step(*inputs, **kwargs))

    def val_step(self, *inputs, **kwargs):
        inputs, kwargs = self._to_device(inputs), self._to_device(kwargs)
        return self.module.val_step(*inputs, **kwargs)

    def forward(self, *inputs, **kwargs):
        inputs, kwargs = self._to_device(inputs), self._to_device(kwargs)
        return self.module(*inputs, **kwargs)


def register_ddp_into_mmseg():
    """Make ``mmseg.utils.util_distribution.build_ddp`` (hence tools/train.py) build this wrapper."""
    from mmseg.utils import util_distribution
    util_distribution.ddp_factory['cuda'] = S4DistributedDataParallel
