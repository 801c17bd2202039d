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


class GradReducer:
    def __init__(self, model, flat_grads, bucket_bytes=25 * 1024 * 1024, group=None):
        self.model, self.fg, self.group = model, flat_grads, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.backend = dist.get_backend(group) if self.world > 1 else None
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
