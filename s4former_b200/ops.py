"""Host-side operator layer: thin Python wrappers and ``torch.autograd.Function``s over the
C ABI (``include/s4former.h``).  PyTorch is used for device memory, streams and autograd
bookkeeping only; every arithmetic step of the path runs in the CUDA library.

Gradient convention: weight/bias gradients are ACCUMULATED IN PLACE into ``param.grad``
(float32, allocated as zeros on first use) by the wgrad kernels; ``Function.backward`` returns
``None`` for them.  This removes one read-modify-write pass per parameter per student pass
(three passes share the weights in the S4Former step) and lets the data-parallel reducer
consume a contiguous gradient buffer.
"""
import ctypes as C
import math

import torch

from . import _lib as L

_cfg = {'compute_dtype': torch.bfloat16, 'backend': L.BACKEND_AUTO, 'wgrad_split_k': 0}


def set_compute_dtype(dtype):
    """torch.bfloat16 (speed mode, tcgen05) or torch.float32 (validation mode, CUDA cores)."""
    assert dtype in (torch.bfloat16, torch.float32)
    _cfg['compute_dtype'] = dtype


def compute_dtype():
    return _cfg['compute_dtype']


def set_backend(backend):
    _cfg['backend'] = {'auto': L.BACKEND_AUTO, 'simt': L.BACKEND_SIMT, 'tc': L.BACKEND_TC}.get(backend, backend)


def backend():
    return _cfg['backend']


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _st():
    # raw handle of the current stream (torch.cuda.current_stream() builds a Stream object and
    # re-resolves the device on every call: a quarter of the step's host time at ~3000 calls)
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _code(dtype):
    return L.BF16 if dtype == torch.bfloat16 else L.F32


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.S4Error('s4former_b200 ops need CUDA tensors: there is no CPU fallback')


# ----------------------------------------------------------------------------------------------
# low-precision / repacked weight cache
# ----------------------------------------------------------------------------------------------
_global_gen = [0]


def bump_generation(p):
    """Call after a raw-pointer kernel (EMA / SGD) has modified ``p`` in place."""
    p._s4_gen = getattr(p, '_s4_gen', 0) + 1


def bump_all_generations():
    """Every parameter may have changed behind the host's back (a CUDA-graph replay of the whole
    step): one counter that is part of every cache tag."""
    _global_gen[0] += 1


def _cache(p, key, make):
    tag = (p._version, getattr(p, '_s4_gen', 0), _global_gen[0], p.data_ptr(), compute_dtype())
    store = p.__dict__.setdefault('_s4_cache', {})
    hit = store.get(key)
    if hit is not None and hit[0] == tag:
        return hit[1]
    val = make()
    store[key] = (tag, val)
    return val


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    _require_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x, dtype=dtype)
    L.call('s4_cast', _p(x), _p(y), x.numel(), _code(x.dtype), _code(dtype), _st())
    return y


def transpose2d(x):
    """[R, C] -> [C, R] contiguous."""
    r, c = x.shape
    y = torch.empty((c, r), dtype=x.dtype, device=x.device)
    L.call('s4_transpose', _p(x), _p(y), 1, r, c, _code(x.dtype), _st())
    return y


def _param_tag(p):
    return (p._version, getattr(p, '_s4_gen', 0), p.data_ptr())     # (shadows are refreshed by the
    # EMA / SGD kernels themselves, also inside a graph replay: no global generation here)


def lowp(p):
    """The weight in the compute dtype ([out, in] as stored).

    bf16: a persistent SHADOW buffer per parameter (stable address).  The fused SGD / EMA kernels
    refresh it in the pass that updates the fp32 master (``shadow_ptrs`` / ``mark_shadows_fresh``),
    so no per-step cast launches remain; any other in-place change of the parameter is detected
    through its version tag and re-cast here."""
    w2d = p.detach().reshape(p.shape[0], -1)
    if compute_dtype() == torch.float32:
        return w2d
    _require_cuda(p)
    sh = getattr(p, '_s4_shadow', None)
    tag = _param_tag(p)
    if sh is not None and getattr(p, '_s4_shadow_tag', None) == tag:
        return sh
    if sh is None:
        sh = torch.empty(w2d.shape, dtype=torch.bfloat16, device=p.device)
        p._s4_shadow = sh
    src = w2d.contiguous()
    L.call('s4_cast', _p(src), _p(sh), src.numel(), _code(src.dtype), L.BF16, _st())
    p._s4_shadow_tag = tag
    return sh


def shadow_list(params):
    """bf16 shadow buffers (or None) of ``params``, for the multi-tensor kernels."""
    return [None if p is None else getattr(p, '_s4_shadow', None) for p in params]


def mark_shadows_fresh(params):
    """Call after a kernel that rewrote both the parameters and their shadows."""
    for p in params:
        if getattr(p, '_s4_shadow', None) is not None:
            p._s4_shadow_tag = _param_tag(p)


def conv_packed(p):
    """(fwd [Cout, 9*Cin], dgrad [Cin, 9*Cout]) repacks of a [Cout, Cin, 3, 3] weight."""
    def make():
        cout, cin = p.shape[0], p.shape[1]
        dt = compute_dtype()
        wf = torch.empty((cout, 9 * cin), dtype=dt, device=p.device)
        wd = torch.empty((cin, 9 * cout), dtype=dt, device=p.device)
        L.call('s4_pack_conv3x3_weight', _p(p.detach()), _p(wf), _p(wd), cin, cout, _code(dt), _st())
        return wf, wd
    return _cache(p, 'conv', make)


# ----------------------------------------------------------------------------------------------
# host-drawn per-step parameters (augmentation boxes / permutations, learning rates)
# ----------------------------------------------------------------------------------------------
class StepParams:
    """Everything the HOST decides per training step (CutMix boxes, PatchShuffle permutations --
    host RNG, reference call order -- and the poly-LR table) lives in ONE persistent device buffer
    refreshed by ONE small host->device copy before the step's first kernel.  The device program
    of a step then depends on the host only through this buffer, which is what allows the whole
    step to be captured in a CUDA graph and replayed (runner.TrainStep).

    The staging side is a ring of pinned slots: a slot is rewritten only after the copy that read
    it has completed on the device (the host may run a step or two ahead of the GPU)."""

    def __init__(self, device, nbytes=1 << 16, slots=4):
        self.device = torch.device(device)
        self.dev = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.host = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.k = -1
        self.fields = {}
        self.off = 0

    def _field(self, name, shape, dtype):
        f = self.fields.get(name)
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if f is None or f[1] != tuple(shape) or f[2] != dtype:
            off = (self.off + 15) // 16 * 16
            if off + nbytes > self.dev.numel():
                raise RuntimeError('StepParams buffer exhausted')
            self.off = off + nbytes
            f = (off, tuple(shape), dtype, nbytes)
            self.fields[name] = f
        return f

    def view(self, name, shape=None, dtype=None):
        """Device view of a field (stable address for the lifetime of this object)."""
        off, shp, dt, nbytes = self.fields[name] if shape is None else self._field(name, shape, dtype)
        return self.dev[off:off + nbytes].view(dt).view(shp)

    def begin(self):
        self.k = (self.k + 1) % len(self.host)
        ev = self.events[self.k]
        if ev is not None:
            ev.synchronize()

    def set(self, name, value):
        """Write a CPU tensor into the current staging slot; returns the device view."""
        value = value.contiguous()
        off, shp, dt, nbytes = self._field(name, value.shape, value.dtype)
        self.host[self.k][off:off + nbytes].view(dt).view(shp).copy_(value)
        return self.dev[off:off + nbytes].view(dt).view(shp)

    def commit(self):
        """One async H2D copy of the used prefix on the current stream."""
        n = (self.off + 15) // 16 * 16
        if n:
            self.dev[:n].copy_(self.host[self.k][:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[self.k] = ev


# ----------------------------------------------------------------------------------------------
# per-step arena of zeroed float32 scratch (BatchNorm / loss partial sums): one memset per step
# instead of one fill launch per use
# ----------------------------------------------------------------------------------------------
_arena = {}
_ARENA_CHUNK = 1 << 18


def arena_zeros(shape, device):
    """A zero-initialised float32 tensor carved from the device's step arena.  Valid until the
    arena is reset at the start of the next training step (``reset_arena``); without resets the
    arena simply keeps allocating fresh zeroed chunks."""
    n = 1
    for d in shape:
        n *= int(d)
    a = _arena.get(device)
    if a is None or a['off'] + n > a['buf'].numel():
        a = dict(buf=torch.zeros(max(n, _ARENA_CHUNK), dtype=torch.float32, device=device), off=0)
        _arena[device] = a
    t = a['buf'][a['off']:a['off'] + n].view(shape)
    a['off'] += (n + 3) // 4 * 4
    return t


def reset_arena():
    """Re-zero the part of every arena used since the last reset (stream-ordered: everything that
    read the old contents was enqueued before this call) and start carving from the top again."""
    for a in _arena.values():
        if a['off']:
            a['buf'][:a['off']].zero_()
            a['off'] = 0


def _add_bn_grads(bn, dgamma_dbeta):
    """bn.weight.grad += dgamma_dbeta[0]; bn.bias.grad += dgamma_dbeta[1] -- one launch when the
    two gradients are adjacent in the flat gradient buffer (weight, then bias: registration order)."""
    gw, gb = grad_buffer(bn.weight), grad_buffer(bn.bias)
    Cc = gw.numel()
    if (gb.data_ptr() == gw.data_ptr() + 4 * Cc and gw.is_contiguous() and gb.is_contiguous()
            and gw.untyped_storage().data_ptr() == gb.untyped_storage().data_ptr()):
        gw.as_strided((2, Cc), (Cc, 1)).add_(dgamma_dbeta)
    else:
        gw.add_(dgamma_dbeta[0])
        gb.add_(dgamma_dbeta[1])


# Modules that run as ONE autograd node count their pending backward passes (three student passes
# share the weights); the data-parallel reducer is told when the count returns to zero.  The
# counters are re-armed at the start of every step (``reset_pending``): a grad-enabled forward
# whose output feeds no loss, or an exception in the middle of backward, must not leave them
# above zero for good (every bucket would silently fall back to ``GradReducer.finalize``).
_pending_mods = {}


def _pending_inc(mod):
    mod._s4_pending = getattr(mod, '_s4_pending', 0) + 1
    _pending_mods[id(mod)] = mod


def _pending_dec(mod):
    mod._s4_pending -= 1
    if mod._s4_pending == 0:
        hook = getattr(mod, '_s4_grad_ready_hook', None)
        if hook is not None:
            hook(mod)


def reset_pending():
    """Returns how many modules still had a pending count (0 in a healthy step)."""
    stale = 0
    for mod in _pending_mods.values():
        if getattr(mod, '_s4_pending', 0):
            stale += 1
            mod._s4_pending = 0
    return stale


_frozen_scratch = {}


def grad_buffer(p):
    """The tensor a kernel accumulates ``p``'s gradient into.  A frozen parameter
    (``requires_grad=False``) gets a shared scratch buffer nobody reads — its ``.grad`` stays None,
    as autograd would leave it; the weight-gradient GEMMs of frozen weights are skipped altogether."""
    if not p.requires_grad:
        key = (tuple(p.shape), p.device)
        buf = _frozen_scratch.get(key)
        if buf is None:
            buf = _frozen_scratch[key] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
        return buf
    if p.grad is None:
        flat = getattr(p, '_s4_flat_view', None)     # a foreign zero_grad(set_to_none=True) dropped it:
        if flat is not None:                         # re-attach the slice of the flat gradient buffer
            p.grad = flat                            # (zeroed at the start of the step by its owner)
        else:
            p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


# ----------------------------------------------------------------------------------------------
# GEMM
# ----------------------------------------------------------------------------------------------
def gemm(a, b, c, M, N, K, a_str, b_str, c_sm, batch=(1, 1), a_bs=(0, 0), b_bs=(0, 0), c_bs=(0, 0),
         bias=None, aux=None, res=None, pre=None, alpha=1.0, act=L.ACT_NONE, accumulate=False,
         split_k=1, dtype=None, backend_override=None, colsum=None):
    """C = epi(alpha * A @ B) with element strides a_str=(sm, sk), b_str=(sk, sn).

    ``colsum`` (float32 [N]): the epilogue also accumulates the column sums of C into it when the
    problem runs on the tcgen05 path; returns True if it did (False: the caller sums C itself)."""
    g = L.GemmParams()
    g.a, g.b, g.c = _p(a), _p(b), _p(c)
    g.bias, g.aux, g.res, g.pre = _p(bias), _p(aux), _p(res), _p(pre)
    g.M, g.N, g.K = M, N, K
    g.nb1, g.nb2 = batch
    g.a_sm, g.a_sk = a_str
    g.a_b1, g.a_b2 = a_bs
    g.b_sk, g.b_sn = b_str
    g.b_b1, g.b_b2 = b_bs
    g.c_sm = c_sm
    g.c_b1, g.c_b2 = c_bs
    g.alpha, g.act, g.accumulate = alpha, act, int(accumulate)
    g.dtype = _code(a.dtype if dtype is None else dtype)
    g.c_dtype = _code(c.dtype)
    g.backend = backend() if backend_override is None else backend_override
    g.split_k = split_k
    fused = False
    if colsum is not None:
        g.colsum = None
        if L.load().s4_gemm_uses_tc(C.byref(g)):
            g.colsum = _p(colsum)
            fused = True
    L.check(L.load().s4_gemm(C.byref(g), _st()), 's4_gemm')
    return fused


def _wgrad_split(M_out, N_out, K_red):
    """split-K factor for weight gradients: -1 = chosen by the library's tile cost model together
    with the tile shape (whole waves of CTA groups)."""
    if _cfg['wgrad_split_k']:
        return _cfg['wgrad_split_k']
    return -1


def linear_fwd(x, w_lp, bias, act=L.ACT_NONE, res=None, want_pre=False):
    """y = act(x @ w^T + bias) + res ; x [M,K], w_lp [N,K]."""
    M, K = x.shape
    N = w_lp.shape[0]
    y = torch.empty((M, N), dtype=x.dtype, device=x.device)
    pre = torch.empty_like(y) if want_pre else None
    gemm(x, w_lp, y, M, N, K, (K, 1), (1, K), N, bias=bias, res=res, pre=pre, act=act)
    return (y, pre) if want_pre else y


def linear_dgrad(dy, w_lp, aux=None, colsum_param=None):
    """dx = (dy @ w) * gelu'(aux) ; dy [M,N], w_lp [N,K] as stored: the weight is the MN-major B
    operand of the GEMM, so no transposed copy is ever made.

    ``colsum_param``: a bias parameter whose gradient is colsum(dx) (dx is the dY of the linear
    layer below): accumulated from the GEMM epilogue when possible, by ``s4_colsum`` otherwise."""
    M, N = dy.shape
    K = w_lp.shape[1]
    dx = torch.empty((M, K), dtype=dy.dtype, device=dy.device)
    gb = grad_buffer(colsum_param) if colsum_param is not None else None
    fused = gemm(dy, w_lp, dx, M, K, N, (N, 1), (K, 1), K, aux=aux, colsum=gb)
    if gb is not None and not fused:
        L.call('s4_colsum', _p(dx), _p(gb), None, M, K, _code(dx.dtype), _st())
    return dx


def linear_wgrad(dy, x, w_param, b_param):
    """w.grad [N,K] += dy^T x ; b.grad [N] += colsum(dy)."""
    M, N = dy.shape
    K = x.shape[1]
    if w_param.requires_grad:
        gw = grad_buffer(w_param)
        gemm(dy, x, gw, N, K, M, (1, N), (K, 1), K, accumulate=True, split_k=_wgrad_split(N, K, M))
    if b_param is not None and b_param.requires_grad:
        gb = grad_buffer(b_param)
        L.call('s4_colsum', _p(dy), _p(gb), None, M, N, _code(dy.dtype), _st())


# ----------------------------------------------------------------------------------------------
# LayerNorm
# ----------------------------------------------------------------------------------------------
def layernorm_fwd(x2d, gamma, beta, eps, row_map=None, out_rows=None):
    D = x2d.shape[1]
    rows = x2d.shape[0] if out_rows is None else out_rows
    y = torch.empty((rows, D), dtype=x2d.dtype, device=x2d.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x2d.device)
    rstd = torch.empty_like(mean)
    L.call('s4_layernorm_fwd', _p(x2d), _p(row_map), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd),
           rows, D, eps, _code(x2d.dtype), _st())
    return y, mean, rstd


def layernorm_bwd(dy, x2d, gamma_p, beta_p, mean, rstd, row_map=None, dres=None, dx=None,
                  dres_bias=None):
    """``dres_bias``: a bias parameter whose gradient is colsum(dres) (the linear layer whose output
    joined the residual stream here); accumulated by the same pass."""
    rows, D = dy.shape
    if dx is None:
        dx = torch.zeros_like(x2d) if row_map is not None else torch.empty_like(x2d)
    L.call('s4_layernorm_bwd', _p(dy), _p(x2d), _p(row_map), _p(gamma_p.detach()), _p(mean), _p(rstd),
           _p(dres), _p(dx), _p(grad_buffer(gamma_p)), _p(grad_buffer(beta_p)),
           _p(grad_buffer(dres_bias)) if dres_bias is not None else None, rows, D,
           _code(dy.dtype), _st())
    return dx


# ----------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------
_ws_cache = {}


def workspace(nbytes, device, tag='ws'):
    """A growing scratch buffer per (device, stream, tag); contents are never kept across ops."""
    key = (device, torch.cuda.current_stream().cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def attention_fwd(qkv, B, L_, H, hd, u0, gate, w):
    out = torch.empty((B * L_, H * hd), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((B, H, L_), dtype=torch.float32, device=qkv.device)
    dt = _code(qkv.dtype)
    nbytes = L.load().s4_attention_workspace(B, H, L_, hd, dt, backend())
    ws = workspace(nbytes, qkv.device, 'attn')
    L.call('s4_attention_fwd', _p(qkv), _p(u0), _p(gate), float(w), _p(out), _p(lse), _p(ws), nbytes,
           B, H, L_, hd, dt, backend(), _st())
    return out, lse


def attention_bwd(dout, qkv, out, lse, B, L_, H, hd, u0, gate, w):
    dqkv = torch.empty_like(qkv)
    dt = _code(qkv.dtype)
    nbytes = L.load().s4_attention_workspace(B, H, L_, hd, dt, backend())
    ws = workspace(nbytes, qkv.device, 'attn')
    L.call('s4_attention_bwd', _p(dout), _p(qkv), _p(out), _p(lse), _p(u0), _p(gate), float(w),
           _p(dqkv), _p(ws), nbytes, B, H, L_, hd, dt, backend(), _st())
    return dqkv


# ----------------------------------------------------------------------------------------------
# transformer encoder layer (reference vit.py:113-127) as ONE autograd node
# ----------------------------------------------------------------------------------------------
class EncoderLayerFn(torch.autograd.Function):
    """x + OutProj(Attn(LN1(x))) then + FFN(LN2(.)); ``layer`` supplies the parameters."""

    @staticmethod
    def forward(ctx, x, layer, B, Ltok, u0, gate, w, _param_probe=None):
        # ``_param_probe`` (one of the layer's parameters) makes ``needs_input_grad`` reflect the
        # parameters too: a trainable layer behind a frozen / non-differentiable input still
        # takes the training path (its weights are not autograd inputs, see the module docstring)
        M, D = x.shape
        H = layer.num_heads
        hd = D // H
        mha = layer.attn.attn
        fc1, fc2 = layer.ffn.layers[0][0], layer.ffn.layers[1]
        eps = layer.ln1.eps
        xl1, mean1, rstd1 = layernorm_fwd(x, layer.ln1.weight, layer.ln1.bias, eps)
        qkv = linear_fwd(xl1, lowp(mha.in_proj_weight), mha.in_proj_bias)
        att, lse = attention_fwd(qkv, B, Ltok, H, hd, u0, gate, w)
        xm = linear_fwd(att, lowp(mha.out_proj.weight), mha.out_proj.bias, res=x)
        xl2, mean2, rstd2 = layernorm_fwd(xm, layer.ln2.weight, layer.ln2.bias, eps)
        if not any(ctx.needs_input_grad):
            # no-grad pass (the EMA teacher): the pre-activation copy (one extra [M, 4D] store) and
            # the saved activations are only needed by backward
            h = linear_fwd(xl2, lowp(fc1.weight), fc1.bias, act=L.ACT_GELU)
            return linear_fwd(h, lowp(fc2.weight), fc2.bias, res=xm)
        h, pre = linear_fwd(xl2, lowp(fc1.weight), fc1.bias, act=L.ACT_GELU, want_pre=True)
        y = linear_fwd(h, lowp(fc2.weight), fc2.bias, res=xm)
        ctx.layer, ctx.dims, ctx.w = layer, (B, Ltok, H, hd), w
        ctx.save_for_backward(x, xl1, mean1, rstd1, qkv, att, lse, xm, xl2, mean2, rstd2, pre, h, u0, gate)
        _pending_inc(layer)
        return y

    @staticmethod
    def backward(ctx, dy):
        layer = ctx.layer
        (x, xl1, mean1, rstd1, qkv, att, lse, xm, xl2, mean2, rstd2, pre, h, u0, gate) = ctx.saved_tensors
        B, Ltok, H, hd = ctx.dims
        mha = layer.attn.attn
        fc1, fc2 = layer.ffn.layers[0][0], layer.ffn.layers[1]
        dy = dy.contiguous()
        # FFN
        # (dy W2) * gelu'(pre); the same epilogue accumulates fc1.bias.grad = colsum(dpre)
        dpre = linear_dgrad(dy, lowp(fc2.weight), aux=pre, colsum_param=fc1.bias)
        linear_wgrad(dy, h, fc2.weight, None)
        dxl2 = linear_dgrad(dpre, lowp(fc1.weight))
        linear_wgrad(dpre, xl2, fc1.weight, None)
        # fc2.bias.grad = colsum(dy) and out_proj.bias.grad = colsum(dxm) come out of the
        # LayerNorm-backward passes that read dy / dxm as their residual gradient
        dxm = layernorm_bwd(dxl2, xm, layer.ln2.weight, layer.ln2.bias, mean2, rstd2, dres=dy,
                            dres_bias=fc2.bias)
        # attention block
        datt = linear_dgrad(dxm, lowp(mha.out_proj.weight))
        linear_wgrad(dxm, att, mha.out_proj.weight, None)
        dqkv = attention_bwd(datt, qkv, att, lse, B, Ltok, H, hd, u0, gate, ctx.w)
        dxl1 = linear_dgrad(dqkv, lowp(mha.in_proj_weight))
        linear_wgrad(dqkv, xl1, mha.in_proj_weight, mha.in_proj_bias)
        dx = layernorm_bwd(dxl1, x, layer.ln1.weight, layer.ln1.bias, mean1, rstd1, dres=dxm,
                           dres_bias=mha.out_proj.bias)
        _pending_dec(layer)
        return dx, None, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# patch embedding + cls/pos assembly (reference embed.py:183-204, vit.py:483-513)
# ----------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_token, bb, img, pos_override=None):
        _require_cuda(img)
        B, Cin, Himg, Wimg = img.shape
        P = bb.patch_size
        gh, gw = (Himg + P - 1) // P, (Wimg + P - 1) // P
        proj = bb.patch_embed.projection
        D = proj.weight.shape[0]
        dt = compute_dtype()
        a = torch.empty((B * gh * gw, Cin * P * P), dtype=dt, device=img.device)
        L.call('s4_patchify', _p(img.contiguous()), _p(a), B, Cin, Himg, Wimg, P, _code(dt), _st())
        tok = linear_fwd(a, lowp(proj.weight), proj.bias)
        Ltok = gh * gw + 1
        x = torch.empty((B * Ltok, D), dtype=dt, device=img.device)
        pos = bb.pos_embed.detach() if pos_override is None else pos_override
        L.call('s4_assemble_tokens', _p(tok), _p(cls_token.detach()), _p(pos), _p(x),
               B, Ltok, D, _code(dt), _st())
        ctx.bb, ctx.dims = bb, (B, Ltok, D)
        ctx.save_for_backward(a)
        if any(ctx.needs_input_grad):       # all False under no_grad: no backward will come
            _pending_inc(bb)
        return x

    @staticmethod
    def backward(ctx, dx):
        bb = ctx.bb
        (a,) = ctx.saved_tensors
        B, Ltok, D = ctx.dims
        dx = dx.contiguous()
        dtok = torch.empty((B * (Ltok - 1), D), dtype=dx.dtype, device=dx.device)
        # cls_token / pos_embed gradients are accumulated straight into their gradient buffers
        # (like every other parameter): returning the cls gradient to autograd instead would let
        # AccumulateGrad add it AFTER the grad-ready hook below has handed the bucket holding
        # cls_token to the asynchronous all-reduce (a data race at N > 1).
        L.call('s4_assemble_tokens_bwd', _p(dx), _p(dtok), _p(grad_buffer(bb.cls_token)),
               _p(grad_buffer(bb.pos_embed)), B, Ltok, D, _code(dx.dtype), _st())
        proj = bb.patch_embed.projection
        linear_wgrad(dtok, a, proj.weight, proj.bias)
        _pending_dec(bb)
        return None, None, None, None


# ----------------------------------------------------------------------------------------------
# SETR-PUP head stages
# ----------------------------------------------------------------------------------------------
def _all_reduce_stats(t, group_info):
    """SyncBN: sum the per-rank statistics over the data-parallel group (reference N2)."""
    if group_info is not None and group_info.get('world', 1) > 1:
        import torch.distributed as dist
        peer = group_info.get('peer')
        if peer is not None and peer.usable(t):
            return peer.all_reduce(t)        # one-shot sum over NVLink peer memory (csrc/peer.cu)
        dist.all_reduce(t, group=group_info.get('group'))
    return t


class HeadLNFn(torch.autograd.Function):
    """Feature tap (drop cls) + PatchMix un-shuffle + LayerNorm, as a row-gathered LN
    (reference setr_up_head.py:96-104, decode_head.py:186-212)."""

    @staticmethod
    def forward(ctx, x_tokens, head, row_map, B, Ltok, _param_probe=None):
        D = x_tokens.shape[1]
        rows = B * (Ltok - 1)
        y, mean, rstd = layernorm_fwd(x_tokens, head.norm.weight, head.norm.bias, head.norm.eps,
                                      row_map=row_map, out_rows=rows)
        ctx.head = head
        ctx.save_for_backward(x_tokens, mean, rstd, row_map)
        if any(ctx.needs_input_grad):
            _pending_inc(head)
        return y

    @staticmethod
    def backward(ctx, dy):
        x_tokens, mean, rstd, row_map = ctx.saved_tensors
        head = ctx.head
        dx = layernorm_bwd(dy.contiguous(), x_tokens, head.norm.weight, head.norm.bias, mean, rstd,
                           row_map=row_map)
        _pending_dec(head)                 # the LayerNorm is the last node of the head's backward
        return dx, None, None, None, None, None


def _conv_fwd(stage, x, B, H, W, Cin, Cout, training):
    """conv3x3 forward; in training mode the BatchNorm batch statistics [sum, sumsq] of its output
    come out of the same launch (GEMM epilogue).  Returns (y, stats or None)."""
    wf, _ = conv_packed(stage.conv.weight)
    dt = _code(x.dtype)
    y = torch.empty((B * H * W, Cout), dtype=x.dtype, device=x.device)
    if not training:
        L.call('s4_conv3x3_fwd', _p(x), _p(wf), _p(y), B, H, W, Cin, Cout, dt, backend(), _st())
        return y, None
    stats = arena_zeros((2, Cout), x.device)
    L.call('s4_conv3x3_fwd_stats', _p(x), _p(wf), _p(y), _p(stats[0]), _p(stats[1]), B, H, W, Cin, Cout,
           dt, backend(), _st())
    return y, stats


def _bn_scale_shift(conv_bn, y2d, rows, training, group_info, stats=None):
    """BatchNorm statistics of the conv output (training: batch stats, all-reduced for SyncBN;
    eval: running stats).  Returns (scale, shift, mean, invstd, count)."""
    bn = conv_bn.bn
    Cc = y2d.shape[1]
    dev = y2d.device
    scale = torch.empty(Cc, dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    if not training:
        L.call('s4_bn_eval_affine', _p(bn.running_mean), _p(bn.running_var), _p(bn.weight.detach()),
               _p(bn.bias.detach()), bn.eps, _p(scale), _p(shift), Cc, _st())
        return scale, shift, None, None, 0.0
    if stats is None:
        stats = arena_zeros((2, Cc), dev)
        L.call('s4_colsum', _p(y2d), _p(stats[0]), _p(stats[1]), rows, Cc, _code(y2d.dtype), _st())
    count = float(rows)
    mean = torch.empty_like(scale)
    invstd = torch.empty_like(scale)
    mom = bn.momentum if bn.momentum is not None else 0.1
    if group_info is not None and group_info.get('world', 1) > 1:
        count *= group_info['world']
        peer = group_info.get('peer')
        if peer is not None and peer.usable(stats):
            # SyncBN forward in one kernel: exchange the local sums over NVLink peer memory, add them in
            # rank order, finalize (csrc/peer.cu)
            L.call('s4_bn_finalize_peer', _p(stats), count, bn.eps, mom, _p(bn.weight.detach()),
                   _p(bn.bias.detach()), _p(mean), _p(invstd), _p(scale), _p(shift), _p(bn.running_mean),
                   _p(bn.running_var), _p(bn.num_batches_tracked), Cc, _p(peer.ptrs), peer.rank, peer.world,
                   _p(peer.seq), _st())
            return scale, shift, mean, invstd, count
        _all_reduce_stats(stats, group_info)
    L.call('s4_bn_finalize', _p(stats[0]), _p(stats[1]), count, bn.eps, mom, _p(bn.weight.detach()),
           _p(bn.bias.detach()), _p(mean), _p(invstd), _p(scale), _p(shift), _p(bn.running_mean),
           _p(bn.running_var), _p(bn.num_batches_tracked), Cc, _st())
    return scale, shift, mean, invstd, count


class ConvBNReLUUpFn(torch.autograd.Function):
    """conv3x3(no bias) -> BN -> ReLU -> bilinear x s  (one SETR-PUP stage, NHWC)."""

    @staticmethod
    def forward(ctx, x, stage, B, H, W, s, training, group_info):
        conv = stage.conv
        Cout, Cin = conv.weight.shape[0], conv.weight.shape[1]
        dt = _code(x.dtype)
        y, stats = _conv_fwd(stage, x, B, H, W, Cin, Cout, training)
        scale, shift, mean, invstd, count = _bn_scale_shift(stage, y, B * H * W, training, group_info, stats)
        out = torch.empty((B * H * s * W * s, Cout), dtype=x.dtype, device=x.device)
        L.call('s4_bn_relu_upsample_fwd', _p(y), _p(scale), _p(shift), _p(out), B, H, W, Cout, s, dt, _st())
        ctx.stage, ctx.dims, ctx.group_info, ctx.count = stage, (B, H, W, s, Cin, Cout), group_info, count
        if training:
            ctx.save_for_backward(x, y, scale, shift, mean, invstd)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, scale, shift, mean, invstd = ctx.saved_tensors
        stage = ctx.stage
        B, H, W, s, Cin, Cout = ctx.dims
        dt = _code(x.dtype)
        dout = dout.contiguous()
        dact = torch.empty_like(y)
        sums = arena_zeros((2, Cout), x.device)      # [0] = sum dact * xhat (dgamma), [1] = sum dact (dbeta)
        L.call('s4_bn_relu_upsample_bwd', _p(dout), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd),
               _p(dact), _p(sums[1]), _p(sums[0]), B, H, W, Cout, s, dt, _st())
        dx = _conv_bn_backward(stage, x, y, dact, sums, mean, invstd, ctx.count, ctx.group_info,
                               B, H, W, Cin, Cout, need_dx=ctx.needs_input_grad[0])
        return dx, None, None, None, None, None, None, None


def _conv_grads(stage, x, dyc, B, H, W, Cin, Cout, need_dx=True):
    """conv3x3 weight gradient (accumulated into conv.weight.grad) and input gradient."""
    conv = stage.conv
    dt = _code(x.dtype)
    if conv.weight.requires_grad:
        L.call('s4_conv3x3_wgrad', _p(x), _p(dyc), _p(grad_buffer(conv.weight)), B, H, W, Cin, Cout, dt,
               backend(), _st())
    dx = None
    if need_dx:
        _, wd = conv_packed(conv.weight)
        dx = torch.empty_like(x)
        L.call('s4_conv3x3_dgrad', _p(dyc), _p(wd), _p(dx), B, H, W, Cin, Cout, dt, backend(), _st())
    return dx


def _conv_bn_backward(stage, x, y, dact, sums, mean, invstd, count, group_info, B, H, W, Cin, Cout,
                      need_dx=True):
    """Shared tail of the stage backward: BN backward (local dgamma/dbeta, global reduction of the
    two sums under SyncBN) -> conv wgrad / dgrad."""
    bn = stage.bn
    dt = _code(x.dtype)
    _add_bn_grads(bn, sums)                 # dgamma = sum dact * xhat, dbeta = sum dact (local, like torch SyncBN)
    gsums = sums
    if group_info is not None and group_info.get('world', 1) > 1:
        gsums = _all_reduce_stats(sums.clone(), group_info)
    dyc = torch.empty_like(y)
    L.call('s4_bn_bwd_apply', _p(dact), _p(y), _p(bn.weight.detach()), _p(mean), _p(invstd),
           _p(gsums[1]), _p(gsums[0]), count, _p(dyc), B * H * W, Cout, dt, _st())
    return _conv_grads(stage, x, dyc, B, H, W, Cin, Cout, need_dx)


class ConvBNReLUClsUpFn(torch.autograd.Function):
    """Last stage: conv3x3 -> BN -> ReLU -> conv_seg (1x1) -> bilinear x s -> NCHW fp32 logits.
    conv_seg is applied before the upsample (they commute), so the 256-channel full-resolution
    tensor of the reference never exists."""

    @staticmethod
    def forward(ctx, x, stage, conv_seg, B, H, W, s, training, group_info):
        conv = stage.conv
        Cout, Cin = conv.weight.shape[0], conv.weight.shape[1]
        NC = conv_seg.weight.shape[0]
        dt = _code(x.dtype)
        rows = B * H * W
        y, stats = _conv_fwd(stage, x, B, H, W, Cin, Cout, training)
        scale, shift, mean, invstd, count = _bn_scale_shift(stage, y, rows, training, group_info, stats)
        z = torch.empty((rows, NC), dtype=torch.float32, device=x.device)
        w2 = conv_seg.weight.detach().reshape(NC, Cout)
        L.call('s4_bn_relu_conv1x1_fwd', _p(y), _p(scale), _p(shift), _p(w2), _p(conv_seg.bias.detach()),
               _p(z), rows, Cout, NC, dt, _st())
        logits = torch.empty((B, NC, H * s, W * s), dtype=torch.float32, device=x.device)
        L.call('s4_upsample_logits_fwd', _p(z), _p(logits), B, H, W, NC, s, _st())
        ctx.stage, ctx.conv_seg, ctx.group_info, ctx.count = stage, conv_seg, group_info, count
        ctx.dims = (B, H, W, s, Cin, Cout, NC)
        if training:
            ctx.save_for_backward(x, y, scale, shift, mean, invstd)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, y, scale, shift, mean, invstd = ctx.saved_tensors
        stage, conv_seg = ctx.stage, ctx.conv_seg
        B, H, W, s, Cin, Cout, NC = ctx.dims
        dt = _code(x.dtype)
        rows = B * H * W
        dlogits = dlogits.contiguous()
        if L.load().s4_cls_supported(Cout, NC, dt):
            # bf16 fast path: y is streamed twice (reduce, apply), the ReLU-masked gradient is
            # never written to memory
            bn = stage.bn
            w2 = conv_seg.weight.detach().reshape(NC, Cout)
            dz16 = torch.empty((rows, 32), dtype=torch.bfloat16, device=x.device)
            L.call('s4_cls_upsample_bwd_padded', _p(dlogits), _p(dz16), B, H, W, NC, s, _st())
            sums = arena_zeros((2, Cout), x.device)  # [0] = dgamma sum, [1] = dbeta sum
            L.call('s4_cls_bwd_reduce', _p(dz16), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd), _p(w2),
                   _p(grad_buffer(conv_seg.weight)), _p(grad_buffer(conv_seg.bias)), _p(sums[1]), _p(sums[0]),
                   rows, Cout, NC, _st())
            _add_bn_grads(bn, sums)
            gsums = sums
            gi = ctx.group_info
            if gi is not None and gi.get('world', 1) > 1:
                gsums = _all_reduce_stats(sums.clone(), gi)
            dyc = torch.empty_like(y)
            L.call('s4_cls_bwd_apply', _p(dz16), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd),
                   _p(bn.weight.detach()), _p(w2), _p(gsums[1]), _p(gsums[0]), ctx.count, _p(dyc), rows, Cout,
                   NC, _st())
            dx = _conv_grads(stage, x, dyc, B, H, W, Cin, Cout, need_dx=ctx.needs_input_grad[0])
            return dx, None, None, None, None, None, None, None, None
        dz = torch.empty((rows, NC), dtype=torch.float32, device=x.device)
        L.call('s4_upsample_logits_bwd', _p(dlogits), _p(dz), B, H, W, NC, s, _st())
        dact = torch.empty_like(y)
        sums = arena_zeros((2, Cout), x.device)      # [0] = dgamma sum, [1] = dbeta sum
        w2 = conv_seg.weight.detach().reshape(NC, Cout)
        L.call('s4_bn_relu_conv1x1_bwd', _p(dz), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd),
               _p(w2), _p(dact), _p(grad_buffer(conv_seg.weight)), _p(grad_buffer(conv_seg.bias)),
               _p(sums[1]), _p(sums[0]), rows, Cout, NC, dt, _st())
        dx = _conv_bn_backward(stage, x, y, dact, sums, mean, invstd, ctx.count, ctx.group_info,
                               B, H, W, Cin, Cout, need_dx=ctx.needs_input_grad[0])
        return dx, None, None, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# losses / pseudo labels / augmentation / EMA
# ----------------------------------------------------------------------------------------------
_zero_scalars = {}


def _zero_scalar(device):
    z = _zero_scalars.get(device)
    if z is None:
        z = torch.zeros((), dtype=torch.float32, device=device)
        _zero_scalars[device] = z
    return z


class CeNcrFn(torch.autograd.Function):
    """(loss_ce, loss_ncr) = (w_ce/P * sum_valid nll, w_ncr/P * sum_valid ||p_s - p_t + eps||).

    When the student logits require grad the forward launch also writes d(loss)/dz_s for unit
    upstream gradients (the logits are streamed once per step instead of twice); backward applies
    the real upstream gradients on the device (``s4_ce_ncr_grad_fixup``: a no-op for (1, 1))."""

    @staticmethod
    def forward(ctx, logits_s, logits_t, label, ce_w, ncr_w, ignore_index):
        _require_cuda(logits_s, label)
        logits_s = logits_s.contiguous()
        assert logits_s.dtype == torch.float32 and label.dtype == torch.int64
        B, Cc, H, W = logits_s.shape
        label = label.reshape(B, H, W).contiguous()
        if logits_t is not None:
            logits_t = logits_t.contiguous()
        ctx.set_materialize_grads(False)     # unused outputs: no zero-filled gradient tensors
        out = torch.empty(3, dtype=torch.float32, device=logits_s.device)
        nbytes = L.load().s4_ce_ncr_workspace(B, H, W)
        ws = workspace(nbytes, logits_s.device, 'loss')
        dz = torch.empty_like(logits_s) if ctx.needs_input_grad[0] else None
        L.call('s4_ce_ncr', _p(logits_s), _p(logits_t), _p(label), _p(dz), _p(out), None, B, Cc, H, W,
               float(ce_w), float(ncr_w), int(ignore_index), _p(ws), nbytes, _st())
        ctx.args = (float(ce_w), float(ncr_w), int(ignore_index))
        ctx.dz = dz
        ctx.save_for_backward(logits_s, logits_t, label)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_ce, g_ncr, _g_cnt):
        logits_s, logits_t, label = ctx.saved_tensors
        ce_w, ncr_w, ignore_index = ctx.args
        B, Cc, H, W = logits_s.shape
        dev = logits_s.device
        dz = ctx.dz
        if dz is None:
            raise RuntimeError('CeNcrFn.backward called twice (the saved gradient buffer is consumed)')
        ctx.dz = None
        zero = _zero_scalar(dev)
        gs = torch.stack([g_ce if g_ce is not None else zero, g_ncr if g_ncr is not None else zero])
        if gs.dtype != torch.float32:
            gs = gs.float()
        L.call('s4_ce_ncr_grad_fixup', _p(logits_s), _p(logits_t), _p(label), _p(dz), _p(gs), B, Cc, H, W,
               ce_w, ncr_w, ignore_index, _st())
        return dz, None, None, None, None, None


def cross_entropy(logits, label, loss_weight=1.0, ignore_index=255):
    return CeNcrFn.apply(logits, None, label, loss_weight, 0.0, ignore_index)[0]


def pseudo_label(logits_t, threshold, patch=16):
    """-> (hard [B,H,W] i64 with 255 where unconfident, conf [B,H,W] i64, u [B,H/p,W/p] f32)."""
    _require_cuda(logits_t)
    logits_t = logits_t.contiguous()
    B, Cc, H, W = logits_t.shape
    dev = logits_t.device
    hard = torch.empty((B, H, W), dtype=torch.int64, device=dev)
    conf = torch.empty_like(hard)
    u = torch.empty((B, H // patch, W // patch), dtype=torch.float32, device=dev)
    L.call('s4_pseudo_label', _p(logits_t), _p(hard), _p(conf), _p(u), B, Cc, H, W, patch,
           float(threshold), _st())
    return hard, conf, u


def cutmix(img, label, boxes):
    """boxes: list of (y0, y1, x0, x1) from the host RNG."""
    _require_cuda(img)
    B, Cc, H, W = img.shape
    if torch.is_tensor(boxes) and boxes.is_cuda:
        bx = boxes                       # [B, 4] int32, already resident (ops.StepParams)
    else:
        bx = torch.as_tensor(boxes, dtype=torch.int32).reshape(B, 4).to(img.device, non_blocking=True)
    out_img = torch.empty_like(img)
    out_lab = torch.empty_like(label) if label is not None else None
    L.call('s4_cutmix', _p(img.contiguous()), _p(label.contiguous() if label is not None else None),
           _p(bx), _p(out_img), _p(out_lab), B, Cc, H, W, _st())
    return out_img, out_lab


def patchshuffle(img, perms, block):
    _require_cuda(img)
    B, Cc, H, W = img.shape
    pm = perms.to(device=img.device, dtype=torch.int64, non_blocking=True).contiguous()
    out = torch.empty_like(img)
    L.call('s4_patchshuffle', _p(img.contiguous()), _p(pm), _p(out), B, Cc, H, W, int(block), _st())
    return out


class TensorTable:
    """Device-resident pointer/chunk table for the multi-tensor kernels."""

    def __init__(self, lists, device, lrs=None):
        chunk = L.load().s4_chunk_elems()
        sizes = [t.numel() for t in lists[0]]
        for lst in lists:
            assert [sizes[i] if t is None else t.numel() for i, t in enumerate(lst)] == sizes
        self.ptrs = [torch.tensor([0 if t is None else t.data_ptr() for t in lst], dtype=torch.int64).to(device)
                     for lst in lists]
        self.sizes = torch.tensor(sizes, dtype=torch.int64).to(device)
        ct, co = [], []
        for i, n in enumerate(sizes):
            for off in range(0, n, chunk):
                ct.append(i)
                co.append(off)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32).to(device)
        self.chunk_off = torch.tensor(co, dtype=torch.int64).to(device)
        self.n_chunks = len(ct)
        if torch.is_tensor(lrs):
            self.lrs = lrs                    # a resident float32 table owned by the caller
        else:
            self.lrs = None if lrs is None else torch.tensor(lrs, dtype=torch.float32).to(device)
        self.key = tuple(0 if t is None else t.data_ptr() for lst in lists for t in lst)
        self.keep = lists


def ema_update(table, momentum):
    """dst = m*dst + (1-m)*src over every tensor pair of the table (one launch); a third list in
    the table holds the bf16 shadows of dst refreshed in the same pass."""
    shadow = table.ptrs[2] if len(table.ptrs) > 2 else None
    L.call('s4_ema_multi_tensor', _p(table.ptrs[0]), _p(table.ptrs[1]), _p(shadow), _p(table.sizes),
           _p(table.chunk_tensor), _p(table.chunk_off), table.n_chunks, float(momentum),
           float(1 - momentum), _st())
    for t in table.keep[0]:
        bump_generation(t)


def sgd_step(table, momentum, weight_decay, first_step, lrs=None):
    """``lrs``: host list -> copied into the table now; None -> the table's ``lrs`` tensor (e.g. a
    ``StepParams`` view refreshed by the step's parameter upload) is used as it is."""
    if lrs is not None:
        table.lrs.copy_(torch.tensor(lrs, dtype=torch.float32), non_blocking=True)
    shadow = table.ptrs[3] if len(table.ptrs) > 3 else None
    L.call('s4_sgd_multi_tensor', _p(table.ptrs[0]), _p(table.ptrs[1]), _p(table.ptrs[2]), _p(shadow),
           _p(table.sizes), _p(table.lrs), _p(table.chunk_tensor), _p(table.chunk_off), table.n_chunks,
           float(momentum), float(weight_decay), int(first_step), _st())
    for t in table.keep[0]:
        bump_generation(t)


def sgd_ema_step(table, momentum, weight_decay, first_step, lrs=None):
    """SGD-momentum + EMA-teacher update in one sweep (``s4_sgd_ema_multi_tensor``).  Table lists:
    [params, grads, momentum buffers, bf16 shadows, ema params, ema bf16 shadows]; ``table.ema_m``
    holds the per-tensor EMA momentum."""
    if lrs is not None:
        table.lrs.copy_(torch.tensor(lrs, dtype=torch.float32), non_blocking=True)
    L.call('s4_sgd_ema_multi_tensor', _p(table.ptrs[0]), _p(table.ptrs[1]), _p(table.ptrs[2]), _p(table.ptrs[3]),
           _p(table.ptrs[4]), _p(table.ptrs[5]), _p(table.ema_m), _p(table.sizes), _p(table.lrs),
           _p(table.chunk_tensor), _p(table.chunk_off), table.n_chunks, float(momentum), float(weight_decay),
           int(first_step), _st())
    for t in table.keep[0]:
        bump_generation(t)
    for t in table.keep[4]:
        if t is not None:
            bump_generation(t)


# ----------------------------------------------------------------------------------------------
# inference / validation (reference encoder_decoder.py:1068-1232, core/evaluation/metrics.py)
# ----------------------------------------------------------------------------------------------
def resize_bilinear(x, size):
    """mmseg ``resize(x, size, mode='bilinear', align_corners=False)`` on NCHW fp32 (any scale)."""
    _require_cuda(x)
    B, Cc, IH, IW = x.shape
    OH, OW = int(size[0]), int(size[1])
    if (IH, IW) == (OH, OW):
        return x
    x = x.float().contiguous()
    out = torch.empty((B, Cc, OH, OW), dtype=torch.float32, device=x.device)
    L.call('s4_resize_bilinear_nchw', _p(x), _p(out), B * Cc, IH, IW, OH, OW, _st())
    return out


def softmax_argmax(logits, want_prob=True, want_pred=True, flip=None):
    """(softmax over dim 1 [flipped], argmax over dim 1): (prob [B,C,H,W] f32 | None, pred [B,H,W] i64 | None)."""
    _require_cuda(logits)
    logits = logits.float().contiguous()
    B, Cc, H, W = logits.shape
    prob = torch.empty_like(logits) if want_prob else None
    pred = torch.empty((B, H, W), dtype=torch.int64, device=logits.device) if want_pred else None
    code = {None: 0, False: 0, 'horizontal': 1, 'vertical': 2}[flip]
    L.call('s4_softmax_argmax_nchw', _p(logits), _p(prob), _p(pred), B, Cc, H, W, code, _st())
    return prob, pred


def accumulate_crop(preds, count, crop, y1, x1):
    B, Cc, H, W = preds.shape
    ch, cw = crop.shape[2], crop.shape[3]
    L.call('s4_accumulate_crop', _p(crop.float().contiguous()), _p(preds), _p(count), B, Cc, H, W, int(y1), int(x1),
           ch, cw, _st())


def intersect_union_hist(pred, label, num_classes, ignore_index, hist=None):
    """Accumulates the three class histograms of ``intersect_and_union`` into ``hist`` ([3, C] int64)."""
    _require_cuda(pred, label)
    pred = pred.to(torch.int64).contiguous()
    label = label.to(device=pred.device, dtype=torch.int64).contiguous()
    assert pred.numel() == label.numel()
    if hist is None:
        hist = torch.zeros((3, num_classes), dtype=torch.int64, device=pred.device)
    L.call('s4_intersect_union', _p(pred), _p(label), pred.numel(), int(num_classes), int(ignore_index), _p(hist), _st())
    return hist
