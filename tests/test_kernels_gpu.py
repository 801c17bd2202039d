"""GPU parity tests of the individual kernels, called through the C ABI (ctypes) and compared
with the oracle / plain fp32 math on the same seeded inputs.  Integer outputs are bit-exact;
fp32 within 1e-3 relative (validation mode); bf16 within 2e-2 relative."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import s4former_oracle as O  # noqa: E402  (the checker)
from s4former_b200 import _lib as L  # noqa: E402
from s4former_b200 import ops  # noqa: E402

DEV = 'cuda'


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.fixture(autouse=True)
def _sync():
    yield
    torch.cuda.synchronize()


def test_library_loads_on_gpu():
    lib = L.load()
    assert lib.s4_built_arch() == 100
    assert torch.cuda.get_device_capability(0)[0] == 10


# ------------------------------------------------------------------------------------------ colsum
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 1e-5)])
@pytest.mark.parametrize('rows,cols', [(1, 8), (77, 768), (8200, 2304), (1000, 256), (33, 21), (4099, 3072)])
def test_colsum_and_sumsq(dtype, tol, rows, cols):
    """bias gradients / BatchNorm statistics: column sums (accumulating) of a row-major matrix."""
    g = gen(3)
    x = torch.randn(rows, cols, generator=g).to(DEV, dtype)
    s = torch.full((cols,), 0.5, device=DEV)
    q = torch.zeros(cols, device=DEV)
    L.call('s4_colsum', x.data_ptr(), s.data_ptr(), q.data_ptr(), rows, cols, ops._code(dtype), ops._st())
    xf = x.double()
    assert rel(s, xf.sum(0) + 0.5) < tol
    assert rel(q, (xf * xf).sum(0)) < tol
    s2 = torch.zeros(cols, device=DEV)
    L.call('s4_colsum', x.data_ptr(), s2.data_ptr(), None, rows, cols, ops._code(dtype), ops._st())
    assert rel(s2, xf.sum(0)) < tol


# ------------------------------------------------------------------------------------------ LN
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize('D', [128, 768])
def test_layernorm_fwd_bwd(dtype, tol, D):
    g = gen(1)
    rows = 77
    x = torch.randn(rows, D, generator=g)
    gamma = torch.randn(D, generator=g) * 0.1 + 1
    beta = torch.randn(D, generator=g) * 0.1
    dy = torch.randn(rows, D, generator=g)
    dres = torch.randn(rows, D, generator=g)
    xd = x.to(DEV, dtype)
    xr = xd.float().cpu().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), gr, br, 1e-6)
    dyd = dy.to(DEV, dtype)
    yr.backward(dyd.float().cpu())
    gp = torch.nn.Parameter(gamma.to(DEV))
    bp = torch.nn.Parameter(beta.to(DEV))
    y, mean, rstd = ops.layernorm_fwd(xd, gp, bp, 1e-6)
    assert rel(y.float(), yr) < tol
    dresd = dres.to(DEV, dtype)
    dx = ops.layernorm_bwd(dyd, xd, gp, bp, mean, rstd, dres=dresd)
    assert rel(dx.float(), xr.grad + dresd.float().cpu()) < max(tol, 1e-4)
    assert rel(gp.grad, gr.grad) < max(tol, 1e-4)
    assert rel(bp.grad, br.grad) < max(tol, 1e-4)
    # the same pass can emit colsum(dres) (+=) : the bias gradient of the layer feeding the residual
    rb = torch.nn.Parameter(torch.zeros(D, device=DEV))
    rb.grad = torch.ones(D, device=DEV)
    dx2 = ops.layernorm_bwd(dyd, xd, gp, bp, mean, rstd, dres=dresd, dres_bias=rb)
    assert torch.equal(dx2, dx)
    assert rel(rb.grad.cpu(), 1.0 + dresd.float().cpu().sum(0)) < 1e-5


def test_layernorm_bwd_many_rows_residual_sum():
    rows, D = 24600, 768
    g = gen(5)
    xd = torch.randn(rows, D, generator=g).to(DEV, torch.bfloat16)
    dyd = torch.randn(rows, D, generator=g).to(DEV, torch.bfloat16)
    dresd = (torch.randn(rows, D, generator=g) + 0.25).to(DEV, torch.bfloat16)
    gp = torch.nn.Parameter(torch.ones(D, device=DEV))
    bp = torch.nn.Parameter(torch.zeros(D, device=DEV))
    rb = torch.nn.Parameter(torch.zeros(D, device=DEV))
    _, mean, rstd = ops.layernorm_fwd(xd, gp, bp, 1e-6)
    ops.layernorm_bwd(dyd, xd, gp, bp, mean, rstd, dres=dresd, dres_bias=rb)
    assert rel(rb.grad, dresd.double().sum(0).float()) < 1e-5
    assert rel(bp.grad, dyd.double().sum(0).float()) < 1e-5


def test_layernorm_fwd_many_rows_lean_kernel():
    """>= 4096 bf16 rows without a row map take the high-occupancy kernel (gamma/beta via L1)."""
    g = gen(3)
    rows, D = 5001, 768
    x = torch.randn(rows, D, generator=g)
    gamma = torch.randn(D, generator=g) * 0.1 + 1
    beta = torch.randn(D, generator=g) * 0.1
    gp, bp = torch.nn.Parameter(gamma.to(DEV)), torch.nn.Parameter(beta.to(DEV))
    y, mean, rstd = ops.layernorm_fwd(x.to(DEV, torch.bfloat16), gp, bp, 1e-6)
    xr = x.to(torch.bfloat16).float()
    want = F.layer_norm(xr, (D,), gamma, beta, 1e-6)
    assert rel(y.float().cpu(), want) < 1e-2
    assert torch.allclose(mean.cpu(), xr.mean(1), atol=1e-4)
    assert torch.allclose(rstd.cpu(), (xr.var(1, unbiased=False) + 1e-6).rsqrt(), rtol=1e-3)


def test_layernorm_row_map():
    g = gen(2)
    x = torch.randn(50, 128, generator=g).to(DEV)
    rm = torch.randperm(50, generator=g)[:30].to(torch.int32).to(DEV)
    gp = torch.nn.Parameter(torch.ones(128, device=DEV))
    bp = torch.nn.Parameter(torch.zeros(128, device=DEV))
    y, mean, rstd = ops.layernorm_fwd(x, gp, bp, 1e-6, row_map=rm, out_rows=30)
    want = F.layer_norm(x[rm.long()], (128,), None, None, 1e-6)
    assert rel(y, want) < 1e-5
    dy = torch.randn(30, 128, generator=g).to(DEV)
    dx = ops.layernorm_bwd(dy, x, gp, bp, mean, rstd, row_map=rm)
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr[rm.long()], (128,), None, None, 1e-6).backward(dy)
    assert rel(dx, xr.grad) < 1e-4


# ------------------------------------------------------------------------------------------ GEMM
def _gemm_case(M, N, K, dtype, backend, a_mn=False, b_mn=False, bias=False, act=False, res=False,
               aux=False, c32=False, accumulate=False, split_k=1, batch=(1, 1), alpha=1.0, seed=3):
    g = gen(seed)
    nb = batch[0] * batch[1]
    A = torch.randn(nb, M, K, generator=g).to(DEV, dtype)
    Bm = torch.randn(nb, K, N, generator=g).to(DEV, dtype)      # logical [K, N]
    ref = alpha * torch.bmm(A.float(), Bm.float())
    a_store = A.transpose(1, 2).contiguous() if a_mn else A    # MN-major: stored [K, M]
    b_store = Bm if b_mn else Bm.transpose(1, 2).contiguous()   # K-major: stored [N, K]
    a_str = (1, M) if a_mn else (K, 1)
    b_str = (N, 1) if b_mn else (1, K)
    cdt = torch.float32 if c32 else dtype
    c = torch.randn(nb, M, N, generator=g).to(DEV, cdt) if accumulate else torch.empty(nb, M, N, device=DEV, dtype=cdt)
    kw = {}
    if bias:
        bv = torch.randn(N, generator=g).to(DEV)
        ref = ref + bv
        kw['bias'] = bv
    if aux:
        av = torch.randn(nb, M, N, generator=g).to(DEV, dtype)
        xa = av.float()
        cdf = 0.5 * (1 + torch.erf(xa / math.sqrt(2)))
        ref = ref * (cdf + xa * torch.exp(-0.5 * xa * xa) / math.sqrt(2 * math.pi))
        kw['aux'] = av
    pre = None
    if act:
        pre = torch.empty(nb, M, N, device=DEV, dtype=dtype)
        kw['pre'] = pre
        pre_ref = ref
        ref = F.gelu(ref)
        kw['act'] = L.ACT_GELU
    if res:
        rv = torch.randn(nb, M, N, generator=g).to(DEV, dtype)
        ref = ref + rv.float()
        kw['res'] = rv
    if accumulate:
        ref = ref + c.float()
    ops.gemm(a_store, b_store, c, M, N, K, a_str, b_str, N, batch=batch,
             a_bs=(batch[1] * M * K, M * K), b_bs=(batch[1] * K * N, K * N), c_bs=(batch[1] * M * N, M * N),
             alpha=alpha, accumulate=accumulate, split_k=split_k, backend_override=backend, **kw)
    torch.cuda.synchronize()
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    r = rel(c.float(), ref)
    assert r < tol, f'rel err {r}'
    if act:
        assert rel(pre.float(), pre_ref) < tol


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_simt_layouts(dtype, a_mn, b_mn):
    _gemm_case(100, 72, 50, dtype, L.BACKEND_SIMT, a_mn=a_mn, b_mn=b_mn, c32=(dtype == torch.float32))


def test_gemm_simt_epilogues_and_batch():
    _gemm_case(65, 40, 33, torch.float32, L.BACKEND_SIMT, bias=True, act=True, res=True, c32=True, batch=(2, 3))
    _gemm_case(65, 40, 33, torch.float32, L.BACKEND_SIMT, aux=True, c32=True, accumulate=True, alpha=0.5)
    _gemm_case(65, 40, 33, torch.bfloat16, L.BACKEND_SIMT, bias=True, res=True, c32=True)


# ------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize('pasa', [False, True])
def test_attention_composed(dtype, tol, pasa):
    g = gen(5)
    B, H, Lt, hd = 2, 2, 65, 64
    D = H * hd
    qkv = (torch.randn(B * Lt, 3 * D, generator=g) * 0.5).to(DEV, dtype)
    u0 = gate = None
    w = 0.0
    if pasa:
        u = torch.rand(B, 8, 8, generator=g)
        u0, gate = O.pasa_gate_u0(u, True)
        u0, gate, w = u0.to(DEV), gate.to(DEV), 5.0
    old = ops.backend()
    ops.set_backend('simt')
    try:
        out, lse = ops.attention_fwd(qkv, B, Lt, H, hd, u0, gate, w)
        dout = torch.randn(B * Lt, D, generator=g).to(DEV, dtype)
        dqkv = ops.attention_bwd(dout, qkv, out, lse, B, Lt, H, hd, u0, gate, w)
    finally:
        ops.set_backend(old)
    qr = qkv.float().cpu().requires_grad_(True)
    q, k, v = qr.view(B, Lt, 3 * D).split(D, dim=-1)
    q = q.view(B, Lt, H, hd).transpose(1, 2) / math.sqrt(hd)
    k = k.view(B, Lt, H, hd).transpose(1, 2)
    v = v.view(B, Lt, H, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if pasa:
        s = s + (w * gate.cpu().unsqueeze(-1) * u0.cpu().unsqueeze(1)).unsqueeze(1)
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * Lt, D)
    o.backward(dout.float().cpu())
    assert rel(out.float(), o) < tol
    assert rel(dqkv.float(), qr.grad) < tol


# ------------------------------------------------------------------------------------------ head
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_conv3x3_simt(dtype, tol):
    g = gen(6)
    B, H, W, Cin, Cout = 2, 8, 8, 16, 24
    x = torch.randn(B, H, W, Cin, generator=g).to(DEV, dtype)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.2).to(DEV)
    dy = torch.randn(B, H, W, Cout, generator=g).to(DEV, dtype)
    code = L.BF16 if dtype == torch.bfloat16 else L.F32
    wf = torch.empty(Cout, 9 * Cin, device=DEV, dtype=dtype)
    wd = torch.empty(Cin, 9 * Cout, device=DEV, dtype=dtype)
    st = torch.cuda.current_stream().cuda_stream
    L.call('s4_pack_conv3x3_weight', w.data_ptr(), wf.data_ptr(), wd.data_ptr(), Cin, Cout, code, st)
    y = torch.empty(B, H, W, Cout, device=DEV, dtype=dtype)
    L.call('s4_conv3x3_fwd', x.data_ptr(), wf.data_ptr(), y.data_ptr(), B, H, W, Cin, Cout, code, L.BACKEND_SIMT, st)
    dx = torch.empty_like(x)
    L.call('s4_conv3x3_dgrad', dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), B, H, W, Cin, Cout, code, L.BACKEND_SIMT, st)
    dw = torch.zeros(Cout, Cin, 3, 3, device=DEV)
    L.call('s4_conv3x3_wgrad', x.data_ptr(), dy.data_ptr(), dw.data_ptr(), B, H, W, Cin, Cout, code, L.BACKEND_SIMT, st)
    # reference on the CPU: cuDNN would silently use TF32 for an fp32 conv
    xr = x.float().cpu().permute(0, 3, 1, 2).clone().requires_grad_(True)
    wr = wf.float().cpu().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1)
    yr.backward(dy.float().cpu().permute(0, 3, 1, 2))
    assert rel(y.float().permute(0, 3, 1, 2), yr) < tol
    assert rel(dx.float().permute(0, 3, 1, 2), xr.grad) < tol
    assert rel(dw, wr.grad) < tol


class _Stage(torch.nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.bn = torch.nn.BatchNorm2d(cout)


@pytest.mark.parametrize('s', [2, 4])
def test_conv_bn_relu_upsample_stage_fp32(s):
    ops.set_compute_dtype(torch.float32)
    try:
        g = gen(7)
        B, H, W, Cin, Cout = 2, 8, 8, 16, 32
        stage = _Stage(Cin, Cout).to(DEV)
        ref = _Stage(Cin, Cout)
        with torch.no_grad():
            stage.bn.weight.copy_(torch.rand(Cout, generator=g) + 0.5)
            stage.bn.bias.copy_(torch.randn(Cout, generator=g) * 0.1)
        ref.load_state_dict({k: v.cpu() for k, v in stage.state_dict().items()})
        x = torch.randn(B, H, W, Cin, generator=g)
        xd = x.to(DEV).reshape(B * H * W, Cin).requires_grad_(True)
        out = ops.ConvBNReLUUpFn.apply(xd, stage, B, H, W, s, True, None)
        dout = torch.randn(B * H * s * W * s, Cout, generator=g)
        out.backward(dout.to(DEV))
        xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
        yr = F.interpolate(F.relu(ref.bn(ref.conv(xr))), scale_factor=s, mode='bilinear', align_corners=False)
        yr.backward(dout.view(B, H * s, W * s, Cout).permute(0, 3, 1, 2))
        assert rel(out.view(B, H * s, W * s, Cout).permute(0, 3, 1, 2), yr) < 1e-4
        assert rel(xd.grad.view(B, H, W, Cin).permute(0, 3, 1, 2), xr.grad) < 1e-3
        assert rel(stage.conv.weight.grad, ref.conv.weight.grad) < 1e-3
        assert rel(stage.bn.weight.grad, ref.bn.weight.grad) < 1e-3
        assert rel(stage.bn.bias.grad, ref.bn.bias.grad) < 1e-3
        assert rel(stage.bn.running_mean, ref.bn.running_mean) < 1e-4
        assert rel(stage.bn.running_var, ref.bn.running_var) < 1e-4
    finally:
        ops.set_compute_dtype(torch.bfloat16)


@pytest.mark.parametrize('dtype,tol', [(torch.bfloat16, 1e-2), (torch.float32, 1e-5)])
@pytest.mark.parametrize('B,H,W,C,s', [(2, 16, 16, 256, 2), (1, 8, 24, 256, 4), (2, 5, 7, 64, 2), (1, 32, 32, 128, 2),
                                       (2, 64, 64, 256, 2), (1, 40, 36, 256, 2), (2, 32, 32, 256, 4),
                                       (12, 32, 32, 256, 2), (1, 2, 2, 256, 2), (1, 3, 5, 256, 4)])
def test_bn_relu_upsample_kernels(dtype, tol, B, H, W, C, s):
    """fused BN + ReLU + bilinear (and its transpose with the BatchNorm-backward sums) against
    F.interpolate on the same inputs, incl. non-square maps and every border case."""
    g = gen(9)
    x = torch.randn(B, H, W, C, generator=g).to(DEV, dtype)
    scale = (torch.rand(C, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(C, generator=g) * 0.3).to(DEV)
    mean = (torch.randn(C, generator=g) * 0.1).to(DEV)
    invstd = (torch.rand(C, generator=g) + 0.5).to(DEV)
    dout = torch.randn(B, H * s, W * s, C, generator=g).to(DEV, dtype)
    st = ops._st()
    dt = ops._code(dtype)
    out = torch.empty(B, H * s, W * s, C, device=DEV, dtype=dtype)
    L.call('s4_bn_relu_upsample_fwd', x.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(),
           B, H, W, C, s, dt, st)
    a = (x.float() * scale + shift).permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(F.relu(a), scale_factor=s, mode='bilinear', align_corners=False)
    assert rel(out.float().permute(0, 3, 1, 2), ref) < tol
    ref.backward(dout.float().permute(0, 3, 1, 2))
    dact = torch.empty_like(x)
    sums = torch.zeros(2, C, device=DEV)
    L.call('s4_bn_relu_upsample_bwd', dout.data_ptr(), x.data_ptr(), scale.data_ptr(), shift.data_ptr(),
           mean.data_ptr(), invstd.data_ptr(), dact.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(),
           B, H, W, C, s, dt, st)
    da_ref = a.grad.permute(0, 2, 3, 1)
    assert rel(dact.float(), da_ref) < tol
    xhat = (x.float() - mean) * invstd
    assert rel(sums[0], da_ref.sum((0, 1, 2))) < max(tol, 1e-4)
    assert rel(sums[1], (da_ref * xhat).sum((0, 1, 2))) < max(tol, 1e-4)
    # BatchNorm backward apply on the same tensors
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV)
    n = float(B * H * W)
    dy = torch.empty_like(x)
    L.call('s4_bn_bwd_apply', dact.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
           sums[0].data_ptr(), sums[1].data_ptr(), n, dy.data_ptr(), B * H * W, C, dt, st)
    want = gamma * invstd * (dact.float() - sums[0] / n - xhat * sums[1] / n)
    assert rel(dy.float(), want) < tol


def test_conv_bn_relu_cls_upsample_stage_fp32():
    ops.set_compute_dtype(torch.float32)
    try:
        g = gen(8)
        B, H, W, Cin, Cout, NC, s = 2, 8, 8, 16, 32, 5, 2
        stage = _Stage(Cin, Cout).to(DEV)
        seg = torch.nn.Conv2d(Cout, NC, 1).to(DEV)
        ref, rseg = _Stage(Cin, Cout), torch.nn.Conv2d(Cout, NC, 1)
        ref.load_state_dict({k: v.cpu() for k, v in stage.state_dict().items()})
        rseg.load_state_dict({k: v.cpu() for k, v in seg.state_dict().items()})
        x = torch.randn(B, H, W, Cin, generator=g)
        xd = x.to(DEV).reshape(B * H * W, Cin).requires_grad_(True)
        out = ops.ConvBNReLUClsUpFn.apply(xd, stage, seg, B, H, W, s, True, None)
        dout = torch.randn(B, NC, H * s, W * s, generator=g)
        out.backward(dout.to(DEV))
        xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
        # the reference order: upsample THEN conv_seg
        yr = rseg(F.interpolate(F.relu(ref.bn(ref.conv(xr))), scale_factor=s, mode='bilinear', align_corners=False))
        yr.backward(dout)
        assert rel(out, yr) < 1e-4
        assert rel(xd.grad.view(B, H, W, Cin).permute(0, 3, 1, 2), xr.grad) < 1e-3
        assert rel(stage.conv.weight.grad, ref.conv.weight.grad) < 1e-3
        assert rel(seg.weight.grad, rseg.weight.grad) < 1e-3
        assert rel(seg.bias.grad, rseg.bias.grad) < 1e-3
        assert rel(stage.bn.weight.grad, ref.bn.weight.grad) < 1e-3
    finally:
        ops.set_compute_dtype(torch.bfloat16)


# ------------------------------------------------------------------------------------------ losses
def test_pseudo_label_bit_exact(golden_dir):
    G = torch.load(os.path.join(golden_dir, 'loss_pseudo.pt'), weights_only=False)
    hard, conf, u = ops.pseudo_label(G['z_t'].to(DEV), 0.95, patch=G['patch'])
    assert torch.equal(hard.cpu(), G['hard'])
    assert torch.equal(conf.cpu(), G['conf'])
    assert torch.equal(u.cpu(), G['u'])


def test_pseudo_label_large_random():
    g = gen(9)
    z = (torch.randn(2, 21, 256, 256, generator=g) * 4)
    hard, conf, u = ops.pseudo_label(z.to(DEV), 0.95, 16)
    h0, c0, mv = O.pseudo_label(z, 0.95)
    near = (mv - 0.95).abs() < 1e-6          # 1-ulp band where expf implementations may differ
    assert torch.equal(conf.cpu()[~near], c0[~near])
    assert torch.equal(hard.cpu()[~near], h0[~near])
    assert int(near.sum()) < 16
    if int(near.sum()) == 0:
        assert torch.equal(u.cpu(), O.patch_unconfidence(c0, 16))
    # W % 64 != 0: scalar kernel; C > 24: wide instantiation of the vector kernel; patch 8
    for shape in ((1, 19, 32, 96), (1, 30, 32, 128)):
        z = (torch.randn(*shape, generator=g) * 4)
        hard, conf, u = ops.pseudo_label(z.to(DEV), 0.95, 8)
        h0, c0, mv = O.pseudo_label(z, 0.95)
        near = (mv - 0.95).abs() < 1e-6
        assert torch.equal(conf.cpu()[~near], c0[~near])
        assert torch.equal(hard.cpu()[~near], h0[~near])
        if int(near.sum()) == 0:
            bb, hh, ww = c0.shape
            want = (1 - c0).view(bb, hh // 8, 8, ww // 8, 8).permute(0, 1, 3, 2, 4).reshape(bb, hh // 8, ww // 8, -1)
            assert torch.equal(u.cpu(), want.sum(-1) / 64)


def test_ce_ncr_golden_and_grad(golden_dir):
    G = torch.load(os.path.join(golden_dir, 'loss_pseudo.pt'), weights_only=False)
    zs = G['z_s'].to(DEV).requires_grad_(True)
    zt = G['z_t'].to(DEV)
    hard = G['hard'].to(DEV)
    ce, ncr, _ = ops.CeNcrFn.apply(zs, zt, hard, 1.0, 1.0, 255)
    assert abs(float(ce) - float(G['loss_seg_unsup'])) < 1e-5 * abs(float(G['loss_seg_unsup']))
    assert abs(float(ncr) - float(G['loss_ncr_unsup'])) < 1e-4 * abs(float(G['loss_ncr_unsup']))
    (0.5 * ce + 0.25 * ncr).backward()
    zr = G['z_s'].clone().requires_grad_(True)
    lo = 0.5 * O.cross_entropy_mean_all(zr, G['hard']) + 0.25 * O.ncr_unsup_only(zr, G['z_t'], G['hard'])
    lo.backward()
    assert rel(zs.grad, zr.grad) < 1e-4
    w = ops.cross_entropy(G['z_s'].to(DEV), hard, 0.4, 255)
    assert abs(float(w) - float(G['ce_w04'])) < 1e-5 * abs(float(G['ce_w04']))


@pytest.mark.parametrize('g_ce,g_ncr,with_t', [(1.0, 1.0, True), (0.5, 0.5, True), (2.0, 0.25, True),
                                               (1.0, 0.0, False), (0.4, 0.0, False)])
def test_ce_ncr_fused_gradient_paths(golden_dir, g_ce, g_ncr, with_t):
    """The forward launch also emits the gradient for unit upstream gradients; backward leaves it
    alone (1,1), rescales it (equal) or recomputes it (different) -- all decided on the device."""
    G = torch.load(os.path.join(golden_dir, 'loss_pseudo.pt'), weights_only=False)
    zs = G['z_s'].to(DEV).requires_grad_(True)
    hard = G['hard'].to(DEV)
    zt = G['z_t'].to(DEV) if with_t else None
    ce, ncr, _ = ops.CeNcrFn.apply(zs, zt, hard, 1.0, 1.0 if with_t else 0.0, 255)
    (g_ce * ce + g_ncr * ncr).backward()
    zr = G['z_s'].clone().requires_grad_(True)
    lo = g_ce * O.cross_entropy_mean_all(zr, G['hard'])
    if with_t:
        lo = lo + g_ncr * O.ncr_unsup_only(zr, G['z_t'], G['hard'])
    lo.backward()
    assert rel(zs.grad, zr.grad) < 1e-4


def test_ce_ncr_extreme_logits():
    """label far above / far below the other classes: no overflow, no 0/0 in the NCR softmax."""
    g = gen(11)
    z = torch.randn(1, 21, 16, 32, generator=g)
    zt = torch.randn(1, 21, 16, 32, generator=g)
    y = torch.randint(0, 21, (1, 16, 32), generator=g)
    z.scatter_(1, y[:, None], 150.0)           # the label dominates
    z[:, :, 8:] = z[:, :, 8:] - 300.0 * torch.nn.functional.one_hot(y[:, 8:], 21).permute(0, 3, 1, 2)
    zs = z.to(DEV).requires_grad_(True)
    ce, ncr, _ = ops.CeNcrFn.apply(zs, zt.to(DEV), y.to(DEV), 1.0, 1.0, 255)
    (ce + ncr).backward()
    zr = z.clone().requires_grad_(True)
    lo_ce, lo_ncr = O.cross_entropy_mean_all(zr, y), O.ncr_unsup_only(zr, zt, y)
    (lo_ce + lo_ncr).backward()
    assert torch.isfinite(zs.grad).all()
    assert abs(float(ce) - float(lo_ce)) < 1e-4 * abs(float(lo_ce))
    assert abs(float(ncr) - float(lo_ncr)) < 1e-4 * abs(float(lo_ncr)) + 1e-7
    assert rel(zs.grad, zr.grad) < 1e-3


def test_ce_known_answers_gpu():
    """reference tests/test_models/test_losses/test_ce_loss.py:25-39 through the CUDA kernel."""
    z = torch.tensor([[100., -100.]]).view(1, 2, 1, 1).to(DEV)
    y = torch.tensor([1]).view(1, 1, 1).to(DEV)
    assert abs(float(ops.cross_entropy(z, y, 1.0, 255)) - 200.0) < 1e-4
    y255 = torch.full((1, 1, 1), 255, dtype=torch.int64, device=DEV)
    assert float(ops.cross_entropy(z, y255, 1.0, 255)) == 0.0


# ------------------------------------------------------------------------------------------ augment / EMA
def test_augment_golden(golden_dir):
    G = torch.load(os.path.join(golden_dir, 'augment.pt'), weights_only=False)
    from s4former_b200.utils import generate_unsup_data as gud
    O.seed_host_rng(G['seed'])
    t, s = gud.generate_unsup_cutmix_data(dict(hard_seg_label=G['lab'].to(DEV)), dict(img=G['img'].to(DEV)))
    assert torch.equal(s['img'].cpu(), G['cut_img'])
    assert torch.equal(t['hard_seg_label'].cpu(), G['cut_lab'])
    metas = [dict() for _ in range(4)]
    res = dict(img=s['img'], img_metas=metas)
    res, _ = gud.generate_unsup_patchmix_data(res, t, PatchMix_N=1, patchmix_ratio=0.5)
    assert torch.equal(torch.stack([m['PatchMixIndex'] for m in metas]), G['perms'])
    assert torch.equal(res['img'].cpu(), G['shuf_img'])


def test_token_unshuffle_rowmap(golden_dir):
    G = torch.load(os.path.join(golden_dir, 'augment.pt'), weights_only=False)
    import s4former_b200 as s4
    head = s4.SETRUPHead(in_channels=16, channels=8, num_classes=3, num_convs=1, up_scale=2, in_index=0,
                         dropout_ratio=0, norm_cfg=dict(type='BN'))
    rm = head._row_map(4, 8, False, 'cpu', 2, G['perms'])
    got = G['tok'].reshape(4 * 64, 16)[rm.long()].view(4, 64, 16)
    assert torch.equal(got, G['unsh'])


def test_ema_multi_tensor():
    g = gen(10)
    shapes = [(1000,), (33, 7), (70000,), (5,)]
    src = [torch.randn(s, generator=g).to(DEV) for s in shapes]
    dst = [torch.randn(s, generator=g).to(DEV) for s in shapes]
    want = [d.clone().mul_(0.999).add_(s, alpha=1 - 0.999) for d, s in zip(dst, src)]
    table = ops.TensorTable([dst, src], DEV)
    ops.ema_update(table, 0.999)
    for d, w in zip(dst, want):
        assert torch.allclose(d, w, rtol=1e-6, atol=1e-7)
    # with bf16 shadows (None = no shadow for that tensor), refreshed in the same pass
    shadows = [torch.zeros(s, device=DEV, dtype=torch.bfloat16) if i != 1 else None for i, s in enumerate(shapes)]
    want2 = [d.clone().mul_(0.5).add_(s, alpha=0.5) for d, s in zip(dst, src)]
    table = ops.TensorTable([dst, src, shadows], DEV)
    ops.ema_update(table, 0.5)
    for d, w, sh in zip(dst, want2, shadows):
        assert torch.allclose(d, w, rtol=1e-6, atol=1e-7)
        if sh is not None:
            assert torch.equal(sh, d.to(torch.bfloat16))


def test_sgd_multi_tensor_with_shadows():
    """fused SGD-momentum step (torch.optim.SGD semantics) + the bf16 weight shadows it refreshes."""
    g = gen(12)
    shapes = [(4099,), (64, 48), (16384 * 2 + 5,)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = torch.optim.SGD(ref, lr=0.1, momentum=0.9, weight_decay=0.01)
    bufs = [torch.zeros_like(p) for p in ps]
    shadows = [torch.zeros(s, device=DEV, dtype=torch.bfloat16) if i != 0 else None for i, s in enumerate(shapes)]
    for step in range(3):
        grads = [torch.randn(s, generator=g).to(DEV) for s in shapes]
        for p, r, gr in zip(ps, ref, grads):
            p.grad = gr.clone()
            r.grad = gr.clone()
        opt.step()
        table = ops.TensorTable([[p.data for p in ps], [p.grad for p in ps], bufs, shadows], DEV, lrs=[0.1] * 3)
        ops.sgd_step(table, 0.9, 0.01, first_step=(step == 0))
        for p, r, sh in zip(ps, ref, shadows):
            assert torch.allclose(p, r, rtol=1e-5, atol=1e-6)
            if sh is not None:
                assert torch.equal(sh, p.detach().to(torch.bfloat16))


def test_weight_shadow_tracks_parameter_updates():
    """ops.lowp: a persistent bf16 buffer, re-cast only when the parameter changed behind its back."""
    p = torch.nn.Parameter(torch.randn(8, 16, device=DEV))
    a = ops.lowp(p)
    assert a.dtype == torch.bfloat16 and torch.equal(a, p.detach().to(torch.bfloat16))
    assert ops.lowp(p).data_ptr() == a.data_ptr()
    with torch.no_grad():
        p.add_(1.0)                      # torch in-place op: version counter moves
    b = ops.lowp(p)
    assert b.data_ptr() == a.data_ptr() and torch.equal(b, p.detach().to(torch.bfloat16))
    p.data.mul_(2.0)                     # raw change + explicit generation bump (what the kernels do)
    ops.bump_generation(p)
    assert torch.equal(ops.lowp(p), p.detach().to(torch.bfloat16))


def test_patchify_and_tokens():
    g = gen(11)
    img = torch.randn(2, 3, 64, 64, generator=g).to(DEV)
    a = torch.empty(2 * 16, 768, device=DEV)
    L.call('s4_patchify', img.data_ptr(), a.data_ptr(), 2, 3, 64, 64, 16, L.F32, torch.cuda.current_stream().cuda_stream)
    want = F.unfold(img, 16, stride=16).transpose(1, 2).reshape(32, 768)
    assert torch.equal(a, want)
    # bf16 vector path (8 pixels per thread), including the reference's corner padding
    # (embed.py:58-80: H, W padded up to a multiple of the patch size with zeros)
    for (Hh, Ww) in ((64, 64), (56, 72)):
        img = torch.randn(3, 3, Hh, Ww, generator=g).to(DEV)
        gh, gw = (Hh + 15) // 16, (Ww + 15) // 16
        a16 = torch.empty(3 * gh * gw, 768, device=DEV, dtype=torch.bfloat16)
        L.call('s4_patchify', img.data_ptr(), a16.data_ptr(), 3, 3, Hh, Ww, 16, L.BF16,
               torch.cuda.current_stream().cuda_stream)
        pad = F.pad(img, (0, gw * 16 - Ww, 0, gh * 16 - Hh))
        want = F.unfold(pad, 16, stride=16).transpose(1, 2).reshape(3 * gh * gw, 768).to(torch.bfloat16)
        assert torch.equal(a16, want)
