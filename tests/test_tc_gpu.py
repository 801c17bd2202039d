"""GPU tests of the tcgen05 / TMEM / TMA kernels (gemm_tc.cu) against fp32 math on the same
bf16 inputs (tolerance 2e-2 relative, the bf16 gate of north_star; typical error is ~3e-3)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from s4former_b200 import _lib as L  # noqa: E402
from s4former_b200 import ops  # noqa: E402
from test_kernels_gpu import _gemm_case, rel, gen, DEV  # noqa: E402

BF = torch.bfloat16


@pytest.fixture(autouse=True)
def _sync():
    yield
    torch.cuda.synchronize()


def test_tc_is_selected():
    g = L.GemmParams()
    a = torch.zeros(256, 64, device=DEV, dtype=BF)
    g.a, g.b, g.c = a.data_ptr(), a.data_ptr(), a.data_ptr()
    g.M, g.N, g.K, g.nb1, g.nb2 = 256, 256, 64, 1, 1
    g.a_sm, g.a_sk, g.b_sk, g.b_sn, g.c_sm = 64, 1, 1, 64, 256
    g.dtype, g.c_dtype, g.backend, g.split_k, g.alpha = L.BF16, L.BF16, L.BACKEND_AUTO, 1, 1.0
    import ctypes
    assert L.load().s4_gemm_uses_tc(ctypes.byref(g)) == 1


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (128, 128, 64), (128, 256, 128), (300, 200, 72),
                                   (1025, 768, 768), (2050, 2304, 768), (520, 3072, 768)])
def test_tc_gemm_kmajor(M, N, K):
    _gemm_case(M, N, K, BF, L.BACKEND_TC)


def test_tc_gemm_fp32_out_and_accumulate():
    _gemm_case(260, 136, 192, BF, L.BACKEND_TC, c32=True)
    _gemm_case(260, 136, 192, BF, L.BACKEND_TC, c32=True, accumulate=True, alpha=0.5)
    _gemm_case(260, 136, 192, BF, L.BACKEND_TC, accumulate=True)


def test_tc_gemm_epilogues():
    _gemm_case(300, 256, 128, BF, L.BACKEND_TC, bias=True)
    _gemm_case(300, 256, 128, BF, L.BACKEND_TC, bias=True, act=True)
    _gemm_case(300, 256, 128, BF, L.BACKEND_TC, bias=True, res=True)
    _gemm_case(300, 256, 128, BF, L.BACKEND_TC, aux=True)
    _gemm_case(300, 256, 128, BF, L.BACKEND_TC, bias=True, act=True, res=True)


def test_tc_gemm_ragged_n():
    """S = scale * Q K^T with N = L = 1025 (not a multiple of 8) into an fp32 buffer whose leading
    dimension is padded to 1032, as attention.cu calls it; an unpadded ldc is routed to CUDA cores."""
    g = gen(4)
    M, N, K, ldc = 130, 1025, 64, 1032
    A = torch.randn(M, K, generator=g).to(DEV, BF)
    Bt = torch.randn(N, K, generator=g).to(DEV, BF)
    c = torch.zeros(M, ldc, device=DEV)
    ops.gemm(A, Bt, c, M, N, K, (K, 1), (1, K), ldc, alpha=0.125, backend_override=L.BACKEND_TC)
    ref = 0.125 * A.float() @ Bt.float().t()
    assert rel(c[:, :N], ref) < 2e-2
    assert float(c[:, N:].abs().max()) == 0.0
    _gemm_case(130, 1025, 64, BF, L.BACKEND_AUTO, c32=True, alpha=0.125)   # falls back to CUDA cores


def test_tc_gemm_batched():
    _gemm_case(200, 64, 136, BF, L.BACKEND_TC, batch=(2, 3))
    _gemm_case(136, 136, 64, BF, L.BACKEND_TC, batch=(2, 3), c32=True)


def test_tc_gemm_b_mnmajor():
    _gemm_case(256, 128, 128, BF, L.BACKEND_TC, b_mn=True)
    _gemm_case(300, 64, 200, BF, L.BACKEND_TC, b_mn=True)


def test_tc_gemm_a_mnmajor():
    _gemm_case(256, 128, 128, BF, L.BACKEND_TC, a_mn=True)
    _gemm_case(200, 72, 264, BF, L.BACKEND_TC, a_mn=True)


def test_tc_gemm_ab_mnmajor_splitk():
    _gemm_case(256, 256, 512, BF, L.BACKEND_TC, a_mn=True, b_mn=True, c32=True, accumulate=True)
    _gemm_case(768, 768, 2048, BF, L.BACKEND_TC, a_mn=True, b_mn=True, c32=True, accumulate=True, split_k=8)
    _gemm_case(136, 200, 1000, BF, L.BACKEND_TC, a_mn=True, b_mn=True, c32=True, accumulate=True, split_k=4)


def _conv_case(B, H, W, Cin, Cout, seed=20):
    g = gen(seed)
    x = torch.randn(B, H, W, Cin, generator=g).to(DEV, BF)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))).to(DEV)
    dy = torch.randn(B, H, W, Cout, generator=g).to(DEV, BF)
    st = torch.cuda.current_stream().cuda_stream
    wf = torch.empty(Cout, 9 * Cin, device=DEV, dtype=BF)
    wd = torch.empty(Cin, 9 * Cout, device=DEV, dtype=BF)
    L.call('s4_pack_conv3x3_weight', w.data_ptr(), wf.data_ptr(), wd.data_ptr(), Cin, Cout, L.BF16, st)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    wr = wf.float().view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    return x, dy, wf, wd, xr, wr, yr, st


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(2, 32, 32, 128, 64), (1, 64, 64, 64, 256), (1, 128, 128, 64, 128),
                                            (1, 256, 256, 64, 64), (2, 16, 16, 64, 64)])
def test_tc_conv3x3_fwd_dgrad(B, H, W, Cin, Cout):
    x, dy, wf, wd, xr, wr, yr, st = _conv_case(B, H, W, Cin, Cout)
    y = torch.empty(B, H, W, Cout, device=DEV, dtype=BF)
    L.call('s4_conv3x3_fwd', x.data_ptr(), wf.data_ptr(), y.data_ptr(), B, H, W, Cin, Cout, L.BF16, L.BACKEND_TC, st)
    assert rel(y.float().permute(0, 3, 1, 2), yr) < 2e-2
    if Cout % 64 == 0:
        dx = torch.empty_like(x)
        L.call('s4_conv3x3_dgrad', dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), B, H, W, Cin, Cout, L.BF16, L.BACKEND_TC, st)
        assert rel(dx.float().permute(0, 3, 1, 2), xr.grad) < 2e-2


@pytest.mark.parametrize('B,H,W,Cin,Cout', [(2, 32, 32, 128, 64), (1, 64, 64, 64, 256), (2, 128, 128, 256, 256)])
def test_tc_conv3x3_wgrad(B, H, W, Cin, Cout):
    x, dy, wf, wd, xr, wr, yr, st = _conv_case(B, H, W, Cin, Cout)
    dw = torch.zeros(Cout, Cin, 3, 3, device=DEV)
    L.call('s4_conv3x3_wgrad', x.data_ptr(), dy.data_ptr(), dw.data_ptr(), B, H, W, Cin, Cout, L.BF16, L.BACKEND_TC, st)
    assert rel(dw, wr.grad) < 2e-2
