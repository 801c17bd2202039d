"""Micro-benchmarks of single ops at the train step's shapes (CUDA events, L2-cold-ish: the
operands of one call exceed nothing special, so a 256 MB scratch write runs between calls).
Usage: python tools/bench_ops.py [attn_fwd|attn_bwd|gemm|all] [reps]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from s4former_b200 import ops, _lib as L

dev = 'cuda'
BF = torch.bfloat16
what = sys.argv[1] if len(sys.argv) > 1 else 'all'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
scratch = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, flops=None, name=''):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        scratch.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    extra = f'  {flops / med / 1e9:8.1f} TF/s' if flops else ''
    print(f'{name:40s} {med * 1e3:9.1f} us (min {ts[0] * 1e3:.1f}){extra}', flush=True)


B, H, Lt, hd = int(os.environ.get('S4_BENCH_B', '8')), 12, 1025, 64
D = H * hd
M = B * Lt
if what in ('attn_fwd', 'all'):
    qkv = (torch.randn(M, 3 * D, device=dev) * 0.5).to(BF)
    u0 = torch.rand(B, Lt, device=dev)
    gate = (torch.rand(B, Lt, device=dev) > 0.5).float()
    fl = 4.0 * B * H * Lt * Lt * hd
    timeit(lambda: ops.attention_fwd(qkv, B, Lt, H, hd, None, None, 0.0), fl, 'attn_fwd (no bias)')
    timeit(lambda: ops.attention_fwd(qkv, B, Lt, H, hd, u0, gate, 5.0), fl, 'attn_fwd (PASA bias)')
if what in ('attn_bwd', 'all'):
    qkv = (torch.randn(M, 3 * D, device=dev) * 0.5).to(BF)
    out, lse = ops.attention_fwd(qkv, B, Lt, H, hd, None, None, 0.0)
    dout = torch.randn(M, D, device=dev).to(BF)
    fl = 10.0 * B * H * Lt * Lt * hd
    timeit(lambda: ops.attention_bwd(dout, qkv, out, lse, B, Lt, H, hd, None, None, 0.0), fl, 'attn_bwd')
if what == 'sweep':
    Ms = 148 * 128
    for K in (768, 3072):
        for N in (256, 512, 1024, 2048):
            a = torch.randn(Ms, K, device=dev).to(BF)
            w = torch.randn(N, K, device=dev).to(BF)
            timeit(lambda: ops.linear_fwd(a, w, None), 2.0 * Ms * N * K, f'M={Ms} N={N} K={K} ({N // 256} tiles/SM)')
if what == 'fc1':
    a = torch.randn(M, D, device=dev).to(BF)
    w = torch.randn(4 * D, D, device=dev).to(BF)
    bias = torch.randn(4 * D, device=dev)
    timeit(lambda: ops.linear_fwd(a, w, bias), 2.0 * M * 4 * D * D, 'fc1 fwd bias')
if what == 'kslope':
    # one full wave of tiles, K swept: slope = steady-state time per k-block, intercept = fixed cost
    lib = L.load()
    Ms = 148 * 128
    for mode in (0, 2):
        lib.s4_set_tc_pair_mode(mode)
        for N in (256, 512):
            for K in (768, 1536, 3072, 6144):
                a = torch.randn(Ms, K, device=dev).to(BF)
                w = torch.randn(N, K, device=dev).to(BF)
                timeit(lambda: ops.linear_fwd(a, w, None), 2.0 * Ms * N * K,
                       f'[pair mode {mode}] M={Ms} N={N} K={K} kb={K // 64}')
    lib.s4_set_tc_pair_mode(1)
if what == 'gemm_modes':
    lib = L.load()
    for N, K, nm in ((3 * D, D, 'qkv'), (D, D, 'out_proj'), (4 * D, D, 'fc1'), (D, 4 * D, 'fc2')):
        a = torch.randn(M, K, device=dev).to(BF)
        w = torch.randn(N, K, device=dev).to(BF)
        bias = torch.randn(N, device=dev)
        dy = torch.randn(M, N, device=dev).to(BF)
        wt = w.t().contiguous()
        wp = torch.nn.Parameter(torch.randn(N, K, device=dev))
        fl = 2.0 * M * N * K
        for mode in (0, 1, 2):
            lib.s4_set_tc_pair_mode(mode)
            timeit(lambda: ops.linear_fwd(a, w, bias), fl, f'[pair mode {mode}] {nm} fwd bias')
            timeit(lambda: ops.linear_dgrad(dy, w), fl, f'[pair mode {mode}] {nm} dgrad')
            timeit(lambda: ops.linear_wgrad(dy, a, wp, None), fl, f'[pair mode {mode}] {nm} wgrad')
    lib.s4_set_tc_pair_mode(1)
if what in ('gemm', 'all'):
    x = torch.randn(M, D, device=dev).to(BF)
    for N, K, nm in ((3 * D, D, 'qkv'), (D, D, 'out_proj'), (4 * D, D, 'fc1'), (D, 4 * D, 'fc2')):
        a = torch.randn(M, K, device=dev).to(BF)
        w = torch.randn(N, K, device=dev).to(BF)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev).to(BF)
        fl = 2.0 * M * N * K
        timeit(lambda: ops.linear_fwd(a, w, bias), fl, f'{nm} fwd bias')
        if nm == 'fc1':
            timeit(lambda: ops.linear_fwd(a, w, bias, act=L.ACT_GELU, want_pre=True), fl, f'{nm} fwd bias+gelu+pre')
        if nm in ('out_proj', 'fc2'):
            timeit(lambda: ops.linear_fwd(a, w, bias, res=res), fl, f'{nm} fwd bias+res')
        dy = torch.randn(M, N, device=dev).to(BF)
        wt = w.t().contiguous()
        timeit(lambda: ops.linear_dgrad(dy, w), fl, f'{nm} dgrad')
        wp = torch.nn.Parameter(torch.randn(N, K, device=dev))
        bp = torch.nn.Parameter(torch.randn(N, device=dev))
        timeit(lambda: ops.linear_wgrad(dy, a, wp, bp), fl, f'{nm} wgrad(+bias colsum)')
