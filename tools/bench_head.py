"""Micro-benchmarks of the HBM-bound kernels (LayerNorm, SETR-PUP head element-wise stages, losses,
optimizer) at the train step's shapes: CUDA events on the launching stream, a 256 MB scratch write
between calls (L2 cold), achieved GB/s over the ALGORITHMIC bytes (each operand once) against the
measured copy bandwidth in MEASURED_PEAKS.json.

    python tools/bench_head.py [filter-substring] [reps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from s4former_b200 import ops, _lib as L  # noqa: E402
from s4former_b200.ops import _p, _st  # noqa: E402

dev = 'cuda'
BF = torch.bfloat16
flt = sys.argv[1] if len(sys.argv) > 1 else ''
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
scratch = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    PEAK = 6500.0
results = []


def timeit(name, fn, nbytes):
    if flt and flt not in name:
        return
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
    gbs = nbytes / med / 1e6
    results.append(dict(name=name, us=med * 1e3, mb=nbytes / 1e6, gbs=gbs, frac=gbs / PEAK))
    print(f'{name:46s} {med * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f}', flush=True)


def rnd(*shape, dtype=BF, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


# ---------------------------------------------------------------------------------------------- LN
D = 768
gamma, beta = torch.nn.Parameter(torch.rand(D, device=dev) + 0.5), torch.nn.Parameter(torch.randn(D, device=dev))
for rows in (24600, 8200):
    x = rnd(rows, D)
    timeit(f'ln_fwd rows={rows}', lambda: ops.layernorm_fwd(x, gamma, beta, 1e-6), rows * D * 4)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-6)
    dy, dres, dx = rnd(rows, D), rnd(rows, D), torch.empty_like(x)
    timeit(f'ln_bwd rows={rows} (+dres)', lambda: ops.layernorm_bwd(dy, x, gamma, beta, mean, rstd, dres=dres, dx=dx),
           rows * D * 8)
    rb = torch.nn.Parameter(torch.zeros(D, device=dev))
    timeit(f'ln_bwd rows={rows} (+dres, +colsum(dres))',
           lambda: ops.layernorm_bwd(dy, x, gamma, beta, mean, rstd, dres=dres, dx=dx, dres_bias=rb), rows * D * 8)
x = rnd(8 * 1025, D)
rm = (torch.arange(8 * 1024, device=dev, dtype=torch.int32) + torch.arange(8, device=dev, dtype=torch.int32).repeat_interleave(1024) + 1)
timeit('ln_fwd head row_map rows=8192', lambda: ops.layernorm_fwd(x, gamma, beta, 1e-6, row_map=rm, out_rows=8192), 8192 * D * 4)
for rows, cols in ((24600, 768), (24600, 2304)):
    x = rnd(rows, cols)
    out = torch.zeros(cols, device=dev)
    timeit(f'colsum {rows}x{cols}', lambda: L.call('s4_colsum', _p(x), _p(out), None, rows, cols, L.BF16, _st()), rows * cols * 2)

# ---------------------------------------------------------------------------------------------- head
C = 256
B = 8
sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
mean_c, invstd_c = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
gam = torch.rand(C, device=dev) + 0.5
sums = torch.zeros(2, C, device=dev)
for (H, s) in ((128, 2), (64, 2), (32, 2), (32, 4)):
    y = rnd(B * H * H, C)
    out = torch.empty(B * H * s * H * s, C, dtype=BF, device=dev)
    nb_in, nb_out = y.numel() * 2, out.numel() * 2
    timeit(f'bn_relu_upsample_fwd H={H} s={s}',
           lambda: L.call('s4_bn_relu_upsample_fwd', _p(y), _p(sc), _p(sh), _p(out), B, H, H, C, s, L.BF16, _st()),
           nb_in + nb_out)
    dout, dact = rnd(B * H * s * H * s, C), torch.empty_like(y)
    timeit(f'bn_relu_upsample_bwd H={H} s={s}',
           lambda: L.call('s4_bn_relu_upsample_bwd', _p(dout), _p(y), _p(sc), _p(sh), _p(mean_c), _p(invstd_c),
                          _p(dact), _p(sums[1]), _p(sums[0]), B, H, H, C, s, L.BF16, _st()),
           nb_out + 2 * nb_in)
    dyc = torch.empty_like(y)
    timeit(f'bn_bwd_apply rows={B * H * H}',
           lambda: L.call('s4_bn_bwd_apply', _p(dact), _p(y), _p(gam), _p(mean_c), _p(invstd_c), _p(sums[1]),
                          _p(sums[0]), float(B * H * H), _p(dyc), B * H * H, C, L.BF16, _st()),
           3 * nb_in)
NC = 21
w2, b2 = torch.randn(NC, C, device=dev) * 0.05, torch.randn(NC, device=dev) * 0.1
gw, gb = torch.zeros(NC, C, device=dev), torch.zeros(NC, device=dev)
for (H, s) in ((256, 2), (128, 4)):
    rows = B * H * H
    y = rnd(rows, C)
    z = torch.empty(rows, NC, device=dev)
    timeit(f'cls_fwd (bn_relu_conv1x1) H={H}',
           lambda: L.call('s4_bn_relu_conv1x1_fwd', _p(y), _p(sc), _p(sh), _p(w2), _p(b2), _p(z), rows, C, NC, L.BF16, _st()),
           rows * C * 2 + rows * NC * 4)
    logits = torch.empty(B, NC, H * s, H * s, device=dev)
    timeit(f'upsample_logits_fwd H={H} s={s}',
           lambda: L.call('s4_upsample_logits_fwd', _p(z), _p(logits), B, H, H, NC, s, _st()),
           rows * NC * 4 + logits.numel() * 4)
    dlog = torch.randn_like(logits)
    dz16 = torch.empty(rows, 32, dtype=BF, device=dev)
    timeit(f'cls_upsample_bwd H={H} s={s}',
           lambda: L.call('s4_cls_upsample_bwd_padded', _p(dlog), _p(dz16), B, H, H, NC, s, _st()),
           logits.numel() * 4 + rows * 64)
    timeit(f'cls_bwd_reduce H={H}',
           lambda: L.call('s4_cls_bwd_reduce', _p(dz16), _p(y), _p(sc), _p(sh), _p(mean_c), _p(invstd_c), _p(w2),
                          _p(gw), _p(gb), _p(sums[1]), _p(sums[0]), rows, C, NC, _st()),
           rows * C * 2 + rows * 64)
    dyc = torch.empty_like(y)
    timeit(f'cls_bwd_apply H={H}',
           lambda: L.call('s4_cls_bwd_apply', _p(dz16), _p(y), _p(sc), _p(sh), _p(mean_c), _p(invstd_c), _p(gam),
                          _p(w2), _p(sums[1]), _p(sums[0]), float(rows), _p(dyc), rows, C, NC, _st()),
           2 * rows * C * 2 + rows * 64)

# ---------------------------------------------------------------------------------------------- losses
Hh = 512
zs = torch.randn(B, NC, Hh, Hh, device=dev).requires_grad_(True)
zt = torch.randn(B, NC, Hh, Hh, device=dev) * 3
lab = torch.randint(0, NC, (B, Hh, Hh), device=dev)
lab[:, :26] = 255
npx = B * Hh * Hh
timeit('ce (fwd+grad) 8x21x512x512', lambda: ops.CeNcrFn.apply(zs, None, lab, 1.0, 0.0, 255), npx * (NC * 8 + 8))
timeit('ce+ncr (fwd+grad) 8x21x512x512', lambda: ops.CeNcrFn.apply(zs, zt, lab, 1.0, 1.0, 255), npx * (NC * 12 + 8))
timeit('pseudo_label 8x21x512x512', lambda: ops.pseudo_label(zt, 0.95, 16), npx * (NC * 4 + 16))

# ---------------------------------------------------------------------------------------------- misc
for cin in (256, 768):
    w = torch.randn(256, cin, 3, 3, device=dev)
    wf, wd = torch.empty(256, 9 * cin, dtype=BF, device=dev), torch.empty(cin, 9 * 256, dtype=BF, device=dev)
    timeit(f'pack_conv3x3_weight cin={cin}',
           lambda: L.call('s4_pack_conv3x3_weight', _p(w), _p(wf), _p(wd), cin, 256, L.BF16, _st()),
           w.numel() * 4 + 2 * w.numel() * 2)

os.makedirs('gpurun_out', exist_ok=True)
json.dump(dict(peak_gbs=PEAK, results=results), open('gpurun_out/bench_head.json', 'w'), indent=1)
