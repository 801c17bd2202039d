"""Actual relative error of the fused bf16 attention backward against fp32 math on the same bf16
inputs, at the shapes of tests/test_tc_gpu.py::test_tc_attention_bwd (the inherent part of that error
is evaluated on the CPU by tests/test_tolerance_yardsticks.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch  # noqa: E402
from s4former_b200 import ops  # noqa: E402
from test_tc_gpu import _attn_ref  # noqa: E402
from test_kernels_gpu import rel, gen  # noqa: E402

DEV, BF = 'cuda', torch.bfloat16
for B, H, Lt in [(2, 2, 65), (1, 3, 128), (2, 2, 200), (2, 2, 257), (1, 4, 1025), (1, 1, 2305)]:
    for pasa in (False, True):
        g = gen(13)
        hd = 64
        D = H * hd
        qkv = (torch.randn(B * Lt, 3 * D, generator=g) * 0.7).to(DEV, BF)
        dout = torch.randn(B * Lt, D, generator=g).to(DEV, BF)
        u0 = gate = None
        w = 0.0
        if pasa:
            u = (torch.rand(B, Lt - 1, generator=g) * 16).round() / 16
            u0 = torch.cat([torch.zeros(B, 1), u], 1)
            gate = (torch.rand(B, Lt, generator=g) > 0.5).float()
            gate[:, 0] = 1.0
            u0, gate, w = u0.to(DEV), gate.to(DEV), 5.0
        out, lse = ops.attention_fwd(qkv, B, Lt, H, hd, u0, gate, w)
        dqkv = ops.attention_bwd(dout, qkv, out, lse, B, Lt, H, hd, u0, gate, w)
        o_ref, _, g_ref = _attn_ref(qkv, B, Lt, H, hd, u0, gate, w, dout)
        gq, gk, gv = dqkv.float().cpu().view(B * Lt, 3, D).unbind(1)
        rq, rk, rv = g_ref.view(B * Lt, 3, D).unbind(1)
        print(f'B={B} H={H} L={Lt} pasa={int(pasa)}: O {rel(out.float().cpu(), o_ref):.2e}  dV {rel(gv, rv):.2e}  '
              f'dK {rel(gk, rk):.2e}  dQ {rel(gq, rq):.2e}', flush=True)
