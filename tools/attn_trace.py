"""Device-side event timeline of one CTA of the fused attention kernels (debug instantiation).
Usage: python tools/attn_trace.py [fwd|fwd_bias|bwd] [block] ; env S4_BENCH_B (default 24)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from s4former_b200 import ops, _lib as L

what = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
B, H, Lt, hd = int(os.environ.get('S4_BENCH_B', '24')), 12, 1025, 64
block = int(sys.argv[2]) if len(sys.argv) > 2 else 101
dev = 'cuda'
lib = L.load()
D = H * hd
qkv = (torch.randn(B * Lt, 3 * D, device=dev) * 0.5).to(torch.bfloat16)
u0 = torch.rand(B, Lt, device=dev)
gate = (torch.rand(B, Lt, device=dev) > 0.5).float()
out, lse = ops.attention_fwd(qkv, B, Lt, H, hd, None, None, 0.0)
dout = torch.randn(B * Lt, D, device=dev).to(torch.bfloat16)
ops.attention_bwd(dout, qkv, out, lse, B, Lt, H, hd, None, None, 0.0)
torch.cuda.synchronize()
import ctypes as C


tmax = lib.s4_attention_set_trace(None, 0)
buf = torch.zeros(4 * tmax * 2, dtype=torch.int64, device=dev)
lib.s4_attention_set_trace(buf.data_ptr(), block)
if what == 'fwd':
    ops.attention_fwd(qkv, B, Lt, H, hd, None, None, 0.0)
elif what == 'fwd_bias':
    ops.attention_fwd(qkv, B, Lt, H, hd, u0, gate, 5.0)
else:
    ops.attention_bwd(dout, qkv, out, lse, B, Lt, H, hd, None, None, 0.0)
torch.cuda.synchronize()
lib.s4_attention_set_trace(None, 0)
t = buf.cpu().view(4, tmax, 2)
evs = []
for role in range(4):
    for k in range(tmax):
        if int(t[role, k, 0]) == 0:
            break
        evs.append((int(t[role, k, 1]), role, int(t[role, k, 0])))
evs.sort()
t0 = evs[0][0]
print(f'{what} B={B} block={block}: {len(evs)} events, span {evs[-1][0] - t0} clk')
last = {}
cols = ['MMA', 'SM0', 'SM1', 'EPI']
for ts, role, eid in evs:
    d = ts - last.get(role, ts)
    last[role] = ts
    print(f'{ts - t0:8d}  ' + '            ' * role + f'{cols[role]}:{eid:<3d}(+{d})')
