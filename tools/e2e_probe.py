"""Where does the end-to-end step lose time against the resident one?  Times variants of the host
loop around TrainStep (CUDA events, 10 steps each)."""
import os, sys, time, warnings
warnings.filterwarnings('ignore')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s4former_b200 as s4
from s4former_b200 import configs, ops
from s4former_b200.runner import TrainStep
from s4former_b200.utils.synthetic import make_batch, fresh_metas
dev = torch.device('cuda', 0)
torch.manual_seed(1999)
m = s4.build_segmentor(configs.setr_pup_deit_base('ours', 512, 21, norm='SyncBN'))
m.init_weights()
m.backbone_ema.load_state_dict(m.backbone.state_dict()); m.decode_head_ema.load_state_dict(m.decode_head.state_dict())
m = m.to(dev).train()
step = TrainStep(m, cuda_graph=True, graph_warmup=2)
img, gt, metas = make_batch(8, 8, 512, 21, seed=1999)
img_h = [img.pin_memory(), img.clone().pin_memory()]; gt_h = [gt.pin_memory(), gt.clone().pin_memory()]
img_d, gt_d = img.to(dev), gt.to(dev)
for i in range(4):
    step(img_d, fresh_metas(metas), gt_d, i, sync=False)
torch.cuda.synchronize()
def timed(name, fn, k=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(k): fn(i)
    e1.record(); th = (time.perf_counter() - t0) * 1e3 / k
    torch.cuda.synchronize()
    print(f'{name:50s} {e0.elapsed_time(e1) / k:7.2f} ms/step (host enqueue {th:6.2f} ms)', flush=True)
timed('resident replay', lambda i: step(img_d, fresh_metas(metas), gt_d, 10 + i, sync=False))
def e2e(i, k=10, defer=True):
    if i + 1 < k: step.prefetch(img_h[(i + 1) & 1], gt_h[(i + 1) & 1])
    loss, p = step.step_from_host(img_h[i & 1], fresh_metas(metas), gt_h[i & 1], 30 + i, deferred=True)
    pend.append(p)
    if len(pend) > 1: pend.pop(0)()
pend = []
timed('e2e (prefetch + deferred logs)', e2e)
while pend: pend.pop(0)()
cs = torch.cuda.Stream()
buf = torch.empty_like(img_d)
def copy_only(i):
    with torch.cuda.stream(cs):
        buf.copy_(img_h[i & 1], non_blocking=True)
timed('H2D copies only (side stream)', copy_only); cs.synchronize()
def res_plus_copy(i):
    with torch.cuda.stream(cs):
        buf.copy_(img_h[i & 1], non_blocking=True)
    step(img_d, fresh_metas(metas), gt_d, 50 + i, sync=False)
timed('resident replay + concurrent unrelated H2D', res_plus_copy); cs.synchronize()
def res_d2d(i):
    buf.copy_(img_d, non_blocking=True)
    step(img_d, fresh_metas(metas), gt_d, 70 + i, sync=False)
timed('resident replay + 126 MB D2D on the same stream', res_d2d)

pend = []
def e2e_nolog(i, k=10):
    if i + 1 < k: step.prefetch(img_h[(i + 1) & 1], gt_h[(i + 1) & 1])
    loss, p = step.step_from_host(img_h[i & 1], fresh_metas(metas), gt_h[i & 1], 90 + i, deferred=True)
    pend.append(p)
timed('e2e, logs never read inside the loop', e2e_nolog)
while pend: pend.pop(0)()
def e2e_noprefetch(i, k=10):
    loss, p = step.step_from_host(img_h[i & 1], fresh_metas(metas), gt_h[i & 1], 110 + i, deferred=True)
    pend.append(p)
    if len(pend) > 1: pend.pop(0)()
timed('e2e, no prefetch (copy enqueued with its step)', e2e_noprefetch)
while pend: pend.pop(0)()
def e2e_late(i, k=10):
    loss, p = step.step_from_host(img_h[i & 1], fresh_metas(metas), gt_h[i & 1], 130 + i, deferred=True)
    if i + 1 < k: step.prefetch(img_h[(i + 1) & 1], gt_h[(i + 1) & 1])
    pend.append(p)
    if len(pend) > 1: pend.pop(0)()
step.prefetch(img_h[0], gt_h[0])
timed('e2e, prefetch issued AFTER enqueueing the step', e2e_late)
while pend: pend.pop(0)()
