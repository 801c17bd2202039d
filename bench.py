#!/usr/bin/env python
"""S4Former train-step benchmark (BASELINE.json metric: train steps/s, 8 labeled + 8 unlabeled
512x512 crops per GPU, DeiT-B SETR-PUP, fraction of the bf16 tensor-core roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle)

One step = EMA update, supervised pass (backbone + main head + 4 aux heads), EMA-teacher forward,
pseudo labels, PASA student pass, CutMix/PatchShuffle student pass, masked CE + NCR, backward,
gradient all-reduce (N > 1), SGD-momentum step.  Synthetic data, random-init weights.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--variant', default='ours', choices=['ours', 'mt', 'sup'])
    ap.add_argument('--sup', type=int, default=8, help='labeled crops per GPU per step')
    ap.add_argument('--unsup', type=int, default=8, help='unlabeled crops per GPU per step')
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--classes', type=int, default=21)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-prof', action='store_true')
    ap.add_argument('--cpu-budget-s', type=float, default=150.0, help='reference arm wall budget')
    ap.add_argument('--cpu-threads', type=int, default=0, help='reference arm threads (0 = all host cores)')
    ap.add_argument('--no-gpu-eager', action='store_true',
                    help='skip the PyTorch-eager-on-this-GPU leg (the oracle math on cuBLAS/cuDNN kernels)')
    ap.add_argument('--no-extra-configs', action='store_true',
                    help='skip the short runs of BASELINE configs 1/2/4 (sup-only, MT, 768x768/19)')
    ap.add_argument('--no-parity', action='store_true', help='skip the in-bench parity block')
    ap.add_argument('--bucket-mb', type=float, default=25.0, help='gradient all-reduce bucket size (N > 1)')
    ap.add_argument('--norm', default='SyncBN', choices=['SyncBN', 'BN'],
                    help='BN = per-rank statistics (diagnostic: isolates the cost of the SyncBN collectives)')
    ap.add_argument('--no-graph', action='store_true',
                    help='launch every kernel from the host each step instead of replaying the captured CUDA graph')
    return ap.parse_args()


def workload_name(a):
    names = dict(ours='S4Former full (MT + PASA + NCR + CutMix/PatchShuffle)', mt='Mean Teacher as shipped',
                 sup='supervised only')
    return (f'{names[a.variant]}: {a.sup} labeled + {a.unsup if a.variant != "sup" else 0} unlabeled '
            f'{a.size}x{a.size} crops per GPU, {a.classes} classes, DeiT-B SETR-PUP '
            f'(BASELINE.json configs[{dict(ours=2, mt=1, sup=0)[a.variant]}] step at the metric\'s 8L+8U batch)')


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                 '-i', str(self.index)], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=max(pw))
        return out


# --------------------------------------------------------------------------------------------
# CPU reference path (the oracle: the reference's algorithm in plain PyTorch fp32)
# --------------------------------------------------------------------------------------------
EMA_CLS_SCALE = 50.0      # random-init conv_seg ~ N(0, 0.01): scale the teacher's so ~half the pixels clear 0.95


def cpu_reference_steps(a, n_sup, n_unsup, max_steps, warmup, budget_s, keep=None):
    """Times the oracle's train step (forward_train + backward + SGD step) on the host cores for a
    BOUNDED sample (n_sup labeled + n_unsup unlabeled crops).  Returns (seconds/step, steps, threads).
    ``keep`` (a dict) receives the initial weights, the batch and the first step's losses / pseudo
    labels for the in-bench parity block."""
    import copy
    import torch
    # The launcher's environment must not decide the baseline's speed: torch.distributed.run
    # exports OMP_NUM_THREADS=1, which would time the reference on ONE core.  Always use all the
    # host cores (or --cpu-threads) and report the count.
    threads = a.cpu_threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    from oracle import s4former_oracle as O            # checker / CPU baseline only
    from s4former_b200 import configs
    from s4former_b200.utils.synthetic import make_batch, fresh_metas
    cfg = configs.setr_pup_deit_base(a.variant, a.size, a.classes, norm='BN')
    torch.manual_seed(1999)
    m = O.OracleEncoderDecoder(**{k: v for k, v in cfg.items() if k != 'type'})
    m.init_weights()
    if m.ema:
        m.backbone_ema.load_state_dict(m.backbone.state_dict())
        m.decode_head_ema.load_state_dict(m.decode_head.state_dict())
        with torch.no_grad():
            m.decode_head_ema.conv_seg.weight.mul_(EMA_CLS_SCALE)
            m.decode_head_ema.conv_seg.bias.mul_(EMA_CLS_SCALE)
    m.train()
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9)
    img, gt, metas = make_batch(n_sup, n_unsup if a.variant != 'sup' else 0, a.size, a.classes, seed=1999)
    if keep is not None:
        keep.update(state_dict={k: v.detach().clone() for k, v in m.state_dict().items()}, img=img, gt=gt,
                    metas=metas, cfg=cfg)
    O.seed_host_rng(1999)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + max_steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        rec = {} if (keep is not None and i == 0) else None
        losses = m.forward_train(img, fresh_metas(metas), gt, record=rec)
        if rec is not None:
            keep.update(losses={k: float(v) for k, v in losses.items()},
                        mask_ratio=float(rec['conf'].float().mean()) if 'conf' in rec else None)
        loss = O.parse_losses(losses)
        loss.backward()
        opt.step()
        loss.item()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        # keep the whole run inside the budget (always at least one timed step)
        if times and (time.perf_counter() - t_start) + dt > budget_s:
            break
    times.sort()
    return times[len(times) // 2], len(times), threads


def reference_arm(a):
    """``--impl reference``: the reference's own algorithm (the oracle port, pinned to the reference
    by tests/golden) on the host cores.  One "step" of this arm is a BOUNDED SAMPLE of the workload
    (2 labeled + 2 unlabeled crops = 1/4 of the 8L+8U batch): ``steps`` / ``ms_per_step`` describe the
    sample steps actually timed; ``value`` is the metric (full 8L+8U steps/s) = sample_fraction /
    seconds-per-sample-step, with the extrapolation spelled out in ``config``."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    s_sup = min(2, a.sup)
    s_unsup = min(2, a.unsup) if a.variant != 'sup' else 0
    sec, n, threads = cpu_reference_steps(a, s_sup, s_unsup, max(1, a.steps), min(a.warmup, 1), a.cpu_budget_s)
    frac = s_sup / a.sup
    value = frac / sec                        # full (8L+8U) steps per second
    unit = 'steps/s'
    sample = (f'{s_sup} labeled + {s_unsup} unlabeled {a.size}x{a.size} crops per sample step '
              f'({frac:.3g} of the {a.sup}L+{a.unsup}U batch), fp32, PyTorch CPU kernels on {threads} threads, '
              f'oracle port of the reference path; median of {n} timed sample step(s)')
    out = dict(impl='reference', metric='train_steps_per_s', value=value, unit=unit, n_gpus=a.gpus,
               gpus_used=0, steps=n, warmup=min(a.warmup, 1), ms_per_step=sec * 1e3, higher_is_better=True,
               scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=dict(workload=workload_name(a), sample=sample, sample_fraction=frac,
                           ms_per_full_step_extrapolated=sec * 1e3 / frac,
                           note='GPU-over-CPU context, not a kernel-quality claim; ms_per_step is one SAMPLE step'),
               cpu_baseline=dict(value=value, unit=unit, cores=threads, kind='port', sample=sample),
               e2e=dict(value=value, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    emit(out)


# --------------------------------------------------------------------------------------------
# legs that put the headline number in context (rank 0, N = 1 only)
# --------------------------------------------------------------------------------------------
def parity_block(a, keep, dev):
    """The GPU path (bf16 tcgen05 kernels) on the SAME weights / sample / host RNG seed as the
    cpu_baseline's oracle step: the 8 losses side by side."""
    import torch
    import s4former_b200 as s4
    from oracle import s4former_oracle as O
    from s4former_b200.utils.synthetic import fresh_metas
    cfg = dict(keep['cfg'])
    m = s4.build_segmentor(cfg)
    m.load_state_dict(keep['state_dict'])
    m = m.to(dev).train()
    m._topk_override = lambda flat: torch.topk(flat.detach().float().cpu(), int(0.5 * flat.size(-1)), dim=-1,
                                               largest=False)[1]      # the reference's CPU tie order
    O.seed_host_rng(1999)
    with torch.no_grad():
        pass
    losses = m.forward_train(keep['img'].to(dev), fresh_metas(keep['metas']), gt_semantic_seg=keep['gt'].to(dev), iter=0)
    got = {k: float(v) for k, v in losses.items()}
    rel = {k: abs(got[k] - v) / max(abs(v), 1e-12) for k, v in keep['losses'].items()}
    del m
    torch.cuda.empty_cache()
    return dict(what='losses of one step on the cpu_baseline sample: oracle fp32 on the host vs this repo in bf16 on the GPU '
                     '(same weights, inputs and host RNG seed)',
                oracle_cpu_fp32=keep['losses'], ours_gpu_bf16=got, max_rel_diff=max(rel.values()) if rel else None,
                teacher_mask_ratio_oracle=keep.get('mask_ratio'), tolerance='2e-2 (north_star bf16 gate)',
                ok=bool(rel) and max(rel.values()) <= 2e-2)


def gpu_eager_leg(a, dev, n_sup, n_unsup, steps=3):
    """SURVEY.md section 8(d) "the true bar": the reference path on THIS GPU -- the oracle port in
    PyTorch eager (cuBLAS / cuDNN / ATen kernels), fp32 (TF32 off, the reference's arithmetic) and
    under bf16 autocast, same batch, forward_train + backward + SGD step, CUDA events."""
    import torch
    from oracle import s4former_oracle as O            # baseline leg only: never the product path
    from s4former_b200 import configs
    from s4former_b200.utils.synthetic import make_batch, fresh_metas
    out = dict(kind='oracle port of the reference path in PyTorch eager on this GPU (cuBLAS/cuDNN/ATen)',
               steps=steps, workload=f'{n_sup} labeled + {n_unsup} unlabeled {a.size}x{a.size} crops')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = configs.setr_pup_deit_base(a.variant, a.size, a.classes, norm='BN')
    for mode in ('bf16_autocast', 'fp32'):
        try:
            torch.manual_seed(1999)
            m = O.OracleEncoderDecoder(**{k: v for k, v in cfg.items() if k != 'type'})
            m.init_weights()
            if m.ema:
                m.backbone_ema.load_state_dict(m.backbone.state_dict())
                m.decode_head_ema.load_state_dict(m.decode_head.state_dict())
                with torch.no_grad():
                    m.decode_head_ema.conv_seg.weight.mul_(EMA_CLS_SCALE)
                    m.decode_head_ema.conv_seg.bias.mul_(EMA_CLS_SCALE)
            m = m.to(dev).train()
            opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3, momentum=0.9)
            img, gt, metas = make_batch(n_sup, n_unsup, a.size, a.classes, seed=1999)
            img, gt = img.to(dev), gt.to(dev)
            O.seed_host_rng(1999)

            def one():
                opt.zero_grad(set_to_none=True)
                if mode == 'fp32':
                    losses = m.forward_train(img, fresh_metas(metas), gt)
                else:
                    with torch.autocast('cuda', dtype=torch.bfloat16):
                        losses = m.forward_train(img, fresh_metas(metas), gt)
                O.parse_losses(losses).backward()
                opt.step()
            one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = dict(ms_per_step=ms, steps_per_s=1e3 / ms, peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
        except Exception as e:      # e.g. out of memory at the full batch: say so instead of dying
            out[mode] = dict(error=f'{type(e).__name__}: {str(e)[:160]}')
        finally:
            m = opt = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(dev)
    return out


def extra_config_runs(a):
    """BASELINE.json configs 1 / 2 / 4 (supervised-only, Mean Teacher as shipped, 768x768 / 19 classes) as
    short runs of this same script, so that they are driver-run numbers too."""
    runs = {}
    specs = dict(sup_only_config1=['--variant', 'sup'], mean_teacher_config2=['--variant', 'mt'],
                 cityscapes_768_config4=['--size', '768', '--classes', '19'])
    for name, extra in specs.items():
        cmd = [sys.executable, os.path.abspath(__file__), '--steps', '5', '--warmup', '3', '--no-cpu-baseline',
               '--no-gpu-eager', '--no-extra-configs', '--no-prof', '--no-parity'] + extra
        if a.no_graph:
            cmd.append('--no-graph')
        try:
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=240)
            line = [ln for ln in r.stdout.decode().splitlines() if ln.startswith('{')][-1]
            j = json.loads(line)
            runs[name] = dict(workload=j['config']['workload'], ms_per_step=j['ms_per_step'], value=j['value'],
                              unit=j['unit'], e2e_ms_per_step=j['e2e']['ms_per_step'], step_tc_frac=j['step_tc_frac'],
                              step_tflops=j['step_tflops'], steps=j['steps'])
        except Exception as e:
            runs[name] = dict(error=f'{type(e).__name__}: {str(e)[:120]}')
    return runs


# --------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------
def calibrate_teacher(model, img_t, target=0.5):
    """Scale decode_head_ema.conv_seg so that ~`target` of the teacher's pixels clear the 0.95
    threshold (random-init conv_seg ~ N(0, 0.01) would leave every pixel unconfident and the
    masked CE / NCR branches idle).  Returns the achieved mask ratio."""
    import torch
    from s4former_b200 import ops
    with torch.no_grad():
        model.set_eval(True)
        z = model.decode_head_ema.forward_get_logits(model.extract_feat_ema(img_t), None, [dict()])
        zs = z[:, :, ::4, ::4].float()
        lo, hi = 1.0, 1e6
        for _ in range(40):
            mid = (lo * hi) ** 0.5
            r = float((torch.softmax(zs * mid, 1).max(1)[0] > 0.95).float().mean())
            if r < target:
                lo = mid
            else:
                hi = mid
        s = (lo * hi) ** 0.5
        model.decode_head_ema.conv_seg.weight.mul_(s)
        model.decode_head_ema.conv_seg.bias.mul_(s)
        ops.bump_generation(model.decode_head_ema.conv_seg.weight)
        z = model.decode_head_ema.forward_get_logits(model.extract_feat_ema(img_t), None, [dict()])
        _, conf, _ = ops.pseudo_label(z, 0.95, 16)
        model.set_train(True)
        return float(conf.float().mean())


_REAL_STDOUT = None


def _quiet_stdout():
    """Route fd 1 to stderr while the benchmark runs (NCCL prints its version banner on stdout);
    the ONE JSON line is written to the real stdout at the end."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(obj):
    line = json.dumps(obj)
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(line + '\n')
    out.flush()


def main():
    a = parse_args()
    if a.impl == 'reference':
        return reference_arm(a)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if a.gpus > 1 and world == 1:      # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={a.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    _quiet_stdout()
    import warnings
    warnings.filterwarnings('ignore')
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py (impl=ours) needs a GPU: there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import s4former_b200 as s4
    from s4former_b200 import _lib, configs, ops
    from s4former_b200.runner import TrainStep
    from s4former_b200.utils.synthetic import make_batch, fresh_metas
    lib = _lib.load()
    ops.set_compute_dtype(torch.bfloat16 if a.dtype == 'bf16' else torch.float32)

    n_unsup = a.unsup if a.variant != 'sup' else 0
    cfg = configs.setr_pup_deit_base(a.variant, a.size, a.classes, norm=a.norm)
    torch.manual_seed(1999)
    model = s4.build_segmentor(cfg)
    model.init_weights()
    model.backbone_ema.load_state_dict(model.backbone.state_dict())
    model.decode_head_ema.load_state_dict(model.decode_head.state_dict())
    model = model.to(dev).train()
    step = TrainStep(model, cuda_graph=not a.no_graph, graph_warmup=2, bucket_mb=a.bucket_mb)

    img, gt, metas = make_batch(a.sup, n_unsup, a.size, a.classes, seed=1999 + rank)
    img_h, gt_h = img.pin_memory(), gt.pin_memory()
    img_h2, gt_h2 = [img_h, img.clone().pin_memory()], [gt_h, gt.clone().pin_memory()]   # two pinned batches
    img_d, gt_d = img_h.to(dev), gt_h.to(dev)
    mask_ratio = None
    if n_unsup:
        t_sel = [i for i, m in enumerate(metas) if m['tag'] == 'unsup_teacher']
        mask_ratio = calibrate_teacher(model, img_d[t_sel])
    import numpy as np
    import random
    random.seed(1999 + rank)
    np.random.seed(1999 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, k, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for i in range(k):
            fn(i)
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / k     # host enqueue time per step
        if after is not None:
            after()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    it = [0]

    def run_resident(_):
        step(img_d, fresh_metas(metas), gt_d, it[0], sync=False)
        it[0] += 1

    last = {}

    pend = []

    def run_e2e(i, k=None):
        # the dataloader's next pinned batch is copied on a side stream while this step runs (every
        # step's H2D copy is issued inside the timed region; the last one prefetches nothing) and
        # the step's log variables are read back one step late (every step's D2H read happens too)
        if k is not None and i + 1 < k:
            step.prefetch(img_h2[(i + 1) & 1], gt_h2[(i + 1) & 1])
        loss, p_ = step.step_from_host(img_h2[i & 1], fresh_metas(metas), gt_h2[i & 1], it[0], deferred=True)
        pend.append(p_)
        if len(pend) > 1:
            last.update(pend.pop(0)())
        it[0] += 1

    def drain_e2e():
        while pend:
            last.update(pend.pop(0)())

    for i in range(max(a.warmup, 3)):
        run_resident(i)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = lib.s4_launch_count()
    r0 = step.replays
    ms = timed(run_resident, a.steps)
    host_enqueue_ms = host_ms[0]
    graph_replays = step.replays - r0
    if graph_replays == a.steps:      # kernel nodes of the captured step (counted while capturing)
        launches = step.graph_kernel_launches
    else:
        launches = (lib.s4_launch_count() - l0) // a.steps
    for i in range(3):          # both staging slots warm (each has its own captured graph)
        run_e2e(i, 3)
    drain_e2e()
    ms_e2e = timed(lambda i: run_e2e(i, a.steps), a.steps, after=drain_e2e)
    clk = clocks.stop() if rank == 0 else None
    loss_val = last.get('loss')
    assert loss_val == loss_val, 'loss is NaN'

    # ---- roofline leg: per-kernel-family device time, measured live with CUDA events ---------
    kinds = []
    if not a.no_prof:
        use_graph, step.cuda_graph = step.cuda_graph, False     # per-launch event pairs need host launches
        lib.s4_prof_enable(1)
        nprof = 2
        for i in range(nprof):
            run_resident(i)
        torch.cuda.synchronize()
        step.cuda_graph = use_graph
        import ctypes as C
        for i in range(lib.s4_prof_num_kinds()):
            name = C.create_string_buffer(64)
            unit, kms, work, n = C.c_int(), C.c_double(), C.c_double(), C.c_longlong()
            lib.s4_prof_get(i, name, 64, C.byref(unit), C.byref(kms), C.byref(work), C.byref(n))
            kinds.append(dict(name=name.value.decode(), unit='flop' if unit.value == 0 else 'byte',
                              ms_per_step=kms.value / nprof, work_per_step=work.value / nprof,
                              launches_per_step=n.value / nprof))
        lib.s4_prof_enable(0)

    def teardown():
        """Captured NCCL collectives keep the communicator busy: destroying the process group
        with a live CUDA graph hung the N > 1 run at exit.  Drop the graph, drain, and leave
        without the collective teardown."""
        if step is not None:
            step._graph = None
            step._graphs = {}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            if _REAL_STDOUT is not None:
                _REAL_STDOUT.flush()
            os._exit(0)

    if rank != 0:
        teardown()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    tc_peak = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback 1.4 PFLOP/s sustained'
    flops = configs.step_flops(a.variant, a.sup, n_unsup, a.size, a.classes)
    sec = ms / 1e3 / a.steps
    sec_e2e = ms_e2e / 1e3 / a.steps
    value = world / sec
    roofline = dict(bound='tensor', achieved=None, peak=tc_peak, unit='TFLOP/s', frac=None, traffic=None,
                    peak_source=peak_src)
    # the dominant kernel is gemm_tc_kernel: the linear-layer GEMMs and the implicit-GEMM 3x3
    # convolutions (forward / dgrad / wgrad) are launches of the same tcgen05 kernel
    fam = ('gemm_tc', 'conv3x3_tc', 'conv3x3_wgrad_tc')
    tcs = [k for k in kinds if k['name'] in fam and k['work_per_step'] > 0 and k['ms_per_step'] > 0]
    if tcs:
        work = sum(k['work_per_step'] for k in tcs)
        kms = sum(k['ms_per_step'] for k in tcs)
        nl = sum(k['launches_per_step'] for k in tcs)
        ach = work / (kms * 1e-3) / 1e12
        roofline.update(kernel='gemm_tc_kernel (linear GEMMs + implicit-GEMM 3x3 conv fwd/dgrad/wgrad)',
                        achieved=ach, frac=ach / tc_peak, launches_per_step=nl, avg_launch_ms=kms / nl,
                        flop_per_launch=work / nl, share_of_step=kms / (sec * 1e3))
        try:      # DRAM bytes per launch of the same launches, from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_gemm_traffic.json')))
            roofline.update(traffic=tr['dram_bytes_per_launch'], traffic_unit='bytes/launch',
                            traffic_source=tr['source'])
        except Exception:
            pass
    out = dict(metric='train_steps_per_s', value=value, unit='steps/s', n_gpus=world, steps=a.steps,
               warmup=max(a.warmup, 3), ms_per_step=sec * 1e3, higher_is_better=True, scaling='weak',
               vs_baseline=None, dtype=a.dtype, data='synthetic',
               config=dict(workload=workload_name(a), parallelism=f'dp{world}',
                           images_per_step_per_gpu=a.sup + 2 * n_unsup,
                           l2='working set (activations > 10 GB per step) far exceeds the 126 MB L2; no flush needed',
                           teacher_mask_ratio=mask_ratio, optimizer='SGD momentum 0.9, poly LR, head lr x10 (fused)'),
               images_per_s=value * (a.sup + n_unsup),
               step_tflops=flops / sec / 1e12, step_tc_frac=flops / sec / 1e12 / tc_peak,
               flop_per_step=flops, loss=loss_val,
               e2e=dict(value=world / sec_e2e, unit='steps/s', ms_per_step=sec_e2e * 1e3,
                        h2d_bytes_per_step=img_h.numel() * 4 + gt_h.numel() * 8,
                        d2h_bytes_per_step=4 * len(last)),
               gpu_launches=int(launches) * a.steps, gpu_launches_per_step=int(launches),
               host_enqueue_ms_per_step=host_enqueue_ms,
               cuda_graph=dict(enabled=not a.no_graph, replays_in_timed_region=int(graph_replays),
                               kernel_nodes_per_replay=getattr(step, 'graph_kernel_launches', None)),
               clocks=clk, roofline=roofline,
               kernels=sorted(kinds, key=lambda k: -k["ms_per_step"])[:40])
    try:    # which multi-GPU mechanisms this run used
        from s4former_b200.parallel import PeerAllReduce
        peer = [v for v in PeerAllReduce._cache.values()]
        out['data_parallel'] = dict(
            gemm_tile_scheduler='dynamic' if lib.s4_set_tc_sched(-1) == 1 else 'static',
            syncbn_statistics=('nvlink_peer_memory' if any(v is not None for v in peer) else 'nccl')
            if world > 1 else None)
    except Exception as e:      # informational only
        out['data_parallel'] = dict(error=str(e))
    if world == 1 and not (a.no_gpu_eager and a.no_extra_configs and a.no_cpu_baseline):
        # the comparison legs need the memory: drop this arm's model, graph and staging buffers
        step._graph = None
        step._graphs = {}
        step = model = img_d = gt_d = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    if world == 1 and not a.no_gpu_eager:
        out['gpu_eager_baseline'] = gpu_eager_leg(a, dev, a.sup, n_unsup)
        for mode in ('bf16_autocast', 'fp32'):
            r = out['gpu_eager_baseline'].get(mode, {})
            if 'ms_per_step' in r:
                r['speedup_of_this_repo'] = r['ms_per_step'] / (sec * 1e3)
    if world == 1 and not a.no_extra_configs:
        out['extra_configs'] = extra_config_runs(a)
    if world == 1 and not a.no_cpu_baseline:
        s_sup, s_unsup = 1, (1 if n_unsup else 0)
        keep = {} if not a.no_parity else None
        csec, n, threads = cpu_reference_steps(a, s_sup, s_unsup, 1, 0, 40.0, keep=keep)
        if keep:
            try:
                out['parity'] = parity_block(a, keep, dev)
            except Exception as e:
                out['parity'] = dict(error=f'{type(e).__name__}: {str(e)[:160]}')
        out['cpu_baseline'] = dict(
            value=(s_sup / a.sup) / csec, unit='steps/s', cores=threads, kind='port',
            sample=f'{s_sup} labeled + {s_unsup} unlabeled {a.size}x{a.size} crops, one oracle step '
                   f'({csec:.1f} s), scaled by 1/{a.sup} to the {a.sup}L+{a.unsup}U step')
    emit(out)
    teardown()


if __name__ == '__main__':
    main()
