"""End-to-end parity of the S4Former train step (forward_train + backward + EMA) on the GPU
against (a) golden vectors produced by the unmodified reference and (b) the oracle run on the
same seeded inputs.  fp32 validation mode: losses/grads within 1e-3 relative; bf16 mode: 2e-2
(north_star tolerances); pseudo-label masks bit-exact given identical logits."""
import copy
import os
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu
warnings.filterwarnings('ignore')

import s4former_b200 as s4  # noqa: E402
from oracle import golden_common as gc  # noqa: E402
from oracle import s4former_oracle as O  # noqa: E402
from s4former_b200 import ops  # noqa: E402

DEV = 'cuda'


def _build(variant):
    m = s4.build_segmentor(gc.tiny_cfg(variant))
    sd = gc.seeded_state_dict(m.state_dict(), seed=5)
    m.load_state_dict(sd)
    return m.to(DEV).train(), sd


def _run(variant, dtype, topk=None, batched=True):
    ops.set_compute_dtype(dtype)
    try:
        m, sd = _build(variant)
        m.batch_student_passes = batched
        img, gt, metas = gc.tiny_batch(variant)
        O.seed_host_rng(1999)
        m._topk_override = topk
        losses = m.forward_train(img.to(DEV), metas, gt_semantic_seg=gt.to(DEV), iter=0)
        total, log_vars = m._parse_losses(losses)
        total.backward()
        torch.cuda.synchronize()
        return m, losses, log_vars, metas
    finally:
        ops.set_compute_dtype(torch.bfloat16)


def _golden(golden_dir, variant):
    return torch.load(os.path.join(golden_dir, f'step_{variant}.pt'), weights_only=False)


def _reference_topk(variant):
    """Top-k index set of the PASA gate as the CPU reference chose it (tie order is device
    specific, Appendix B-1): recomputed by the oracle on the same inputs."""
    if variant == 'sup':
        return None
    cfg = gc.tiny_cfg(variant)
    orc = O.OracleEncoderDecoder(**{k: v for k, v in cfg.items() if k != 'type'})
    orc.load_state_dict(gc.seeded_state_dict(orc.state_dict(), seed=5))
    orc.train()
    img, gt, metas = gc.tiny_batch(variant)
    O.seed_host_rng(1999)
    with torch.no_grad():
        O.ema_update(orc.backbone, orc.backbone_ema, orc.momentum)
        O.ema_update(orc.decode_head, orc.decode_head_ema, orc.momentum)
        sel = [i for i, mm in enumerate(metas) if mm['tag'] == 'unsup_teacher']
        orc.backbone_ema.eval()
        orc.decode_head_ema.eval()
        z = orc.decode_head_ema.forward(orc.backbone_ema(img[sel]))
        _, conf, _ = O.pseudo_label(z, 0.95)
        u = O.patch_unconfidence(conf, 16).reshape(len(sel), -1)
    return torch.topk(u, int(0.5 * u.shape[-1]), dim=-1, largest=False)[1]


@pytest.mark.parametrize('variant', ['sup', 'mt', 'ours'])
def test_train_step_fp32_vs_reference_golden(golden_dir, variant):
    G = _golden(golden_dir, variant)
    m, losses, log_vars, metas = _run(variant, torch.float32, topk=_reference_topk(variant))
    assert set(losses) == set(G['losses'])
    for k, v in G['losses'].items():
        got = float(losses[k])
        assert abs(got - float(v)) <= 1e-3 * abs(float(v)) + 1e-6, (k, got, float(v))
    named = dict(m.named_parameters())
    for k, g in G['grads'].items():
        r = float((named[k].grad.cpu() - g).norm() / (g.norm() + 1e-12))
        assert r < 1e-3, (k, r)
    for k, n in G['grad_norms'].items():
        got = float(named[k].grad.norm())
        assert abs(got - n) <= 2e-3 * n + 1e-8, (k, got, n)
    post = m.state_dict()
    for k, v in G['ema_after'].items():
        assert torch.allclose(post[k].float().cpu(), v.float(), rtol=1e-5, atol=1e-7), k
    for k, v in G['bn_after'].items():
        assert torch.allclose(post[k].cpu(), v, rtol=1e-3, atol=1e-5), k
    if variant == 'ours':
        sm = [mm for mm in metas if mm['tag'] == 'unsup_student']
        for mm, p in zip(sm, G['perms']):
            assert torch.equal(torch.as_tensor(mm['PatchMixIndex']), torch.as_tensor(p))


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_batched_student_passes_equal_pass_by_pass(golden_dir, dtype, tol):
    """The S4Former-full step runs its three student backbone passes as one batch; the
    reference's pass-by-pass order (batch_student_passes=False) must give the same losses,
    gradients, BN running statistics and PatchShuffle permutations.  The fp32 pass-by-pass run is
    also held to the reference goldens."""
    topk = _reference_topk('ours')
    ma, la, _, metas_a = _run('ours', dtype, topk=topk, batched=True)
    mb, lb, _, metas_b = _run('ours', dtype, topk=topk, batched=False)
    assert list(la) == list(lb)
    for k in la:
        a, b = float(la[k]), float(lb[k])
        assert abs(a - b) <= tol * abs(b) + 1e-6, (k, a, b)
    na, nb = dict(ma.named_parameters()), dict(mb.named_parameters())
    bad = []
    for k, pb in nb.items():
        if pb.grad is None:
            assert na[k].grad is None, k
            continue
        r = float((na[k].grad - pb.grad).norm() / (pb.grad.norm() + 1e-12))
        if r > (1e-4 if dtype == torch.float32 else 6e-2):
            bad.append((k, r))
    assert not bad, bad
    sa, sb = ma.state_dict(), mb.state_dict()
    for k in sb:
        if 'running_' in k:
            assert torch.allclose(sa[k], sb[k], rtol=1e-3 if dtype == torch.float32 else 3e-2, atol=1e-4), k
    for x, y in zip(metas_a, metas_b):
        if 'PatchMixIndex' in y:
            assert torch.equal(torch.as_tensor(x['PatchMixIndex']), torch.as_tensor(y['PatchMixIndex']))
    if dtype == torch.float32:
        G = _golden(golden_dir, 'ours')
        for k, v in G['losses'].items():
            assert abs(float(lb[k]) - float(v)) <= 1e-3 * abs(float(v)) + 1e-6, k
        for k, g in G['grads'].items():
            assert float((nb[k].grad.cpu() - g).norm() / (g.norm() + 1e-12)) < 1e-3, k


@pytest.mark.parametrize('variant', ['sup', 'ours'])
def test_train_step_bf16_vs_reference_golden(golden_dir, variant):
    G = _golden(golden_dir, variant)
    m, losses, log_vars, metas = _run(variant, torch.bfloat16, topk=_reference_topk(variant))
    for k, v in G['losses'].items():
        got = float(losses[k])
        assert abs(got - float(v)) <= 2e-2 * abs(float(v)) + 1e-3, (k, got, float(v))
    named = dict(m.named_parameters())
    # Gradients of this tiny, noisy problem: PyTorch's own bf16 autocast of the reference math
    # loses up to ~13% (stored per tensor in the golden file); the CUDA path must not be worse
    # than that yardstick (25% slack: the two bf16 roundings are independent noise of the same
    # size), and must meet 2e-2 where autocast itself does.
    bad = []
    for k, g in G['grads'].items():
        r = float((named[k].grad.cpu() - g).norm() / (g.norm() + 1e-12))
        if r > max(2e-2, 1.25 * G["bf16_autocast_err"][k]):
            bad.append((k, r, G['bf16_autocast_err'][k]))
    assert not bad, bad


def test_forward_logits_bf16_within_2e_2(golden_dir):
    """north_star: bf16 logits within 2e-2 relative of the fp32 reference."""
    G = _golden(golden_dir, 'ours')
    m, _ = _build('ours')
    m.eval()
    g2 = torch.Generator().manual_seed(G['vit_seed'])
    torch.rand(2, 8, 8, generator=g2)
    x = torch.randn(2, 3, 128, 128, generator=g2)
    with torch.no_grad():
        logits = m.decode_head.forward(m.backbone(x.to(DEV)))
    got, want = logits[:, :, ::2, ::2].float().cpu(), G['head_logits_eval']
    assert float((got - want).norm() / want.norm()) < 2e-2
    # argmax maps agree wherever the fp32 margin exceeds the bf16 error
    top2 = want.topk(2, dim=1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > 0.05 * want.abs().max()
    assert torch.equal(got.argmax(1)[safe], want.argmax(1)[safe])


def test_backbone_head_forward_fp32_vs_golden(golden_dir):
    G = _golden(golden_dir, 'ours')
    ops.set_compute_dtype(torch.float32)
    try:
        m, _ = _build('ours')
        m.eval()
        g2 = torch.Generator().manual_seed(G['vit_seed'])
        u = torch.rand(2, 8, 8, generator=g2).mul(16).round().div(16)
        x = torch.randn(2, 3, 128, 128, generator=g2)
        with torch.no_grad():
            feats = m.backbone(x.to(DEV), attn_mask=u.to(DEV), attn_mask_weight=5, adaptive_attn_mask=True,
                               topk_idx=G['vit_topk'])
            plain = m.backbone(x.to(DEV))
            logits = m.decode_head.forward(plain)
        for a, b in zip(feats, G['vit_feats']):
            assert torch.allclose(a.float().cpu(), b, rtol=1e-3, atol=1e-4)
        for a, b in zip(plain, G['vit_feats_plain']):
            assert torch.allclose(a.float().cpu(), b, rtol=1e-3, atol=1e-4)
        assert torch.allclose(logits[:, :, ::2, ::2].cpu(), G['head_logits_eval'], rtol=1e-3, atol=1e-3)
    finally:
        ops.set_compute_dtype(torch.bfloat16)


def test_missing_cuda_tensor_fails_loudly():
    m, _ = _build('sup')
    img, gt, metas = gc.tiny_batch('sup')
    with pytest.raises(Exception):
        m.cpu().forward_train(img, metas, gt_semantic_seg=gt, iter=0)


def test_train_step_runner_host_path_matches_resident():
    """TrainStep.step_from_host (side-stream prefetch into the staging slots, deferred log
    read-back) must produce exactly the numbers of the device-resident call, step after step,
    and num_batches_tracked must advance once per BatchNorm invocation (fused into bn_finalize)."""
    import copy as _copy
    from s4former_b200.runner import TrainStep

    def run(host):
        m, _ = _build('ours')
        step = TrainStep(m)
        img, gt, metas = gc.tiny_batch('ours')
        img_h, gt_h = img.pin_memory(), gt.pin_memory()
        img_h2, gt_h2 = img.clone().pin_memory(), gt.clone().pin_memory()
        hosts = [(img_h, gt_h), (img_h2, gt_h2)]
        O.seed_host_rng(1999)
        out, pend = [], []
        for it in range(3):
            if host:
                if it + 1 < 3:
                    step.prefetch(*hosts[(it + 1) & 1])
                _, p = step.step_from_host(hosts[it & 1][0], _copy.deepcopy(metas), hosts[it & 1][1], it, deferred=True)
                pend.append(p)
            else:
                _, lv = step(img.to(DEV), _copy.deepcopy(metas), gt.to(DEV), it, sync=True)
                out.append(lv)
        out += [p() for p in pend]
        torch.cuda.synchronize()
        nbt = [int(b) for n, b in m.named_buffers() if n.endswith('num_batches_tracked') and 'ema' not in n]
        return out, {n: p.detach().float().cpu().clone() for n, p in m.named_parameters()}, nbt

    a, pa, nbt_a = run(False)
    b, pb, nbt_b = run(True)
    assert len(a) == len(b) == 3
    # (fp32 atomics in the split-K weight gradients / BN statistics make two runs agree to rounding,
    # not bit for bit)
    # the first step sees identical weights and inputs; later steps inherit the bf16 rounding noise
    # of the previous update through a tiny, badly conditioned model
    for i, (la, lb) in enumerate(zip(a, b)):
        assert la.keys() == lb.keys()
        tol = 1e-3 if i == 0 else 5e-2
        for k in la:
            assert abs(la[k] - lb[k]) <= tol * abs(la[k]) + 1e-5, (i, k, la[k], lb[k])
    for k in pa:
        assert float((pa[k] - pb[k]).norm()) <= 1e-3 * float(pa[k].norm()) + 1e-6, k
    assert nbt_a == nbt_b and max(nbt_a) >= 3


@pytest.mark.parametrize('variant', ['ours', 'mt', 'sup'])
def test_cuda_graph_replay_equals_eager_steps(variant):
    """TrainStep(cuda_graph=True) captures the whole step (EMA, teacher, student passes, losses,
    backward, SGD) after two eager iterations and replays it; losses of every step and the weights
    after six steps must equal the all-eager run (same host RNG draws: boxes / permutations /
    learning rates reach the device through ops.StepParams in both modes)."""
    import copy as _copy
    from s4former_b200.runner import TrainStep

    def run(graph):
        m, _ = _build(variant)
        step = TrainStep(m, cuda_graph=graph, graph_warmup=2)
        img, gt, metas = gc.tiny_batch(variant)
        img_d, gt_d = img.to(DEV), gt.to(DEV)
        O.seed_host_rng(1999)
        logs, perms = [], []
        for it in range(6):
            mm = _copy.deepcopy(metas)
            _, lv = step(img_d, mm, gt_d, it, sync=True)
            logs.append(lv)
            perms.append([torch.as_tensor(x['PatchMixIndex']).clone() for x in mm if 'PatchMixIndex' in x])
        torch.cuda.synchronize()
        return logs, perms, {n: p.detach().float().cpu().clone() for n, p in m.named_parameters()}, step

    a, pa, wa, _ = run(False)
    b, pb, wb, step = run(True)
    assert step.replays == 4 and step.graph_kernel_launches > 50
    for i, (la, lb) in enumerate(zip(a, b)):
        assert list(la) == list(lb)
        tol = 1e-3 if i == 0 else 5e-2       # later steps inherit bf16 / atomic-order noise (tiny model)
        for k in la:
            assert abs(la[k] - lb[k]) <= tol * abs(la[k]) + 1e-5, (i, k, la[k], lb[k])
    for x, y in zip(pa, pb):                  # the reference's meta side effect survives the replay
        assert len(x) == len(y) and all(torch.equal(p, q) for p, q in zip(x, y))
    for k in wa:
        assert float((wa[k] - wb[k]).norm()) <= 2e-3 * float(wa[k].norm()) + 1e-6, k


def test_fused_sgd_ema_sweep_equals_separate_updates():
    """f1: TrainStep(fused_ema=True) applies the EMA-teacher update
    inside the SGD sweep of the SAME step; the reference does it at the start of the next
    iteration (encoder_decoder.py:416-423) on the same weights.  After n steps + the (n+1)-th
    step's EMA, student, teacher and BatchNorm statistics of both schedules must agree."""
    import copy as _copy
    from s4former_b200.runner import TrainStep

    def run(fused):
        m, _ = _build('ours')
        step = TrainStep(m, fused_ema=fused)
        img, gt, metas = gc.tiny_batch('ours')
        img_d, gt_d = img.to(DEV), gt.to(DEV)
        O.seed_host_rng(1999)
        logs = []
        for it in range(3):
            _, lv = step(img_d, _copy.deepcopy(metas), gt_d, it, sync=True)
            logs.append(lv)
        if not fused:      # bring the separate schedule to the same point: the next step's EMA
            with torch.no_grad():
                m.update_ema_variables(m.backbone, m.backbone_ema, m.momentum_backbone)
                m.update_ema_variables(m.decode_head, m.decode_head_ema, m.momentum_head)
        else:
            with torch.no_grad():
                m.update_ema_variables(m.backbone, m.backbone_ema, m.momentum_backbone, buffers_only=True)
                m.update_ema_variables(m.decode_head, m.decode_head_ema, m.momentum_head, buffers_only=True)
        torch.cuda.synchronize()
        return logs, {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}

    la, sa = run(False)
    lb, sb = run(True)
    for i, (x, y) in enumerate(zip(la, lb)):
        for k in x:
            assert abs(x[k] - y[k]) <= (1e-3 if i == 0 else 5e-2) * abs(x[k]) + 1e-5, (i, k, x[k], y[k])
    for k, v in sa.items():
        if v.dtype.is_floating_point and 'num_batches' not in k:
            tol = 2e-3 if 'ema' not in k else 1e-4      # the teacher moves by 1e-3 of the student's noise
            if 'running_' in k:                          # batch statistics of a tiny model amplify bf16 /
                tol = 2e-2 if 'ema' not in k else 1e-3   # atomic-order noise between two runs
            assert float((v - sb[k]).norm()) <= tol * float(v.norm()) + 1e-6, k


def test_sgd_ema_kernel_vs_torch_three_steps():
    """s4_sgd_ema_multi_tensor against torch.optim.SGD(momentum) followed by t = m t + (1-m) s."""
    from s4former_b200.optim import FusedSGD
    g = torch.Generator().manual_seed(3)
    shapes = [(300, 77), (5,), (16384 * 2 + 3,), (64, 3, 3, 3)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    es = [torch.nn.Parameter(p.detach().clone() + 0.1, requires_grad=False) for p in ps[:3]]
    ref_p = [p.detach().clone().requires_grad_(True) for p in ps]
    ref_e = [e.detach().clone() for e in es]
    opt = FusedSGD([(f'p{i}', p) for i, p in enumerate(ps)], lr=0.1, momentum=0.9)
    opt.attach_ema({id(ps[i]): (es[i], 0.99 if i else 0.9) for i in range(3)})
    ropt = torch.optim.SGD(ref_p, lr=0.1, momentum=0.9)
    for it in range(3):
        grads = [torch.randn(s, generator=g).to(DEV) for s in shapes]
        opt.zero_grad()
        for p, rp, gr in zip(ps, ref_p, grads):
            p.grad.copy_(gr)
            rp.grad = gr.clone()
        lrs = opt.current_lrs(it)
        for grp in ropt.param_groups:
            grp['lr'] = lrs[0]
        opt.step(it)
        ropt.step()
        for i in range(3):
            m = 0.99 if i else 0.9
            ref_e[i].mul_(m).add_(ref_p[i].detach(), alpha=1 - m)
    torch.cuda.synchronize()
    for p, rp in zip(ps, ref_p):
        assert torch.allclose(p.detach(), rp.detach(), rtol=1e-5, atol=1e-6)
    for e, re_ in zip(es, ref_e):
        assert torch.allclose(e.detach(), re_, rtol=1e-5, atol=1e-6)


def _data_batch(variant, it=0):
    """The reference's ``data_batch`` layout after ``collate(flatten=True)`` + scatter
    (datasets/builder.py:295-302): one dict of img / img_metas / gt_semantic_seg."""
    img, gt, metas = gc.tiny_batch(variant)
    return dict(img=img.to(DEV), img_metas=[dict(m) for m in metas], gt_semantic_seg=gt.to(DEV))


def test_train_step_and_val_step_api():
    """BaseSegmentor.train_step(data_batch, optimizer, iter=i) (base.py:155-206): returns
    dict(loss, log_vars, num_samples) with host-float log variables whose 'loss' is the sum of the
    entries containing 'loss'; val_step mirrors it."""
    m, _ = _build('ours')
    O.seed_host_rng(1999)
    out = m.train_step(_data_batch('ours'), None, iter=3)
    assert set(out) == {'loss', 'log_vars', 'num_samples'} and out['num_samples'] == 6
    lv = out['log_vars']
    assert all(isinstance(v, float) for v in lv.values())
    want_keys = {'decode.loss_ce', 'aux_0.loss_ce', 'aux_1.loss_ce', 'aux_2.loss_ce', 'aux_3.loss_ce',
                 'loss_seg_unsup_attn_mask', 'loss_ncr_unsup', 'loss_seg_unsup', 'loss'}
    assert set(lv) == want_keys
    assert abs(lv['loss'] - sum(v for k, v in lv.items() if 'loss' in k and k != 'loss')) < 1e-4 * abs(lv['loss'])
    assert abs(float(out['loss']) - lv['loss']) < 1e-5 * abs(lv['loss'])
    assert m.current_iter == 3
    out['loss'].backward()
    assert float(m.backbone.layers[0].ffn.layers[1].weight.grad.abs().sum()) > 0
    # the same numbers as the golden step (same seeds / weights / batch)
    G = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'step_ours.pt'), weights_only=False)
    for k, v in G['losses'].items():
        assert abs(lv[k] - float(v)) <= 3e-2 * abs(float(v)) + 1e-3, k
    O.seed_host_rng(1999)
    vout = m.val_step(dict(_data_batch('ours'), iter=4))
    assert set(vout['log_vars']) == {k + '_val' for k in want_keys}


def test_ddp_wrapper_world1_runner_contract():
    """S4DistributedDataParallel under the runner's protocol (mmcv OptimizerHook.after_train_iter:
    optimizer.zero_grad(); loss.backward(); optimizer.step()) with a stock torch.optim.SGD and
    zero_grad(set_to_none=True): two iterations equal the un-wrapped model driven the same way."""
    from s4former_b200.parallel import S4DistributedDataParallel

    def run(wrap):
        m, _ = _build('ours')
        model = S4DistributedDataParallel(m, device_ids=[0], broadcast_buffers=False,
                                          find_unused_parameters=False) if wrap else m
        opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3, momentum=0.9)
        O.seed_host_rng(1999)
        logs = []
        for it in range(2):
            if not wrap:
                opt.zero_grad(set_to_none=False)
                ops.reset_arena()
            out = model.train_step(_data_batch('ours'), opt, iter=it)
            if wrap:
                opt.zero_grad(set_to_none=True)          # after the forward, like the hook
            out['loss'].backward()
            opt.step()
            logs.append(out['log_vars'])
        torch.cuda.synchronize()
        if wrap:      # gradients still live in the wrapper's flat buffer
            fg = model.grads
            for p in fg.params:
                off, n = fg.offsets[id(p)]
                assert p.grad.data_ptr() == fg.flat.data_ptr() + 4 * off
        return logs, {n: p.detach().float().cpu().clone() for n, p in m.named_parameters()}

    la, wa = run(False)
    lb, wb = run(True)
    for x, y in zip(la, lb):
        for k in x:
            assert abs(x[k] - y[k]) <= 2e-2 * abs(x[k]) + 1e-4, (k, x[k], y[k])
    for k in wa:
        assert float((wa[k] - wb[k]).norm()) <= 2e-3 * float(wa[k].norm()) + 1e-6, k


def test_pasa_bias_only_on_peeled_key():
    """Regression for the peeled-key detection (L = 128 k + 1: the last key is folded in on the CUDA
    cores and is not part of the staged u0 row): a batch whose ONLY non-zero u0 entry is that last
    key must still take the biased path."""
    B, H, hd, L_ = 2, 2, 64, 257
    g = torch.Generator().manual_seed(4)
    qkv = (torch.randn(B * L_, 3 * H * hd, generator=g) * 0.5).to(DEV, torch.bfloat16)
    u0 = torch.zeros(B, L_)
    u0[0, L_ - 1] = 1.0                    # image 0: bias on the peeled key only; image 1: no bias at all
    gate = torch.ones(B, L_)
    gate[0, 5:40] = 0.0
    w = 5.0
    out, lse = ops.attention_fwd(qkv, B, L_, H, hd, u0.to(DEV), gate.to(DEV), w)
    q, k, v = qkv.float().view(B, L_, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    s = q @ k.transpose(-1, -2) / hd ** 0.5 + (w * gate.to(DEV)[:, None, :, None] * u0.to(DEV)[:, None, None, :])
    want = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * L_, H * hd)
    assert float((out.float() - want).norm() / want.norm()) < 2e-2
    # and it differs measurably from the unbiased result for image 0's gated-on rows
    plain, _ = ops.attention_fwd(qkv, B, L_, H, hd, None, None, 0.0)
    assert float((out.float()[:L_] - plain.float()[:L_]).norm()) > 1e-2 * float(plain.float()[:L_].norm())
    assert float((out.float()[L_:] - plain.float()[L_:]).norm()) < 1e-3 * float(plain.float()[L_:].norm())


def test_frozen_parameters_get_no_gradient_and_change_nothing_else():
    """requires_grad=False on an encoder layer / a head conv / a LayerNorm: their .grad stays None
    (autograd semantics; the weight-gradient GEMMs are skipped), every other gradient is unchanged."""
    def run(freeze):
        ops.set_compute_dtype(torch.float32)
        try:
            m, _ = _build('sup')
            frozen = []
            if freeze:
                for p in list(m.backbone.layers[1].parameters()) + list(m.decode_head.up_convs[0][0].conv.parameters()) \
                        + list(m.backbone.layers[0].ln1.parameters()):
                    p.requires_grad_(False)
                    frozen.append(id(p))
            img, gt, metas = gc.tiny_batch('sup')
            O.seed_host_rng(1999)
            losses = m.forward_train(img.to(DEV), metas, gt_semantic_seg=gt.to(DEV), iter=0)
            total, _ = m._parse_losses(losses)
            total.backward()
            torch.cuda.synchronize()
            return m, frozen
        finally:
            ops.set_compute_dtype(torch.bfloat16)
    m0, _ = run(False)
    m1, frozen = run(True)
    g0 = dict(m0.named_parameters())
    checked = 0
    for n, p in m1.named_parameters():
        if id(p) in frozen:
            assert p.grad is None, n
            continue
        if g0[n].grad is None:
            continue
        assert torch.allclose(p.grad, g0[n].grad, rtol=1e-4, atol=1e-7), n
        checked += 1
    assert frozen and checked > 50
