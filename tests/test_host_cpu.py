"""CPU tests of the host-side logic and of the C-ABI boundary (no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys
import warnings

import numpy as np
import pytest
import torch

warnings.filterwarnings('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='session')
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    from s4former_b200 import _lib
    return _lib


def test_abi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, 'include', 's4former.h')).read()
    declared = set(re.findall(r'\b(s4_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 35
    lib = built_lib.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert set(built_lib.EXPORTED_SYMBOLS) == declared
    assert lib.s4_version() >= 100 and lib.s4_built_arch() == 100


def test_abi_argument_counts_match_header(built_lib):
    hdr = open(os.path.join(ROOT, 'include', 's4former.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    for name, (_res, args) in built_lib._SPEC.items():
        m = re.search(r'\b' + name + r'\s*\(([^;]*?)\)\s*;', hdr, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        assert n == len(args), (name, n, len(args))


def test_gemm_params_struct_layout(built_lib):
    # 7 pointers, 5 ints (+pad), 11 long longs, float + 6 ints
    assert ctypes.sizeof(built_lib.GemmParams) == 7 * 8 + 5 * 4 + 4 + 11 * 8 + 7 * 4 + 4 + 8   # + colsum pointer


def test_library_is_sm100a_only():
    out = subprocess.run(['cuobjdump', '-lelf', os.path.join(ROOT, 's4former_b200', 'libs4former_b200.so')],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_sass_has_tcgen05_tma_and_no_serialised_mma_issue():
    """The hot kernels really are tcgen05 / TMEM / TMA code (SASS: UTCHMMA = tcgen05.mma, UTCBAR =
    tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store /
    reduce), the GEMM uses CTA pairs (.2CTA), and no tcgen05.mma sits inside one of ptxas'
    per-instruction elect loops (BRA.U.ANY right after it: what a `lane == 0` issue branch produces
    and what cost ~100 clk per MMA before the issue warps were made warp-uniform)."""
    import shutil
    lib = os.path.join(ROOT, 's4former_b200', 'libs4former_b200.so')
    if shutil.which('cuobjdump') is None or not os.path.exists(lib):
        pytest.skip('cuobjdump or the built library is not available')
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    for mnem in ('UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG',
                 'UBLKRED'):
        assert mnem in sass, mnem
    lines = [ln for ln in sass.splitlines() if '/*' in ln and ';' in ln]
    bad = 0
    for i, ln in enumerate(lines):
        if 'UTCHMMA' in ln:
            nxt = ' '.join(lines[i + 1:i + 3])
            if 'BRA.U.ANY' in nxt:
                bad += 1
    assert bad == 0, f'{bad} tcgen05.mma instructions are wrapped in serialising elect loops'


def test_product_package_never_imports_oracle():
    for dp, _dn, fns in os.walk(os.path.join(ROOT, 's4former_b200')):
        for fn in fns:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, fn)).read()
                assert 'oracle' not in src.lower(), os.path.join(dp, fn)


def test_ops_fail_loudly_without_cuda(built_lib):
    from s4former_b200 import ops
    with pytest.raises(built_lib.S4Error):
        ops.pseudo_label(torch.zeros(1, 3, 16, 32), 0.95)
    with pytest.raises(built_lib.S4Error):
        ops.cross_entropy(torch.zeros(1, 3, 16, 32), torch.zeros(1, 16, 32, dtype=torch.int64))


def test_registry_builds_reference_configs_and_state_dict_keys():
    import s4former_b200 as s4
    from oracle import golden_common as gc
    from oracle import s4former_oracle as O
    for variant in ('sup', 'mt', 'ours'):
        cfg = gc.tiny_cfg(variant)
        m = s4.build_segmentor(cfg)
        m.init_weights()
        o = O.OracleEncoderDecoder(**{k: v for k, v in cfg.items() if k != 'type'})
        sm, so = m.state_dict(), o.state_dict()
        assert set(sm) == set(so)
        for k in sm:
            assert sm[k].shape == so[k].shape, k
        for p in list(m.backbone_ema.parameters()) + list(m.decode_head_ema.parameters()):
            assert not p.requires_grad


def test_full_size_config_keys():
    """The DeiT-B SETR-PUP checkpoint layout of SURVEY.md section 8(b)."""
    import s4former_b200 as s4
    bb = s4.build_backbone(dict(type='VisionTransformer', img_size=(512, 512), patch_size=16, in_channels=3,
                                norm_cfg=dict(type='LN', eps=1e-6, requires_grad=True), with_cls_token=True,
                                interpolate_mode='bilinear', drop_rate=0., embed_dims=768, num_heads=12,
                                num_layers=12, out_indices=(4, 7, 9, 11)))
    sd = bb.state_dict()
    assert sd['patch_embed.projection.weight'].shape == (768, 3, 16, 16)
    assert sd['cls_token'].shape == (1, 1, 768) and sd['pos_embed'].shape == (1, 1025, 768)
    assert sd['layers.11.attn.attn.in_proj_weight'].shape == (2304, 768)
    assert sd['layers.0.ffn.layers.0.0.weight'].shape == (3072, 768)
    assert sd['layers.0.ffn.layers.1.weight'].shape == (768, 3072)
    assert sum(p.numel() for p in bb.parameters()) == 86_433_024
    head = s4.build_head(dict(type='SETRUPHead', align_corners=False, num_convs=4, in_channels=768, num_classes=21,
                              channels=256, in_index=3, dropout_ratio=0, norm_cfg=dict(type='SyncBN', requires_grad=True),
                              up_scale=2, kernel_size=3,
                              loss_decode=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0)))
    hs = head.state_dict()
    assert hs['up_convs.0.0.conv.weight'].shape == (256, 768, 3, 3)
    assert hs['up_convs.3.0.bn.running_var'].shape == (256,)
    assert hs['conv_seg.weight'].shape == (21, 256, 1, 1)


def test_unsupported_options_raise():
    import s4former_b200 as s4
    from oracle import golden_common as gc
    cfg = gc.tiny_cfg('ours')
    with pytest.raises(NotImplementedError):
        s4.build_segmentor(dict(cfg, unimatch=True))
    with pytest.raises(NotImplementedError):
        s4.build_backbone(dict(cfg['backbone'], final_norm=True))
    with pytest.raises(AssertionError):
        s4.build_backbone(dict(cfg['backbone'], output_cls_token=True, with_cls_token=False))
    with pytest.raises(KeyError):
        s4.build_backbone(dict(type='ResNet'))


def test_row_map_equals_reference_unshuffle(golden_dir):
    import s4former_b200 as s4
    G = torch.load(os.path.join(golden_dir, 'augment.pt'), weights_only=False)
    head = s4.SETRUPHead(in_channels=16, channels=8, num_classes=3, num_convs=1, up_scale=2, in_index=0,
                         dropout_ratio=0, norm_cfg=dict(type='BN'))
    rm = head._row_map(4, 8, False, 'cpu', 2, G['perms'])
    got = G['tok'].reshape(4 * 64, 16)[rm.long()].view(4, 64, 16)
    assert torch.equal(got, G['unsh'])
    # with a cls row in front of every image
    tok_cls = torch.cat([torch.full((4, 1, 16), -7.0), G['tok']], 1).reshape(4 * 65, 16)
    rm = head._row_map(4, 8, True, 'cpu', 2, G['perms'])
    assert torch.equal(tok_cls[rm.long()].view(4, 64, 16), G['unsh'])
    rm0 = head._row_map(4, 8, True, 'cpu', 0, None)
    assert torch.equal(tok_cls[rm0.long()].view(4, 64, 16), G['tok'])


def test_pasa_vectors_equal_oracle():
    import s4former_b200 as s4
    from oracle import s4former_oracle as O
    g = torch.Generator().manual_seed(3)
    u = torch.rand(3, 8, 8, generator=g).mul(16).round().div(16)
    u0, gate = s4.VisionTransformer.pasa_bias_vectors(u, True)
    ou0, ogate = O.pasa_gate_u0(u, True)
    assert torch.equal(u0, ou0) and torch.equal(gate, ogate)
    u0, gate = s4.VisionTransformer.pasa_bias_vectors(u, False)
    assert gate is None and torch.equal(u0, ou0)


def test_host_rng_order_equals_oracle():
    from oracle import s4former_oracle as O
    from s4former_b200.utils.generate_unsup_data import generate_cutout_box
    O.seed_host_rng(5)
    a = [O.cutout_box((512, 512), 2) for _ in range(4)]
    O.seed_host_rng(5)
    b = [generate_cutout_box([512, 512], 2) for _ in range(4)]
    assert a == b
    for (y0, y1, x0, x1) in a:
        assert 257 <= x1 - x0 <= 511 and 0 <= y0 < y1 <= 512


def test_dict_split_and_weighted_loss():
    from s4former_b200.utils.structual_utils import add_prefix, dict_split, weighted_loss
    img = torch.arange(6).float().view(6, 1)
    metas = [dict(tag=t, i=i) for i, t in enumerate(['sup', 'sup', 'unsup_student', 'unsup_teacher',
                                                     'unsup_student', 'unsup_teacher'])]
    groups = dict_split(dict(img=img, img_metas=metas, tag=[m['tag'] for m in metas]), 'tag')
    assert set(groups) == {'sup', 'unsup_student', 'unsup_teacher'}
    assert groups['unsup_student']['img'].view(-1).tolist() == [2., 4.]
    assert [m['i'] for m in groups['unsup_teacher']['img_metas']] == [3, 5]
    out = weighted_loss({'loss_a': torch.tensor(2.), 'mask_ratio': torch.tensor(3.)}, 0.5)
    assert float(out['loss_a']) == 1.0 and float(out['mask_ratio']) == 3.0
    assert add_prefix({'loss_ce': 1}, 'decode') == {'decode.loss_ce': 1}


def _parse_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from s4former_b200.segmentors.base import BaseSegmentor
    losses = {'decode.loss_ce': torch.tensor(1.0 + rank), 'aux_0.loss_ce': torch.tensor(0.5),
              'mask_ratio': torch.tensor(0.25 * (rank + 1))}
    loss, log_vars = BaseSegmentor._parse_losses(losses)
    q.put((rank, float(loss), dict(log_vars)))
    dist.destroy_process_group()


def test_parse_losses_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_parse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    (r0, l0, v0), (r1, l1, v1) = res
    assert l0 == 1.5 and l1 == 2.5            # the local loss drives backward
    assert v0 == v1                             # logged values are the cross-rank means
    assert abs(v0['decode.loss_ce'] - 1.5) < 1e-6 and abs(v0['loss'] - 2.0) < 1e-6
    assert abs(v0['mask_ratio'] - 0.375) < 1e-6


def _parse_deferred_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from s4former_b200.segmentors import base
    ok = {'decode.loss_ce': torch.tensor(1.0 + rank), 'mask_ratio': torch.tensor(0.5)}
    out = []
    # sync=False: values stay tensors, the cross-rank key-count assertion runs one call late
    _, lv = base.BaseSegmentor._parse_losses(dict(ok), sync=False)
    out.append(float(lv['decode.loss_ce']))
    assert len(base._pending_key_checks) == 1
    # a mismatch recorded by one call (here: forged, a real one would hang gloo's all_reduce) is
    # raised by the NEXT call
    cnt, want, names = base._pending_key_checks[0]
    base._pending_key_checks[0] = (cnt, want + 1, names)
    err = None
    try:
        base.BaseSegmentor._parse_losses(dict(ok), sync=False)
    except AssertionError:
        err = 'late'
    # and the queue keeps exactly one (fresh, good) entry afterwards
    base._pending_key_checks[:] = []
    base.BaseSegmentor._parse_losses(dict(ok), sync=False)
    base.BaseSegmentor._parse_losses(dict(ok), sync=False)
    assert len(base._pending_key_checks) == 1
    q.put((rank, out, err))
    try:
        dist.destroy_process_group()
    except Exception:
        pass


def test_parse_losses_deferred_key_check_two_ranks_gloo():
    """sync=False must not read the all-reduced key count in the same call (that is a mid-step
    host sync at N > 1) but must still catch a cross-rank mismatch, one call late."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_parse_deferred_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        if p.is_alive():
            p.terminate()
    for rank, out, err in res:
        assert abs(out[0] - 1.5) < 1e-6
        assert err == 'late', (rank, err)


def test_step_arena_and_bn_grad_add():
    """Host helpers behind the per-step scratch: arena_zeros hands out disjoint zeroed slices, a
    reset re-zeroes what was used; _add_bn_grads takes the one-launch path when weight.grad and
    bias.grad are adjacent in the flat gradient buffer and the two-add path otherwise."""
    from s4former_b200 import ops
    from s4former_b200.optim import FlatGrads
    dev = torch.device('cpu')
    ops._arena.pop(dev, None)
    a = ops.arena_zeros((2, 8), dev)
    b = ops.arena_zeros((3,), dev)
    assert float(a.abs().sum()) == 0 and float(b.abs().sum()) == 0
    a.fill_(1.0)
    b.fill_(2.0)
    assert a.data_ptr() != b.data_ptr() and float(a.sum()) == 16.0 and float(b.sum()) == 6.0
    ops.reset_arena()
    c = ops.arena_zeros((2, 8), dev)
    assert c.data_ptr() == a.data_ptr() and float(c.abs().sum()) == 0
    big = ops.arena_zeros((ops._ARENA_CHUNK + 5,), dev)       # larger than a chunk: own allocation
    assert big.numel() == ops._ARENA_CHUNK + 5 and float(big.abs().sum()) == 0
    ops._arena.pop(dev, None)

    bn = torch.nn.BatchNorm2d(4)
    FlatGrads(list(bn.parameters()))                          # weight.grad, bias.grad adjacent views
    sums = torch.arange(8, dtype=torch.float32).view(2, 4)    # [dgamma; dbeta]
    ops._add_bn_grads(bn, sums)
    ops._add_bn_grads(bn, sums)
    assert torch.equal(bn.weight.grad, 2 * sums[0]) and torch.equal(bn.bias.grad, 2 * sums[1])
    bn2 = torch.nn.BatchNorm2d(4)                             # separate gradient tensors: fallback
    ops._add_bn_grads(bn2, sums)
    assert torch.equal(bn2.weight.grad, sums[0]) and torch.equal(bn2.bias.grad, sums[1])


def test_init_weights_statistics():
    """Random-init path used by the parity tests and the bench (reference
    vit.py:369-414, setr_up_head.py:34-40, mmcv ConvModule default): trunc-normal(0.02) Linear /
    pos / cls, FFN biases ~ N(0, 1e-6), LayerNorm 1/0, conv_seg ~ N(0, 0.01), kaiming convs."""
    import s4former_b200 as s4
    from s4former_b200 import configs
    torch.manual_seed(3)
    m = s4.build_segmentor(configs.setr_pup_deit_base('ours', 512, 21, norm='BN'))
    m.init_weights()
    bb, head = m.backbone, m.decode_head
    lyr = bb.layers[3]
    # in_proj_weight is a bare Parameter of nn.MultiheadAttention (not an nn.Linear): the reference's
    # init loop never touches it, it keeps torch's xavier_uniform (bound sqrt(6 / (768 + 2304)))
    w = lyr.attn.attn.in_proj_weight
    bound = (6.0 / (768 + 2304)) ** 0.5
    assert float(w.abs().max()) <= bound + 1e-6 and abs(float(w.std()) - bound / 3 ** 0.5) < 1e-3
    wo = lyr.attn.attn.out_proj.weight                                               # an nn.Linear: trunc-normal
    assert abs(float(wo.std()) - 0.02) < 2e-3 and float(wo.abs().max()) <= 2.0   # cut at +-2 ABSOLUTE, as mmcv
    assert abs(float(bb.pos_embed.std()) - 0.02) < 2e-3
    fc1 = lyr.ffn.layers[0][0]
    assert abs(float(fc1.weight.std()) - 0.02) < 2e-3
    assert 0 < float(fc1.bias.abs().max()) < 1e-5                                      # N(0, 1e-6)
    assert float(lyr.attn.attn.out_proj.bias.abs().max()) == 0.0
    assert torch.equal(lyr.ln1.weight, torch.ones_like(lyr.ln1.weight)) and float(lyr.ln1.bias.abs().max()) == 0
    pe = bb.patch_embed.projection.weight                                             # kaiming, fan_in = 3*16*16
    assert abs(float(pe.std()) - (2.0 / (3 * 16 * 16)) ** 0.5) < 5e-3
    assert abs(float(head.conv_seg.weight.std()) - 0.01) < 2e-3 and float(head.conv_seg.bias.abs().max()) == 0
    c0 = head.up_convs[0][0].conv.weight                                              # kaiming, fan_out = 256*9
    assert abs(float(c0.std()) - (2.0 / (256 * 9)) ** 0.5) < 2e-3
    assert torch.equal(head.up_convs[0][0].bn.weight, torch.ones(256))


def test_parse_losses_single_process():
    from s4former_b200.segmentors.base import BaseSegmentor
    loss, lv = BaseSegmentor._parse_losses({'a.loss_ce': torch.tensor([1., 3.]), 'x': torch.tensor(5.)})
    assert float(loss) == 2.0 and lv['loss'] == 2.0 and lv['x'] == 5.0


def _reducer_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.nn as nn
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from s4former_b200.optim import FlatGrads
    from s4former_b200.parallel import GradReducer

    class SETRUPHead(nn.Module):          # name-matched stand-ins: the reducer keys on class names
        def __init__(self):
            super().__init__()
            self.conv = nn.Linear(8, 8)

    class TransformerEncoderLayer(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = nn.Linear(16, 16)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = nn.ModuleList([TransformerEncoderLayer() for _ in range(3)])
            self.head = SETRUPHead()
    torch.manual_seed(0)
    net = Net()
    fg = FlatGrads(list(net.parameters()))
    red = GradReducer(net, fg, bucket_bytes=600)     # several buckets
    assert len(red.buckets) >= 3
    for step in range(2):
        fg.zero()
        for i, p in enumerate(net.parameters()):
            p.grad.add_(float(rank + 1) * (i + 1) + step)
        red.module_ready(net.head)                   # backward order: head first, layers reversed
        red.module_ready(net.layers[2])
        red.module_ready(net.layers[1])
        # layers[0] never signals: finalize() must still reduce its bucket
        red.finalize()
        vals = [float(p.grad.flatten()[0]) for p in net.parameters()]
        want = [1.5 * (i + 1) + step for i in range(len(vals))]
        assert all(abs(a - b) < 1e-6 for a, b in zip(vals, want)), (vals, want)
    # SyncBN statistics: no peer-memory reducer without NCCL; the plain collective is used
    from s4former_b200.parallel import PeerAllReduce
    from s4former_b200 import ops
    assert PeerAllReduce.get(None) is None
    t = torch.full((2, 4), float(rank + 1))
    ops._all_reduce_stats(t, dict(world=world, group=None, peer=None))
    assert torch.equal(t, torch.full((2, 4), 3.0))
    q.put(rank)
    dist.destroy_process_group()


def test_grad_reducer_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert got == [0, 1]


def test_poly_lr_and_lr_mult():
    from s4former_b200.optim import poly_lr
    assert abs(poly_lr(1e-3, 0, 80000) - 1e-3) < 1e-12
    assert abs(poly_lr(1e-3, 80000, 80000) - 1e-4) < 1e-12
    assert abs(poly_lr(1e-2, 40000, 80000) - ((1e-2 - 1e-4) * 0.5 ** 0.9 + 1e-4)) < 1e-12


def test_row_map_covers_segformer_level_unshuffles():
    """The head's row-gather map (host index arithmetic, no cls row) reproduces the reference's
    ``_repatchmix_inputs`` at the per-level block sizes SegformerHead uses (PatchMix_N * 4 / 2^level
    tokens, segformer_head.py:167-170) — the building block a SegFormer path would reuse."""
    import s4former_b200 as s4
    from oracle import golden_common as gc
    from oracle import s4former_oracle as O
    head = s4.build_segmentor(gc.tiny_cfg('ours')).decode_head
    g0 = torch.Generator().manual_seed(4)
    B, N = 2, 4
    for level, g in enumerate((32, 16, 8, 4)):                # 128-pixel crop: strides 4 / 8 / 16 / 32
        n = int(N * (4 / (2 ** level)))
        gb = g // n
        perms = torch.stack([torch.randperm(gb * gb, generator=g0) for _ in range(B)])
        tok = torch.randn(B, g * g, 3, generator=g0)
        want = O.token_unshuffle(tok, perms, n)
        m = head._row_map(B, g, False, 'cpu', n, perms).long()
        got = tok.reshape(B * g * g, 3)[m].reshape(B, g * g, 3)
        assert torch.equal(got, want), (level, g, n)
