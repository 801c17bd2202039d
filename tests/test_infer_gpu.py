"""Inference / validation path and the checkpoint pos_embed resize
against golden vectors produced by the unmodified reference files (oracle/make_golden_infer.py):
``encode_decode`` + whole / slide inference + rescale + softmax + flip + arg-max
(encoder_decoder.py:270-308, 1068-1232) and ``intersect_and_union`` / ``mean_iou``
(core/evaluation/metrics.py:26-165).  Arg-max maps must be bit-exact wherever the reference's own
top-2 probability margin exceeds the arithmetic tolerance of the mode (fp32 validation: 1e-4;
bf16: 2e-2); the class histograms are exact integers."""
import os
import warnings

import numpy as np
import pytest
import torch

warnings.filterwarnings('ignore')

import s4former_b200 as s4  # noqa: E402
from oracle import golden_common as gc  # noqa: E402
from oracle.make_golden_infer import infer_inputs  # noqa: E402
from s4former_b200 import ops  # noqa: E402
from s4former_b200.backbones.vit import VisionTransformer  # noqa: E402

DEV = 'cuda'


@pytest.fixture(scope='module')
def G(golden_dir):
    return torch.load(os.path.join(golden_dir, 'infer.pt'), weights_only=False)


def _pe():
    return torch.randn(1, 197, 24, generator=torch.Generator().manual_seed(41))


def test_resize_pos_embed_cpu_vs_reference(G):
    """vit.py:447-477 on the host (checkpoint load path)."""
    for (mode, hw), want in G['pos_embed'].items():
        got = VisionTransformer.resize_pos_embed(_pe(), hw, (14, 14), mode)
        assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-6, atol=1e-7), (mode, hw)


def test_pretrained_checkpoint_with_pos_embed_resize(G, tmp_path):
    """init_weights() with init_cfg=Pretrained (vit.py:369-395): a 14 x 14-grid checkpoint loads into a
    model with a 8 x 8 grid, pos_embed resized as the reference does (the shipped config's DeiT case)."""
    bb = dict(gc.tiny_cfg('sup')['backbone'])
    bb.pop('type')
    bb.update(embed_dims=24, num_heads=2, num_layers=1, out_indices=(0,))
    src = VisionTransformer(**bb)
    sd = {k: torch.randn(v.shape, generator=torch.Generator().manual_seed(3)) for k, v in src.state_dict().items()}
    sd['pos_embed'] = _pe()
    path = os.path.join(tmp_path, 'deit_like.pth')
    torch.save(dict(state_dict=sd), path)
    m = VisionTransformer(**dict(bb, pretrained=path))
    m.init_weights()
    want = VisionTransformer.resize_pos_embed(_pe(), (8, 8), (14, 14), 'bilinear')
    assert torch.equal(m.pos_embed.detach(), want)
    assert torch.equal(m.cls_token.detach(), sd['cls_token'])
    assert torch.equal(m.layers[0].attn.attn.in_proj_weight.detach(), sd['layers.0.attn.attn.in_proj_weight'])


@pytest.mark.gpu
def test_resize_bilinear_device_vs_reference(G):
    for (mode, hw), want in G['pos_embed'].items():
        if mode != 'bilinear':
            continue
        got = VisionTransformer.resize_pos_embed(_pe().to(DEV), hw, (14, 14), mode)
        assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-6), hw
    x = torch.randn(2, 3, 37, 53, generator=torch.Generator().manual_seed(5))
    for size in ((90, 96), (20, 31), (37, 53), (74, 106)):
        got = ops.resize_bilinear(x.to(DEV), size).cpu()
        want = torch.nn.functional.interpolate(x, size=size, mode='bilinear', align_corners=False)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), size


def _model(test_cfg, dtype):
    cfg = gc.tiny_cfg('ours')
    cfg['test_cfg'] = test_cfg
    m = s4.build_segmentor(cfg)
    m.load_state_dict(gc.seeded_state_dict(m.state_dict(), seed=5))
    ops.set_compute_dtype(dtype)
    return m.to(DEV).eval()


@pytest.mark.gpu
@pytest.mark.parametrize('case', [0, 1, 2])
@pytest.mark.parametrize('dtype,ptol,margin', [(torch.float32, 1e-3, 1e-4), (torch.bfloat16, 3e-2, 2e-2)])
def test_inference_vs_reference_golden(G, case, dtype, ptol, margin):
    want = G['infer'][case]
    img, metas, cfg = infer_inputs(case)
    assert abs(float(img.double().abs().sum()) - want['img_checksum']) < 1e-6 * want['img_checksum']
    try:
        m = _model(cfg, dtype)
        prob, pred = m.inference(img.to(DEV), metas, rescale=True, want_pred=True)
        seg = m.simple_test(img.to(DEV), metas, rescale=True)
        ft = m.forward_test([img.to(DEV)], [metas], rescale=True)
    finally:
        ops.set_compute_dtype(torch.bfloat16)
    assert prob.shape == want['prob'].shape
    err = float((prob.cpu() - want['prob']).abs().max())
    assert err < ptol, err
    safe = want['margin'] > margin
    assert float(safe.float().mean()) > (0.5 if dtype == torch.float32 else 0.02)   # (random-init logits: small margins)
    assert torch.equal(pred.cpu()[safe], want['pred'].long()[safe])
    assert isinstance(seg, list) and len(seg) == img.shape[0] and seg[0].dtype == np.int64
    assert np.array_equal(np.stack(seg), pred.cpu().numpy()) and np.array_equal(np.stack(ft), np.stack(seg))


@pytest.mark.gpu
def test_softmax_argmax_bit_exact_given_logits():
    """north_star: argmax maps bit-exact given identical logits (ATen arithmetic order, first index on ties)."""
    g = torch.Generator().manual_seed(9)
    z = torch.randn(2, 21, 33, 47, generator=g) * 3
    z[0, 3] = z[0, 7]                      # exact ties: the first index must win
    for flip, dims in ((None, None), ('horizontal', (3,)), ('vertical', (2,))):
        prob, pred = ops.softmax_argmax(z.to(DEV), flip=flip)
        wp = torch.softmax(z, 1)
        if dims:
            wp = wp.flip(dims=dims)
        # probabilities: the host's vectorised exp differs from expf by an ulp; arg-max: identical
        # wherever the top-2 gap exceeds that, and the FIRST index on the planted exact ties
        assert torch.allclose(prob.cpu(), wp, rtol=2e-6, atol=1e-9), flip
        assert torch.equal(pred, prob.argmax(1)), flip
        top2 = wp.topk(2, dim=1)[0]
        clear = (top2[:, 0] - top2[:, 1]) > 1e-6
        assert torch.equal(pred.cpu()[clear], wp.argmax(1)[clear]), flip
    z2 = torch.zeros(1, 5, 4, 4)
    z2[0, 1] = 2.0
    z2[0, 3] = 2.0                         # classes 1 and 3 tie everywhere: 1 must win
    _, pred = ops.softmax_argmax(z2.to(DEV))
    assert int(pred.min()) == 1 and int(pred.max()) == 1


@pytest.mark.gpu
def test_intersect_and_union_and_miou_vs_reference(G):
    from s4former_b200.core import intersect_and_union, mean_iou, total_intersect_and_union
    want = G['metrics']
    rng = np.random.RandomState(want['seed'])
    pred = rng.randint(0, 5, (3, 40, 50))
    lab = rng.randint(0, 6, (3, 40, 50))
    lab[:, :3] = 255
    got = intersect_and_union(pred[0], lab[0], 5, 255)
    for a, b in zip(got, want['iau0']):
        assert torch.equal(a.cpu(), b)
    got = intersect_and_union(torch.from_numpy(pred[1]).to(DEV), lab[1].copy(), 5, 255, reduce_zero_label=True)
    for a, b in zip(got, want['iau_rzl']):
        assert torch.equal(a.cpu(), b)
    tot = total_intersect_and_union(list(pred), list(lab), 5, 255)
    for a, b in zip(tot, want['total']):
        assert torch.equal(a, b)
    mi = mean_iou(list(pred), list(lab), 5, 255)
    for k, v in want['miou'].items():
        assert np.allclose(mi[k], v, rtol=0, atol=1e-12), k
