"""Generate tests/golden/segformer_{ours,sup}.pt by running the UNMODIFIED reference
(``mit.py``, ``segformer_head.py``, ``encoder_decoder.py`` through oracle/ref_harness) on seeded
inputs, and check ``oracle/segformer_oracle.py`` against it on the spot.  Build container only
(/root/reference):

    python -m oracle.make_golden_segformer

SURVEY.md section 8(f) rank 2 / BASELINE config 5: this pins the ORACLE of the SegFormer / MiT
variant; there is no CUDA path for it yet.  TEST INFRASTRUCTURE ONLY.
"""
import copy
import os
import sys
import warnings

import torch

from oracle import golden_common as gc
from oracle import s4former_oracle as O
from oracle.ref_harness import load_reference
from oracle.check_segformer_path import load_segformer

warnings.filterwarnings('ignore')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
GRAD_KEYS = ('backbone.layers.0.0.projection.weight', 'backbone.layers.0.1.0.attn.sr.weight',
             'backbone.layers.0.1.0.ffn.layers.1.weight', 'backbone.layers.1.1.1.attn.attn.in_proj_weight',
             'backbone.layers.3.1.0.attn.attn.in_proj_weight', 'backbone.layers.3.1.1.ffn.layers.4.weight',
             'backbone.layers.3.2.weight', 'decode_head.convs.3.conv.weight', 'decode_head.convs.0.bn.weight',
             'decode_head.fusion_conv.conv.weight', 'decode_head.conv_seg.weight', 'decode_head.conv_seg.bias')


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = load_reference.load()
    load_segformer(ns)
    torch.set_num_threads(8)
    ok = True
    for variant in ('ours', 'sup'):
        cfg = gc.tiny_segformer_cfg(variant)
        ref = ns.builder.build_segmentor(copy.deepcopy(cfg))
        ref.train()
        sd = gc.seeded_state_dict(ref.state_dict(), seed=5, ema_cls_std=20.0)
        ref.load_state_dict(sd)
        orc = O.OracleEncoderDecoder(**{k: v for k, v in cfg.items() if k != 'type'})
        missing = set(sd) ^ set(orc.state_dict())
        assert not missing, ('state_dict keys differ between the reference and the oracle', sorted(missing)[:8])
        orc.load_state_dict(sd)
        orc.train()
        extra = {}
        if variant == 'ours':
            # stand-alone backbone with the patch-adaptive mask, and the head with a PatchMix un-shuffle
            ref.eval()
            orc.eval()
            with torch.no_grad():
                g2 = torch.Generator().manual_seed(23)
                u = torch.rand(2, 4, 4, generator=g2).mul(64).round().div(64)     # 128 / 32 = 4 x 4 patches
                x = torch.randn(2, 3, 128, 128, generator=g2)
                feats = ref.backbone(x, attn_mask=u, attn_mask_weight=5, adaptive_attn_mask=True)
                feats_plain = ref.backbone(x)
                perms = torch.stack([torch.randperm(4, generator=g2) for _ in range(2)])    # (128 / 64)^2 blocks
                logits = ref.decode_head.forward(feats_plain, PatchMix_N=4, PatchMixIndex=perms)
                idx = torch.topk(u.reshape(2, -1)[:, 1:], int(0.5 * 15), dim=-1, largest=False)[1]
                fo = orc.backbone(x, attn_mask=u, attn_mask_weight=5, adaptive_attn_mask=True)
                fp = orc.backbone(x)
                lo = orc.decode_head.forward(fp, PatchMix_N=4, PatchMixIndex=perms)
            d1 = max(float((a - b).abs().max()) for a, b in zip(feats, fo))
            d2 = max(float((a - b).abs().max()) for a, b in zip(feats_plain, fp))
            d3 = float((logits - lo).abs().max())
            print(f'backbone (masked) max abs diff {d1:.2e}, plain {d2:.2e}, head+unshuffle logits {d3:.2e}',
                  ' mask changes the features by', float((feats[3] - feats_plain[3]).abs().max()))
            ok &= d1 < 1e-4 and d2 < 1e-4 and d3 < 1e-4
            ref.train()
            orc.train()
            extra.update(mit_seed=23, mit_x_checksum=float(x.double().abs().sum()), mit_u=u, mit_topk=idx, mit_perms=perms,
                         mit_feats=[f.clone() for f in feats], mit_feats_plain=[f.clone() for f in feats_plain],
                         head_logits_unshuffled=logits.clone())
        img, gt, metas = gc.tiny_batch(variant)
        O.seed_host_rng(1999)
        ref.zero_grad()
        metas_ref = copy.deepcopy(metas)
        losses = ref.forward_train(img, metas_ref, gt_semantic_seg=gt, iter=0)
        total = sum(v for k, v in losses.items() if 'loss' in k)
        total.backward()
        grads = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
        post = ref.state_dict()
        rec = dict(variant=variant, losses={k: v.detach() for k, v in losses.items()},
                   grads={k: v for k, v in grads.items() if k in GRAD_KEYS},
                   grad_norms={k: float(v.norm()) for k, v in grads.items()},
                   sd_checksum=gc.checksum(sd), img_checksum=float(img.double().abs().sum()),
                   bn_after={k: post[k].clone() for k in post if 'running_' in k and 'ema' not in k},
                   perms=[m.get('PatchMixIndex') for m in metas_ref if m['tag'] == 'unsup_student'])
        rec.update(extra)
        torch.save(rec, os.path.join(OUT, f'segformer_{variant}.pt'))

        O.seed_host_rng(1999)
        lo = orc.forward_train(img, copy.deepcopy(metas), gt)
        tot_o = O.parse_losses(lo)
        tot_o.backward()
        print(variant, 'loss ref', float(total), 'oracle', float(tot_o))
        for k in losses:
            if k not in lo:
                print('    (reference-only key)', k, float(losses[k]))
                continue
            print('   ', k, float(losses[k]), float(lo[k]))
            if 'loss' in k:
                ok &= abs(float(losses[k]) - float(lo[k])) < 1e-5 * max(1.0, abs(float(losses[k])))
        # (a per-channel shift in front of conv -> BatchNorm(train) has an analytically ZERO gradient: the
        # stage norms' biases and the head convs' ... hold pure rounding noise, ~1e-9 of the others; they
        # are compared against the largest gradient norm instead of their own)
        gmax = max(float(v.norm()) for v in grads.values())
        worst, wk, noise = 0.0, None, []
        for k, p in orc.named_parameters():
            if p.grad is not None and k in grads:
                gn = float(grads[k].norm())
                if gn < 1e-6 * gmax:
                    noise.append(k)
                    ok &= float(p.grad.norm()) < 1e-6 * gmax
                    continue
                d = float((p.grad - grads[k]).norm() / (gn + 1e-12))
                if d > worst:
                    worst, wk = d, k
        print('    worst grad rel diff', worst, wk, ' tensors with a gradient:', len(grads),
              ' analytically-zero gradients:', noise)
        rec['zero_grad_keys'] = noise
        torch.save(rec, os.path.join(OUT, f'segformer_{variant}.pt'))
        ok &= worst < 1e-4
    print('oracle == reference:', ok)
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
