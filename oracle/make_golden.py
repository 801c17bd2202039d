"""Generate tests/golden/*.pt by running the UNMODIFIED reference (through
oracle/ref_harness) on seeded inputs.  Runs only in the build container, where
/root/reference exists:

    python -m oracle.make_golden

Each golden file stores the reference's outputs plus checksums of the seeded inputs /
weights (regenerated in the tests by ``golden_common``), so a replay that generates
different inputs fails loudly instead of silently comparing different things.
TEST INFRASTRUCTURE ONLY.
"""
import copy
import os
import sys
import warnings

import numpy as np
import torch

from oracle import golden_common as gc
from oracle import s4former_oracle as O
from oracle.ref_harness import load_reference

warnings.filterwarnings('ignore')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def build_reference(ns, cfg):
    m = ns.builder.build_segmentor(copy.deepcopy(cfg))
    m.train()
    return m


def ref_forward_train(ns, model, img, gt, metas, it=0):
    metas = copy.deepcopy(metas)
    losses = model.forward_train(img, metas, gt_semantic_seg=gt, iter=it)
    return losses, metas


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = load_reference.load()
    torch.set_num_threads(8)

    # ---------------------------------------------------------------- losses / pseudo label
    g = torch.Generator().manual_seed(11)
    z_t = torch.randn(2, 5, 32, 32, generator=g) * 4.0
    z_s = torch.randn(2, 5, 32, 32, generator=g) * 2.0
    ed = build_reference(ns, gc.tiny_cfg('ours'))
    p = torch.softmax(z_t, 1)
    max_value, hard = torch.max(p, 1)
    conf = (max_value > 0.95) * 1
    hard_i = hard.clone()
    hard_i[conf == 0] = 255
    # the reference's own compute_pseudo_loss with a stub head returning z_s
    teacher = dict(seg_logits=z_t, hard_seg_label=hard_i, conf_mask=conf)

    class _Stub(torch.nn.Module):
        def forward_get_logits(self, *a, **k):
            return z_s
    real_head = ed.decode_head
    ed.decode_head = _Stub()
    out = ed.compute_pseudo_loss(dict(backbone_feature=None, img_metas=[{}], img=z_s), teacher)
    ed.decode_head = real_head
    # patch unconfidence exactly as encoder_decoder.py:547-555 (patch 16 -> use 8 for 32px)
    ps = 8
    cm = conf.view(conf.size(0), conf.size(1) // ps, ps, conf.size(1) // ps, ps)
    cm = (1 - cm).permute(0, 1, 3, 2, 4).reshape(conf.size(0), conf.size(1) // ps, conf.size(1) // ps, -1)
    u = torch.sum(cm, -1) / (ps * ps)
    ce_ref = ns.CrossEntropyLoss(use_sigmoid=False, loss_weight=0.4)(z_s, hard_i, ignore_index=255)
    torch.save(dict(z_t=z_t, z_s=z_s, hard=hard_i, conf=conf, u=u, patch=ps,
                    loss_seg_unsup=out['loss_seg_unsup'], loss_ncr_unsup=out['loss_ncr_unsup'],
                    mask_ratio=out['mask_ratio'], ce_w04=ce_ref),
               os.path.join(OUT, 'loss_pseudo.pt'))
    print('loss golden:', float(out['loss_seg_unsup']), float(out['loss_ncr_unsup']),
          'oracle:', float(O.cross_entropy_mean_all(z_s, hard_i)), float(O.ncr_unsup_only(z_s, z_t, hard_i)))

    # ---------------------------------------------------------------- augmentation
    img = torch.randn(4, 3, 64, 64, generator=g)
    lab = torch.randint(0, 5, (4, 64, 64), generator=g)
    O.seed_host_rng(7)
    tinfo, sinfo = ns.gen.generate_unsup_cutmix_data(dict(hard_seg_label=lab.clone()),
                                                     dict(img=img.clone()), ratio=2, patchwise=False)
    metas = [dict() for _ in range(4)]
    sinfo2 = dict(img=sinfo['img'].clone(), img_metas=metas)
    sinfo2, _ = ns.gen.generate_unsup_patchmix_data(sinfo2, tinfo, PatchMix_N=1, patchmix_ratio=0.5)
    perms = torch.stack([torch.as_tensor(m['PatchMixIndex']) for m in metas])
    # feature un-shuffle through the reference head method
    head = ed.decode_head
    tok = torch.randn(4, 64, 16, generator=g)
    unsh = head._repatchmix_inputs(tok, 2, perms)
    torch.save(dict(img=img, lab=lab, seed=7, cut_img=sinfo['img'], cut_lab=tinfo['hard_seg_label'],
                    shuf_img=sinfo2['img'], perms=perms, tok=tok, unsh=unsh),
               os.path.join(OUT, 'augment.pt'))

    # ---------------------------------------------------------------- backbone / head / full step
    for variant in ('ours', 'mt', 'sup'):
        cfg = gc.tiny_cfg(variant)
        ref = build_reference(ns, cfg)
        sd = gc.seeded_state_dict(ref.state_dict(), seed=5)
        ref.load_state_dict(sd)
        extra = {}
        if variant == 'ours':
            # stand-alone backbone with PASA and head outputs for unit parity
            ref.eval()
            with torch.no_grad():
                g2 = torch.Generator().manual_seed(23)
                u = torch.rand(2, 8, 8, generator=g2).mul(16).round().div(16)
                x = torch.randn(2, 3, 128, 128, generator=g2)
                feats = ref.backbone(x, attn_mask=u, attn_mask_weight=5, adaptive_attn_mask=True)
                feats_plain = ref.backbone(x)
                logits = ref.decode_head.forward(feats_plain)
                idx = torch.topk(u.reshape(2, -1), 32, dim=-1, largest=False)[1]
            ref.train()
            extra.update(vit_seed=23, vit_x_checksum=float(x.double().abs().sum()), vit_u=u, vit_topk=idx, vit_feats=[f.clone() for f in feats],
                       vit_feats_plain=[f.clone() for f in feats_plain], head_logits_eval=logits[:, :, ::2, ::2].clone())
        img, gt, metas = gc.tiny_batch(variant)
        O.seed_host_rng(1999)
        ref.zero_grad()
        losses, metas_after = ref_forward_train(ns, ref, img, gt, metas)
        total = sum(v for k, v in losses.items() if 'loss' in k)
        total.backward()
        grads = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
        gsel = {k: v for k, v in grads.items() if k in gc.GRAD_KEYS}
        gnorm = {k: float(v.norm()) for k, v in grads.items()}
        post = ref.state_dict()
        rec = dict(variant=variant, losses={k: v.detach() for k, v in losses.items()},
                   grads=gsel, grad_norms=gnorm,
                   sd_checksum=gc.checksum(sd), img_checksum=float(img.double().abs().sum()),
                   ema_after={k: post[k].clone() for k in gc.EMA_KEYS if k in post},
                   bn_after={k: post[k].clone() for k in post if 'running_' in k and 'ema' not in k
                             and k.startswith('decode_head')},
                   perms=[m.get('PatchMixIndex') for m in metas_after if m['tag'] == 'unsup_student'])
        rec.update(extra)
        torch.save(rec, os.path.join(OUT, f'step_{variant}.pt'))

        # immediate check of the oracle against the reference
        orc = O.OracleEncoderDecoder(**{k: v for k, v in cfg.items() if k != 'type'})
        orc.load_state_dict(sd)
        orc.train()
        O.seed_host_rng(1999)
        lo = orc.forward_train(img, copy.deepcopy(metas), gt)
        tot_o = O.parse_losses(lo)
        tot_o.backward()
        print(variant, 'loss ref', float(total), 'oracle', float(tot_o))
        for k in losses:
            print('   ', k, float(losses[k]), float(lo[k]))
        worst = 0.0
        for k, p in orc.named_parameters():
            if p.grad is not None and k in grads:
                d = float((p.grad - grads[k]).norm() / (grads[k].norm() + 1e-12))
                worst = max(worst, d)
        print('    worst grad rel diff', worst)
        # what PyTorch's own bf16 autocast of the same step loses against fp32: the yardstick for
        # the bf16 gradient gate on this tiny, noisy problem (tests/test_step_gpu.py)
        orb = O.OracleEncoderDecoder(**{k: v for k, v in cfg.items() if k != 'type'})
        orb.load_state_dict(sd)
        orb.train()
        O.seed_host_rng(1999)
        with torch.autocast('cpu', dtype=torch.bfloat16):
            lb = orb.forward_train(img, copy.deepcopy(metas), gt)
        O.parse_losses(lb).backward()
        nb = dict(orb.named_parameters())
        rec['bf16_autocast_err'] = {k: float((nb[k].grad - g).norm() / (g.norm() + 1e-12))
                                    for k, g in gsel.items()}
        torch.save(rec, os.path.join(OUT, f'step_{variant}.pt'))
    sz = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print('golden bytes', sz)


if __name__ == '__main__':
    sys.exit(main())
