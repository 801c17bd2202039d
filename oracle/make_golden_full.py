"""Full-size golden fixtures: the UNMODIFIED reference (through oracle/ref_harness) runs one
S4Former-full train step at the BASELINE shapes

    full512 : DeiT-B SETR-PUP, 512x512, 21 classes, 2 labeled + 2 unlabeled crops
    full768 : DeiT-B SETR-PUP, 768x768, 19 classes, 1 labeled + 1 unlabeled crop   (--with-768)

on seeded inputs / weights and stores, in a few hundred KB per shape (not GBs):

  * the 8 loss values (``encoder_decoder.py:386-514, 516-687``),
  * the teacher outputs as checksums + strided samples: logits, ``hard_seg_label`` (after
    ``:541-542``), ``conf_mask``, the patch unconfidence ``u`` (``:547-555``) and the PASA
    top-k index set the reference's CPU ``torch.topk`` picked (``vit.py:526``),
  * for every parameter: the gradient norm; for ~30 tensors a strided sample (<= 4096 elements)
    of the gradient itself,
  * the same gradient samples from the ORACLE under ``torch.autocast(bfloat16)`` and their error
    against fp32: the yardstick for "what bf16 arithmetic of the reference math loses" at this
    shape (tests/test_full_parity_gpu.py gates the tcgen05 path at 2e-2 wherever that yardstick
    itself is below 2e-2),
  * checksums of the seeded inputs and weights (the tests regenerate both from seeds).

Runs only in the build container (needs /root/reference):

    python -m oracle.make_golden_full [--with-768]

TEST INFRASTRUCTURE ONLY.
"""
import argparse
import copy
import os
import sys
import time
import warnings

import torch

from oracle import golden_common as gc
from oracle import s4former_oracle as O

warnings.filterwarnings('ignore')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def run_reference(ns, shape):
    cfg = gc.full_cfg(shape)
    ref = ns.builder.build_segmentor(copy.deepcopy(cfg))
    ref.train()
    sd = gc.seeded_state_dict(ref.state_dict(), seed=gc.FULL[shape]['wseed'], ema_cls_std=gc.FULL[shape]['ema_cls_std'])
    ref.load_state_dict(sd)
    img, gt, metas = gc.full_batch(shape)
    rec = {}
    # capture the teacher outputs the reference's own extract_teacher_info_ema produces
    orig = ref.extract_teacher_info_ema

    def spy(*a, **k):
        out = orig(*a, **k)
        rec['teacher_logits'] = out['seg_logits'].detach().clone()
        rec['conf'] = out['conf_mask'].detach().clone()
        rec['hard_before'] = out['hard_seg_label'].detach().clone()
        return out
    ref.extract_teacher_info_ema = spy
    # ... and the attention mask / top-k the backbone sees in the PASA pass
    bb_fwd = ref.backbone.forward

    def bb_spy(*a, **k):
        if k.get('attn_mask') is not None and 'u' not in rec:
            rec['u'] = k['attn_mask'].detach().clone()
        return bb_fwd(*a, **k)
    ref.backbone.forward = bb_spy
    O.seed_host_rng(1999)
    ref.zero_grad()
    t0 = time.time()
    metas_run = copy.deepcopy(metas)
    losses = ref.forward_train(img, metas_run, gt_semantic_seg=gt, iter=0)
    total = sum(v for k, v in losses.items() if 'loss' in k)
    total.backward()
    print(f'[{shape}] reference step {time.time() - t0:.1f} s; loss {float(total):.6f}')
    grads = {k: p.grad.detach() for k, p in ref.named_parameters() if p.grad is not None}
    post = ref.state_dict()
    return cfg, sd, img, gt, metas, metas_run, losses, grads, rec, post


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--with-768', action='store_true')
    ap.add_argument('--only-768', action='store_true')
    a = ap.parse_args()
    from oracle.ref_harness import load_reference
    ns = load_reference.load()
    torch.set_num_threads(os.cpu_count() or 8)
    os.makedirs(OUT, exist_ok=True)
    shapes = (['full512'] if not a.only_768 else []) + (['full768'] if (a.with_768 or a.only_768) else [])
    for shape in shapes:
        cfg, sd, img, gt, metas, metas_run, losses, grads, rec, post = run_reference(ns, shape)
        conf, zt = rec['conf'], rec['teacher_logits']
        hard = rec['hard_before'].clone()
        hard[conf == 0] = 255          # encoder_decoder.py:541-542
        u = rec['u']
        flat = u.reshape(u.shape[0], -1)
        topk = torch.topk(flat, int(0.5 * flat.size(-1)), dim=-1, largest=False)[1]   # vit.py:526
        print(f'[{shape}] mask ratio {float(conf.float().mean()):.4f}; losses:',
              {k: round(float(v), 6) for k, v in losses.items()})
        out = dict(
            shape=shape, spec=gc.FULL[shape],
            losses={k: float(v) for k, v in losses.items()},
            sd_checksum=gc.checksum(sd), img_checksum=float(img.double().abs().sum()),
            gt_checksum=int(gt.sum()),
            teacher=dict(
                logits_sample=gc.strided_sample(zt, 8192), logits_norm=float(zt.double().norm()),
                logits_abs_max=float(zt.abs().max()),
                conf_sum=int(conf.sum()), hard_sum=int(hard.sum()),
                hard_hist=torch.bincount(hard.reshape(-1), minlength=256),
                conf_rowsum=conf.sum(-1).to(torch.int32),       # per image row: localises a mismatch
                u=u.clone(), topk=topk.to(torch.int32)),
            grad_norms={k: float(v.double().norm()) for k, v in grads.items()},
            grad_samples={k: gc.strided_sample(grads[k], 4096) for k in gc.full_grad_keys(grads)},
            perms=[m.get('PatchMixIndex') for m in metas_run if m['tag'] == 'unsup_student'],
            ema_after={k: gc.strided_sample(post[k], 1024) for k in gc.EMA_KEYS if k in post},
            bn_after={k: post[k].clone() for k in post if 'running_' in k and 'ema' not in k
                      and k.startswith('decode_head')})
        # ---- oracle (restatement) against the reference at this shape, and the bf16 yardstick ----
        ocfg = {k: v for k, v in cfg.items() if k != 'type'}
        worst = {}
        for mode in ('fp32', 'bf16_autocast'):
            orc = O.OracleEncoderDecoder(**ocfg)
            orc.load_state_dict(sd)
            orc.train()
            O.seed_host_rng(1999)
            t0 = time.time()
            if mode == 'fp32':
                lo = orc.forward_train(img, copy.deepcopy(metas), gt)
            else:
                with torch.autocast('cpu', dtype=torch.bfloat16):
                    lo = orc.forward_train(img, copy.deepcopy(metas), gt, topk_idx=topk.long())
            O.parse_losses(lo).backward()
            named = dict(orc.named_parameters())
            err = {}
            for k, g in grads.items():
                if named[k].grad is None:
                    continue
                err[k] = float((named[k].grad.double() - g.double()).norm() / (g.double().norm() + 1e-30))
            worst[mode] = max(err.values())
            print(f'[{shape}] oracle {mode}: {time.time() - t0:.1f} s, loss {float(O.parse_losses(lo)):.6f}, '
                  f'worst grad rel err {worst[mode]:.3e}')
            out[f'oracle_{mode}_grad_err'] = err
            out[f'oracle_{mode}_losses'] = {k: float(v) for k, v in lo.items()}
        assert worst['fp32'] < 1e-3, 'the oracle restatement disagrees with the reference at full size'
        path = os.path.join(OUT, f'step_{shape}.pt')
        torch.save(out, path)
        print(f'[{shape}] wrote {path}: {os.path.getsize(path)} bytes')


if __name__ == '__main__':
    sys.exit(main())
