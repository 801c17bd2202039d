"""Golden vectors for the inference / validation path and the checkpoint pos_embed resize, produced by
the UNMODIFIED reference files through oracle/ref_harness (runs only where /root/reference exists):

    python -m oracle.make_golden_infer

  * ``VisionTransformer.resize_pos_embed`` (vit.py:447-477) 14x14 -> 32x32 / 8x12, bilinear + bicubic;
  * ``encode_decode`` (:270-297), then the lines of ``whole_inference`` (:1131-1152), ``slide_inference``
    (:1068-1115) and ``inference`` (:1196-1204) + ``simple_test``'s arg-max (:1221) applied as written --
    the shipped ``whole_inference`` / ``slide_inference`` call ``encode_decode`` without its
    ``adaptive_attn_mask`` argument and raise TypeError (SURVEY.md hazard 5), so the harness calls
    ``encode_decode(img, metas, False)`` and restates the ~20 glue lines around it;
  * ``intersect_and_union`` / ``mean_iou`` (core/evaluation/metrics.py:26-165) on seeded label maps.
TEST INFRASTRUCTURE ONLY.
"""
import copy
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

from oracle import golden_common as gc

warnings.filterwarnings('ignore')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def infer_inputs(case):
    g = torch.Generator().manual_seed(100 + case)
    if case == 0:       # whole, padded + rescaled + horizontally flipped, at the training size
        img = torch.randn(2, 3, 128, 128, generator=g)
        meta = dict(img_shape=(120, 128, 3), ori_shape=(90, 96, 3), pad_shape=(128, 128, 3), flip=True,
                    flip_direction='horizontal')
        cfg = dict(mode='whole')
    elif case == 1:     # whole at a different, non-square size: on-the-fly pos_embed resize (vit.py:416-445)
        img = torch.randn(1, 3, 160, 192, generator=g)
        meta = dict(img_shape=(160, 192, 3), ori_shape=(160, 192, 3), pad_shape=(160, 192, 3), flip=False)
        cfg = dict(mode='whole')
    else:               # slide: 128 x 128 windows, stride 96, vertical flip
        img = torch.randn(1, 3, 160, 224, generator=g)
        meta = dict(img_shape=(160, 224, 3), ori_shape=(120, 168, 3), pad_shape=(160, 224, 3), flip=True,
                    flip_direction='vertical')
        cfg = dict(mode='slide', crop_size=(128, 128), stride=(96, 96))
    return img, [dict(meta) for _ in range(img.shape[0])], cfg


def main():
    from oracle.ref_harness import load_reference
    ns = load_reference.load()
    resize = sys.modules['mmseg.ops.wrappers'].resize
    out = {}
    # ---- pos_embed resize
    pe = torch.randn(1, 197, 24, generator=torch.Generator().manual_seed(41))
    out['pos_embed'] = {(mode, hw): ns.VisionTransformer.resize_pos_embed(pe, hw, (14, 14), mode)
                        for mode in ('bilinear', 'bicubic') for hw in ((32, 32), (8, 12))}
    # ---- inference on the tiny S4Former model (eval mode)
    ref = ns.builder.build_segmentor(copy.deepcopy(gc.tiny_cfg('ours')))
    ref.load_state_dict(gc.seeded_state_dict(ref.state_dict(), seed=5))
    ref.eval()
    cases = []
    with torch.no_grad():
        for case in range(3):
            img, metas, cfg = infer_inputs(case)
            if cfg['mode'] == 'whole':
                seg_logit = ref.encode_decode(img, metas, False)
            else:
                h_stride, w_stride = cfg['stride']
                h_crop, w_crop = cfg['crop_size']
                batch_size, _, h_img, w_img = img.size()
                h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
                w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
                preds = img.new_zeros((batch_size, ref.num_classes, h_img, w_img))
                count_mat = img.new_zeros((batch_size, 1, h_img, w_img))
                for h_idx in range(h_grids):
                    for w_idx in range(w_grids):
                        y1 = h_idx * h_stride
                        x1 = w_idx * w_stride
                        y2 = min(y1 + h_crop, h_img)
                        x2 = min(x1 + w_crop, w_img)
                        y1 = max(y2 - h_crop, 0)
                        x1 = max(x2 - w_crop, 0)
                        crop_seg_logit = ref.encode_decode(img[:, :, y1:y2, x1:x2], metas, False)
                        preds += F.pad(crop_seg_logit, (int(x1), int(preds.shape[3] - x2), int(y1),
                                                        int(preds.shape[2] - y2)))
                        count_mat[:, :, y1:y2, x1:x2] += 1
                assert (count_mat == 0).sum() == 0
                seg_logit = preds / count_mat
            resize_shape = metas[0]['img_shape'][:2]
            seg_logit = seg_logit[:, :, :resize_shape[0], :resize_shape[1]]
            seg_logit = resize(seg_logit, size=metas[0]['ori_shape'][:2], mode='bilinear',
                               align_corners=ref.align_corners, warning=False)
            output = F.softmax(seg_logit, dim=1)
            if metas[0]['flip']:
                output = output.flip(dims=(3,)) if metas[0]['flip_direction'] == 'horizontal' else output.flip(dims=(2,))
            pred = output.argmax(dim=1)
            top2 = output.topk(2, dim=1)[0]
            cases.append(dict(prob=output.clone(), pred=pred.to(torch.uint8), margin=(top2[:, 0] - top2[:, 1]).clone(),
                              img_checksum=float(img.double().abs().sum())))
    out['infer'] = cases
    # ---- metrics
    spec = importlib.util.spec_from_file_location('ref_metrics', os.path.join(load_reference.REF,
                                                                            'mmseg/core/evaluation/metrics.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.RandomState(7)
    pred = rng.randint(0, 5, (3, 40, 50))
    lab = rng.randint(0, 6, (3, 40, 50))
    lab[:, :3] = 255
    out['metrics'] = dict(seed=7, iau0=m.intersect_and_union(pred[0], lab[0], 5, 255),
                          total=m.total_intersect_and_union(list(pred), list(lab), 5, 255),
                          miou=m.mean_iou(list(pred), list(lab), 5, 255),
                          iau_rzl=m.intersect_and_union(pred[1], lab[1].copy(), 5, 255, reduce_zero_label=True))
    path = os.path.join(OUT, 'infer.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes;',
          {i: (tuple(c['pred'].shape), float(c['margin'].median())) for i, c in enumerate(cases)})


if __name__ == '__main__':
    sys.exit(main())
