"""mIoU evaluation -- mirror of ``mmseg/core/evaluation/metrics.py`` (``intersect_and_union`` :26-91,
``total_intersect_and_union`` :94-131, ``mean_iou`` :134-165, ``eval_metrics`` :244-289,
``total_area_to_metrics`` :322-390) for in-memory predictions.

The class histograms are accumulated ON THE DEVICE by ``s4_intersect_union`` (predictions come out
of ``simple_test`` there; no per-image device->host copy of the label maps); counts are exact
integers, so the areas are bit-identical to the reference's ``torch.histc`` results."""
from collections import OrderedDict

import numpy as np
import torch

from .. import ops


def _to_device_long(x, device):
    if isinstance(x, str):
        raise NotImplementedError('file-name inputs (np.load / mmcv.imread) are dataset plumbing, out of scope')
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=torch.int64)


def intersect_and_union(pred_label, label, num_classes, ignore_index, label_map=dict(), reduce_zero_label=False,
                        device=None):
    """-> (area_intersect, area_union, area_pred_label, area_label), float32 tensors of [num_classes]
    like the reference's ``torch.histc`` outputs."""
    if device is None:
        device = pred_label.device if torch.is_tensor(pred_label) and pred_label.is_cuda else torch.device('cuda')
    pred = _to_device_long(pred_label, device)
    lab = _to_device_long(label, device)
    if label_map:
        lab = lab.clone()
        for old_id, new_id in label_map.items():
            lab[lab == old_id] = new_id
    if reduce_zero_label:
        lab = lab.clone()
        lab[lab == 0] = 255
        lab = lab - 1
        lab[lab == 254] = 255
    hist = ops.intersect_union_hist(pred.reshape(-1), lab.reshape(-1), num_classes, ignore_index)
    inter, pred_a, lab_a = hist[0].float(), hist[1].float(), hist[2].float()
    return inter, pred_a + lab_a - inter, pred_a, lab_a


def total_intersect_and_union(results, gt_seg_maps, num_classes, ignore_index, label_map=dict(),
                              reduce_zero_label=False, device=None):
    tot = None
    for result, gt in zip(results, gt_seg_maps):
        areas = intersect_and_union(result, gt, num_classes, ignore_index, label_map, reduce_zero_label, device)
        areas = [a.double() for a in areas]
        tot = areas if tot is None else [t + a for t, a in zip(tot, areas)]
    if tot is None:
        z = torch.zeros((num_classes,), dtype=torch.float64)
        return z, z.clone(), z.clone(), z.clone()
    return tuple(t.cpu() for t in tot)


def total_area_to_metrics(total_area_intersect, total_area_union, total_area_pred_label, total_area_label,
                          metrics=['mIoU'], nan_to_num=None, beta=1):
    if isinstance(metrics, str):
        metrics = [metrics]
    allowed = ['mIoU', 'mDice']
    if not set(metrics).issubset(set(allowed)):
        raise KeyError(f'metrics {metrics} is not supported')
    all_acc = total_area_intersect.sum() / total_area_label.sum()
    ret = OrderedDict({'aAcc': all_acc})
    for metric in metrics:
        if metric == 'mIoU':
            ret['IoU'] = total_area_intersect / total_area_union
            ret['Acc'] = total_area_intersect / total_area_label
        elif metric == 'mDice':
            ret['Dice'] = 2 * total_area_intersect / (total_area_pred_label + total_area_label)
            ret['Acc'] = total_area_intersect / total_area_label
    ret = {k: v.numpy() for k, v in ret.items()}
    if nan_to_num is not None:
        ret = OrderedDict({k: np.nan_to_num(v, nan=nan_to_num) for k, v in ret.items()})
    return ret


def eval_metrics(results, gt_seg_maps, num_classes, ignore_index, metrics=['mIoU'], nan_to_num=None,
                 label_map=dict(), reduce_zero_label=False, beta=1):
    areas = total_intersect_and_union(results, gt_seg_maps, num_classes, ignore_index, label_map, reduce_zero_label)
    return total_area_to_metrics(*areas, metrics, nan_to_num, beta)


def mean_iou(results, gt_seg_maps, num_classes, ignore_index, nan_to_num=None, label_map=dict(),
             reduce_zero_label=False):
    return eval_metrics(results, gt_seg_maps, num_classes, ignore_index, ['mIoU'], nan_to_num, label_map,
                        reduce_zero_label)
