"""Host-side glue mirroring ``mmseg/models/utils/structual_utils.py:42-53,132-154`` and
``mmseg/core/utils/misc.py:2-18``."""
from collections import Counter
from collections.abc import Mapping
from numbers import Number

import torch

_step_counter = Counter()


def dict_fuse(obj_list, reference_obj):
    if isinstance(reference_obj, torch.Tensor):
        return torch.stack(obj_list)
    return obj_list


def dict_select(dict1, key, value):
    flag = [v == value for v in dict1[key]]
    return {k: dict_fuse([vv for vv, ff in zip(v, flag) if ff], v) for k, v in dict1.items()}


def dict_split(dict1, key):
    """Group a flattened batch by ``dict1[key]`` (the per-image tag)."""
    group_names = list(dict.fromkeys(dict1[key]))
    return {k: dict_select(dict1, key, k) for k in group_names}


def sequence_mul(obj, multiplier):
    if isinstance(obj, (list, tuple)):
        return [o * multiplier for o in obj]
    return obj * multiplier


def weighted_loss(loss, weight, ignore_keys=(), warmup=0):
    """Scale every entry whose key contains "loss" (structual_utils.py:132-154)."""
    _step_counter['weight'] += 1

    def lambda_weight(x):
        if _step_counter['weight'] <= warmup:
            return x * (_step_counter['weight'] - 1) / warmup
        return x
    if isinstance(weight, Mapping):
        for k, v in weight.items():
            for name in loss:
                if (k in name) and ('loss' in name):
                    loss[name] = sequence_mul(loss[name], lambda_weight(v))
    elif isinstance(weight, Number):
        for name in loss:
            if 'loss' in name:
                if not any(kw in name for kw in ignore_keys):
                    loss[name] = sequence_mul(loss[name], lambda_weight(weight))
                else:
                    loss[name] = sequence_mul(loss[name], 0.0)
    else:
        raise NotImplementedError()
    return loss


def add_prefix(inputs, prefix):
    return {f'{prefix}.{name}': value for name, value in inputs.items()}
