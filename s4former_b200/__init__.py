"""s4former_b200 -- B200-native (sm_100a) implementation of the S4Former semi-supervised
train step behind mmseg's module API.

Public surface (same registry type names as the reference, ``mmseg/models/builder.py:8-15``):
``EncoderDecoder``, ``VisionTransformer``, ``SETRUPHead``, ``CrossEntropyLoss``, built with
``build_segmentor(cfg)`` from the unchanged ``configs/setr/*.py`` model dicts.

All arithmetic runs in ``libs4former_b200.so`` (``include/s4former.h``); importing this package
does not require a GPU, running any op does, and there is no CPU fallback.
"""
from . import _lib, ops
from .builder import (BACKBONES, HEADS, LOSSES, MODELS, SEGMENTORS, build_backbone, build_head,
                      build_loss, build_segmentor)
from .backbones.vit import VisionTransformer
from .decode_heads.setr_up_head import SETRUPHead
from .losses.cross_entropy_loss import CrossEntropyLoss
from .segmentors.encoder_decoder import EncoderDecoder
from .ops import set_compute_dtype, set_backend
from .parallel import S4DistributedDataParallel, register_ddp_into_mmseg

__all__ = ['EncoderDecoder', 'VisionTransformer', 'SETRUPHead', 'CrossEntropyLoss', 'MODELS',
           'BACKBONES', 'HEADS', 'LOSSES', 'SEGMENTORS', 'build_backbone', 'build_head', 'build_loss',
           'build_segmentor', 'set_compute_dtype', 'set_backend', 'register_into_mmseg', 'ops',
           'S4DistributedDataParallel', 'register_ddp_into_mmseg']


def register_into_mmseg():
    """Register the four classes in mmseg's own registry (needs mmcv + mmseg importable), so
    ``tools/train.py CONFIG`` builds the B200 implementation from the unchanged configs."""
    from mmseg.models.builder import MODELS as MMSEG_MODELS
    for cls in (EncoderDecoder, VisionTransformer, SETRUPHead, CrossEntropyLoss):
        MMSEG_MODELS.register_module(name=cls.__name__, force=True, module=cls)
