"""Import the reference's hot-path modules UNMODIFIED from /root/reference.

TEST INFRASTRUCTURE ONLY (runs in the build container; /root/reference does not exist
on the GPU box).  Used by ``oracle/make_golden.py`` to pin ``oracle/s4former_oracle.py``
against the reference's own lines and to generate ``tests/golden/*.pt``.

A synthetic ``mmseg`` package skeleton is placed in ``sys.modules`` so that only the
files on the path are executed:

  mmseg.ops.wrappers                       (real file)
  mmseg.utils.generate_unsup_data          (real file; re-exported from mmseg.utils, which
                                            also works around hazard 9: mmseg/utils/__init__.py
                                            never exports generate_unsup_patchmix_data)
  mmseg.core.utils.misc.add_prefix         (real file)
  mmseg.models.builder                     (real file, on the shim Registry)
  mmseg.models.utils.{embed,structual_utils}            (real files)
  mmseg.models.losses.{utils,cross_entropy_loss}        (real files)
  mmseg.models.backbones.vit               (real file)
  mmseg.models.decode_heads.{decode_head,setr_up_head}  (real files)
  mmseg.models.segmentors.{base,encoder_decoder}        (real files)
"""
import importlib.util
import logging
import os
import sys
import types

REF = os.environ.get('S4_REFERENCE_ROOT', '/root/reference')


def _pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load():
    """Returns a namespace with the reference classes/functions on the hot path."""
    if 'mmseg.models.segmentors.encoder_decoder' in sys.modules:
        return _namespace()
    if not os.path.isdir(REF):
        raise RuntimeError(f'reference tree not found at {REF}')
    from . import mmcv_shim
    mmcv_shim.install()

    mmseg = _pkg('mmseg')
    ops = _pkg('mmseg.ops')
    wr = _load('mmseg.ops.wrappers', 'mmseg/ops/wrappers.py')
    ops.resize, ops.Upsample = wr.resize, wr.Upsample

    utils = _pkg('mmseg.utils')
    utils.get_root_logger = lambda *a, **k: logging.getLogger('mmseg')
    gen = _load('mmseg.utils.generate_unsup_data', 'mmseg/utils/generate_unsup_data.py')
    for k, v in vars(gen).items():
        if k.startswith('generate_') or k.startswith('cut_mix'):
            setattr(utils, k, v)

    core = _pkg('mmseg.core')
    _pkg('mmseg.core.utils')
    misc = _load('mmseg.core.utils.misc', 'mmseg/core/utils/misc.py')
    core.add_prefix = misc.add_prefix

    def build_pixel_sampler(cfg, **kw):
        raise NotImplementedError('sampler=None on the hot path')
    core.build_pixel_sampler = build_pixel_sampler

    models = _pkg('mmseg.models')
    builder = _load('mmseg.models.builder', 'mmseg/models/builder.py')
    models.builder = builder

    mutils = _pkg('mmseg.models.utils')
    embed = _load('mmseg.models.utils.embed', 'mmseg/models/utils/embed.py')
    su = _load('mmseg.models.utils.structual_utils', 'mmseg/models/utils/structual_utils.py')
    mutils.PatchEmbed = embed.PatchEmbed
    mutils.structual_utils = su

    losses = _pkg('mmseg.models.losses')
    lutils = _load('mmseg.models.losses.utils', 'mmseg/models/losses/utils.py')
    ce = _load('mmseg.models.losses.cross_entropy_loss',
               'mmseg/models/losses/cross_entropy_loss.py')
    losses.accuracy = lambda *a, **k: None   # call site is commented out (decode_head.py:353)
    losses.CrossEntropyLoss = ce.CrossEntropyLoss

    _pkg('mmseg.models.backbones')
    _load('mmseg.models.backbones.vit', 'mmseg/models/backbones/vit.py')
    _pkg('mmseg.models.decode_heads')
    _load('mmseg.models.decode_heads.decode_head', 'mmseg/models/decode_heads/decode_head.py')
    _load('mmseg.models.decode_heads.setr_up_head', 'mmseg/models/decode_heads/setr_up_head.py')
    _pkg('mmseg.models.segmentors')
    _load('mmseg.models.segmentors.base', 'mmseg/models/segmentors/base.py')
    _load('mmseg.models.segmentors.encoder_decoder',
          'mmseg/models/segmentors/encoder_decoder.py')
    return _namespace()


def _namespace():
    ns = types.SimpleNamespace()
    ns.vit = sys.modules['mmseg.models.backbones.vit']
    ns.VisionTransformer = ns.vit.VisionTransformer
    ns.SETRUPHead = sys.modules['mmseg.models.decode_heads.setr_up_head'].SETRUPHead
    ns.CrossEntropyLoss = sys.modules['mmseg.models.losses.cross_entropy_loss'].CrossEntropyLoss
    ns.EncoderDecoder = sys.modules['mmseg.models.segmentors.encoder_decoder'].EncoderDecoder
    ns.BaseSegmentor = sys.modules['mmseg.models.segmentors.base'].BaseSegmentor
    ns.gen = sys.modules['mmseg.utils.generate_unsup_data']
    ns.structual_utils = sys.modules['mmseg.models.utils.structual_utils']
    ns.builder = sys.modules['mmseg.models.builder']
    ns.PatchEmbed = sys.modules['mmseg.models.utils.embed'].PatchEmbed
    return ns
