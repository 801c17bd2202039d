"""Registry / builders mirroring ``mmseg/models/builder.py:8-47``.

When mmcv is importable the classes are ALSO registered in mmseg's own ``MODELS`` registry
(``force=True``) by ``s4former_b200.register_into_mmseg()`` so ``configs/setr/*.py`` and
``tools/train.py`` build them unchanged.  Without mmcv a small stand-alone registry with the
same ``build(cfg)`` semantics (``cls(**cfg_without_type)``) is used.
"""
import copy


class Registry:
    def __init__(self, name):
        self.name = name
        self._module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._module_dict and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._module_dict[key] = cls
            return cls
        return _register(module) if module is not None else _register

    def get(self, key):
        return self._module_dict.get(key)

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict):
            raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
        cfg = dict(copy.deepcopy(cfg))
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        if 'type' not in cfg:
            raise KeyError('`cfg` must contain the key "type"')
        typ = cfg.pop('type')
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f'{typ} is not in the {self.name} registry')
        return cls(**cfg)


MODELS = Registry('models')
BACKBONES = NECKS = HEADS = LOSSES = SEGMENTORS = MODELS


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    assert cfg.get('train_cfg') is None or train_cfg is None, \
        'train_cfg specified in both outer field and model field '
    assert cfg.get('test_cfg') is None or test_cfg is None, \
        'test_cfg specified in both outer field and model field '
    return SEGMENTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))
