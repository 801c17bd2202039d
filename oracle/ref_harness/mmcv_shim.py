"""Minimal stand-in for the mmcv 1.x bricks the S4Former hot path imports.

TEST INFRASTRUCTURE ONLY.  This file lets ``oracle/ref_harness/load_reference.py``
import the reference's own ``vit.py``, ``setr_up_head.py``, ``decode_head.py``,
``cross_entropy_loss.py``, ``encoder_decoder.py`` and ``generate_unsup_data.py``
UNMODIFIED from ``/root/reference`` (mmcv itself is not installable in this image:
no wheel, no network -- SURVEY.md section 8(c)).  Only the mmcv bricks are restated,
as the thin wrappers over torch that mmcv-full 1.4.4..1.6.0 defines:

* ``mmcv.cnn.bricks.transformer.MultiheadAttention``: wraps
  ``nn.MultiheadAttention(embed_dims, num_heads, attn_drop, bias=...)``, handles
  ``batch_first`` by transposing, returns ``identity + dropout(proj_drop(out))``.
  The reference reads ``layer.attn.self_attn`` (vit.py:550), which only a patched
  mmcv provides; we store the head-averaged weights returned by torch there.
* ``mmcv.cnn.bricks.transformer.FFN``: ``Sequential(Sequential(Linear, act, Dropout),
  Linear, Dropout)`` + identity.
* ``mmcv.cnn.ConvModule``: conv (no bias when a norm follows) -> norm ('bn') -> ReLU.
* ``build_norm_layer``: LN -> ('ln', nn.LayerNorm), BN/SyncBN -> ('bn', nn.BatchNorm2d)
  (tools/train.py:207-213 reverts SyncBN to BN when not distributed).
* ``Registry`` / ``BaseModule`` / ``ModuleList`` / init helpers.

Nothing here is imported by the product package.
"""
import copy
import math
import sys
import types
import warnings

import torch
import torch.nn as nn


# --------------------------------------------------------------------------- registry
class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name = name
        self._module_dict = {}
        self.parent = parent

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, default_args=None):
        cfg = dict(copy.deepcopy(cfg))
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        typ = cfg.pop('type')
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f'{typ} is not in the {self.name} registry')
        return cls(**cfg)


MODELS = Registry('model')
ATTENTION = Registry('attention')


# --------------------------------------------------------------------------- init helpers
def constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.normal_(module.weight, mean, std)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0,
                 distribution='normal'):
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


# --------------------------------------------------------------------------- runner bits
class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    @property
    def is_init(self):
        return self._is_init

    def _apply_init_cfg(self):
        cfgs = self.init_cfg
        if cfgs is None:
            return
        if isinstance(cfgs, dict):
            cfgs = [cfgs]
        for cfg in cfgs:
            typ = cfg.get('type')
            layers = cfg.get('layer')
            if isinstance(layers, str):
                layers = [layers]
            override = cfg.get('override')

            def _do(m):
                if typ == 'Constant':
                    constant_init(m, cfg.get('val', 0), cfg.get('bias', 0))
                elif typ == 'Normal':
                    normal_init(m, cfg.get('mean', 0), cfg.get('std', 1), cfg.get('bias', 0))
                elif typ == 'Kaiming':
                    kaiming_init(m)
            if layers:
                for m in self.modules():
                    if m.__class__.__name__ in layers:
                        _do(m)
            if override is not None:
                ovs = override if isinstance(override, list) else [override]
                for ov in ovs:
                    m = getattr(self, ov['name'], None)
                    if m is not None:
                        _do(m)

    def init_weights(self):
        if not self._is_init:
            if self.init_cfg and not (isinstance(self.init_cfg, dict)
                                      and self.init_cfg.get('type') == 'Pretrained'):
                self._apply_init_cfg()
            for m in self.children():
                if hasattr(m, 'init_weights'):
                    m.init_weights()
            self._is_init = True


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class CheckpointLoader:
    @staticmethod
    def load_checkpoint(filename, map_location=None, logger=None):
        return torch.load(filename, map_location=map_location)


def load_state_dict(module, state_dict, strict=False, logger=None):
    module.load_state_dict(state_dict, strict=strict)


def auto_fp16(apply_to=None, out_fp32=False):
    def wrapper(func):
        return func
    return wrapper


def force_fp32(apply_to=None, out_fp16=False):
    def wrapper(func):
        return func
    return wrapper


# --------------------------------------------------------------------------- cnn bricks
def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if typ == 'LN':
        abbr, layer = 'ln', nn.LayerNorm(num_features, **cfg)
    elif typ in ('BN', 'BN2d', 'SyncBN'):
        # SyncBN == BN in a single process (tools/train.py:207-213 does the same revert)
        abbr, layer = 'bn', nn.BatchNorm2d(num_features, **cfg)
    else:
        raise KeyError(typ)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type='Conv2d'))
    typ = cfg.pop('type')
    assert typ in ('Conv2d', 'Conv', None)
    return nn.Conv2d(*args, **kwargs, **cfg)


def build_activation_layer(cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ == 'ReLU':
        return nn.ReLU(**cfg)
    if typ == 'GELU':
        return nn.GELU()
    raise KeyError(typ)


def build_dropout(cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ == 'Dropout':
        return nn.Dropout(p=cfg.get('drop_prob', 0.5))
    if typ == 'DropPath':
        assert cfg.get('drop_prob', 0.) == 0.
        return nn.Identity()
    raise KeyError(typ)


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias='auto', conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), inplace=True, with_spectral_norm=False,
                 padding_mode='zeros', order=('conv', 'norm', 'act')):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=padding, dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            act_cfg_ = dict(act_cfg)
            if act_cfg_['type'] == 'ReLU':
                act_cfg_.setdefault('inplace', inplace)
            self.activate = build_activation_layer(act_cfg_)
        self.init_weights()

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.with_norm else None

    def init_weights(self):
        kaiming_init(self.conv, a=0, nonlinearity='relu')
        if self.with_norm:
            constant_init(self.norm, 1, bias=0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = self.norm(x)
        if self.with_activation:
            x = self.activate(x)
        return x


class MultiheadAttention(BaseModule):
    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_pos=None, attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if self.batch_first:
            query = query.transpose(0, 1)
            key = key.transpose(0, 1)
            value = value.transpose(0, 1)
        out, w = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                           key_padding_mask=key_padding_mask)
        self.self_attn = w  # patched-mmcv attribute read by vit.py:550
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        self.embed_dims = embed_dims
        self.activate = build_activation_layer(act_cfg)
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(in_channels, feedforward_channels),
                                     self.activate, nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


def to_2tuple(x):
    if isinstance(x, (tuple, list)):
        return tuple(x)
    return (x, x)


def install():
    """Register synthetic ``mmcv`` modules in ``sys.modules``."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mmcv = mod('mmcv', __version__='1.6.0')
    mmcv.load = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    cnn = mod('mmcv.cnn', build_norm_layer=build_norm_layer, ConvModule=ConvModule,
              build_conv_layer=build_conv_layer, MODELS=MODELS, constant_init=constant_init,
              kaiming_init=kaiming_init, normal_init=normal_init,
              build_activation_layer=build_activation_layer)
    bricks = mod('mmcv.cnn.bricks')
    transformer = mod('mmcv.cnn.bricks.transformer', FFN=FFN,
                      MultiheadAttention=MultiheadAttention)
    registry = mod('mmcv.cnn.bricks.registry', ATTENTION=ATTENTION)
    cnn_utils = mod('mmcv.cnn.utils')
    weight_init = mod('mmcv.cnn.utils.weight_init', constant_init=constant_init,
                      kaiming_init=kaiming_init, trunc_normal_=trunc_normal_,
                      normal_init=normal_init)
    runner = mod('mmcv.runner', BaseModule=BaseModule, CheckpointLoader=CheckpointLoader,
                 ModuleList=ModuleList, Sequential=Sequential, load_state_dict=load_state_dict,
                 auto_fp16=auto_fp16, force_fp32=force_fp32)
    base_module = mod('mmcv.runner.base_module', BaseModule=BaseModule, ModuleList=ModuleList,
                      Sequential=Sequential)
    utils = mod('mmcv.utils', Registry=Registry, to_2tuple=to_2tuple)
    mmcv.cnn, mmcv.runner, mmcv.utils = cnn, runner, utils
    cnn.bricks, cnn.utils = bricks, cnn_utils
    bricks.transformer, bricks.registry = transformer, registry
    cnn_utils.weight_init = weight_init
    runner.base_module = base_module
    # Python >= 3.10 removed collections.Mapping (structual_utils.py:2 imports it)
    import collections
    import collections.abc
    for n in ('Mapping', 'Sequence'):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    return mmcv
