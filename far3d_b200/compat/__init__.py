"""Minimal stand-in for the slice of mmcv / mmdet the reference's plugin surface needs (registries, build_from_cfg,
python-file Config), so `projects/configs/far3d.py`'s `model` dict builds unchanged without mmcv (absent here and
not installable offline).  When real mmcv / mmdet are importable the classes are ALSO registered
into their registries under the same names (force=True), which is how the reference loads a plugin
(projects/mmdet3d_plugin/__init__.py:1-9, tools/test.py:134-155)."""
import copy
import importlib.util
import os


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._modules[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)

    def __contains__(self, key):
        return key in self._modules


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or 'type' not in cfg:
        raise KeyError(f'cfg must be a dict with a "type" key, got {cfg!r}')
    args = copy.deepcopy(dict(cfg))
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop('type')
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f'{t} is not in the {registry.name} registry')
    return cls(**args)


DETECTORS = Registry('detector')
BACKBONES = Registry('backbone')
NECKS = Registry('neck')
HEADS = Registry('head')
TRANSFORMER = Registry('Transformer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')
TRANSFORMER_LAYER = Registry('transformerLayer')
ATTENTION = Registry('attention')
FEEDFORWARD_NETWORK = Registry('feed-forward Network')
BBOX_CODERS = Registry('bbox_coder')

_MM_TARGETS = {   # our registry -> (module path, attribute) of the real one, if installed
    'detector': ('mmdet.models.builder', 'DETECTORS'), 'backbone': ('mmdet.models.builder', 'BACKBONES'),
    'neck': ('mmdet.models.builder', 'NECKS'), 'head': ('mmdet.models.builder', 'HEADS'),
    'Transformer': ('mmdet.models.utils.builder', 'TRANSFORMER'),
    'transformer-layers sequence': ('mmcv.cnn.bricks.registry', 'TRANSFORMER_LAYER_SEQUENCE'),
    'transformerLayer': ('mmcv.cnn.bricks.registry', 'TRANSFORMER_LAYER'),
    'attention': ('mmcv.cnn.bricks.registry', 'ATTENTION'),
    'bbox_coder': ('mmdet.core.bbox.builder', 'BBOX_CODERS'),
}


def mirror_into_mmcv():
    """Register every class also into the real mmcv/mmdet registries when those packages exist."""
    done = []
    for reg in (DETECTORS, BACKBONES, NECKS, HEADS, TRANSFORMER, TRANSFORMER_LAYER_SEQUENCE, TRANSFORMER_LAYER,
                ATTENTION, BBOX_CODERS):
        mod, attr = _MM_TARGETS[reg.name]
        try:
            real = getattr(importlib.import_module(mod), attr)
        except Exception:
            continue
        for key, cls in reg._modules.items():
            real.register_module(name=key, force=True, module=cls)
            done.append(key)
    return done


class Config(dict):
    """`Config.fromfile(path)`: executes a python config file (mmcv style) and exposes its top-level names.
    `_base_` files that do not exist (the reference points at ../../../mmdetection3d/..., far3d.py:1-3) are skipped."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @staticmethod
    def fromfile(path):
        ns = {}
        with open(path) as f:
            code = compile(f.read(), path, 'exec')
        exec(code, ns)
        cfg = Config()
        for base in ns.get('_base_', []) if isinstance(ns.get('_base_', []), (list, tuple)) else [ns['_base_']]:
            bp = os.path.join(os.path.dirname(path), base)
            if os.path.exists(bp):
                cfg.update(Config.fromfile(bp))
        cfg.update({k: v for k, v in ns.items() if not k.startswith('__') and k != '_base_'
                    and not callable(v) and not isinstance(v, type(os))})
        return cfg
