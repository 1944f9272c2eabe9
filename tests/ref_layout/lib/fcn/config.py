"""Stand-in for the reference's lib/fcn/config.py: one global attribute-dict `cfg` with the keys the path reads, and a
cfg_from_file() that merges a yaml file into it AFTER the modules were imported (the order the reference's tools use)."""
import numpy as np
import yaml


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


cfg = AttrDict()
cfg.INPUT = 'RGBD'
cfg.RNG_SEED = 3
cfg.PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])
cfg.MODE = 'TRAIN'
cfg.TRAIN = AttrDict(NUM_UNITS=64, FUSION_TYPE='add', EMBEDDING_NORMALIZATION=True, EMBEDDING_METRIC='cosine', CLASSES=(0, 1))
cfg.TEST = AttrDict(VISUALIZE=False, CLASSES=())


def _merge(src, dst):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(v, dst[k])
        else:
            dst[k] = v


def cfg_from_file(filename):
    with open(filename) as f:
        _merge(yaml.safe_load(f) or {}, cfg)
