"""Config / logging / seeding helpers with the reference's names (src/utils.py:12-70).  YAML files
of the reference load unchanged; ``easydict`` is not a dependency (a small attribute-dict of the same
behaviour is defined here)."""
import logging
import os
import random

import numpy as np
import torch
import yaml

class EasyDict(dict):
    """Attribute-access dict (what the third-party ``easydict`` package provides; not a dependency here --
    compat/easydict.py re-exports this class for the reference's ``from easydict import EasyDict``)."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        for k, v in dict(d or {}, **kwargs).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) else x for x in v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


def load_config(cfg_file):
    with open(cfg_file, "r") as fin:
        raw_text = fin.read()
    if "---" in raw_text:
        raise NotImplementedError("grid configs ('---' sections) are dead code in the reference (utils.py:16-23)")
    return [EasyDict(yaml.safe_load(raw_text))]


def _plain(obj):
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    return obj


def save_config(cfg, path):
    with open(os.path.join(path, "config.yaml"), "w") as fo:
        yaml.dump(_plain(cfg), fo)


def set_seed(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def set_logger(save_path):
    log_file = os.path.join(save_path, "run.log")
    logging.basicConfig(format="%(asctime)s %(levelname)-8s %(message)s", level=logging.INFO,
                        datefmt="%Y-%m-%d %H:%M:%S", filename=log_file, filemode="w")
    console = logging.StreamHandler()
    console.setLevel(logging.INFO)
    console.setFormatter(logging.Formatter("%(asctime)s %(levelname)-8s %(message)s"))
    logging.getLogger("").addHandler(console)
