"""Minimal omegaconf stand-in (DictConfig / ListConfig / OmegaConf) -- see ../README.md."""
import copy

import yaml


class ListConfig(list):
    pass


class DictConfig(dict):
    """dict with attribute access; nested dicts/lists are wrapped on construction."""

    def __init__(self, content=None):
        super().__init__()
        for k, v in (content or {}).items():
            dict.__setitem__(self, k, _wrap(v))

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = _wrap(value)

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, _wrap(value))

    def get(self, key, default=None):
        return dict.get(self, key, default)

    def __deepcopy__(self, memo):
        return DictConfig(copy.deepcopy(_unwrap(self), memo))


def _wrap(v):
    if isinstance(v, DictConfig) or isinstance(v, ListConfig):
        return v
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, (list, tuple)):
        return ListConfig(_wrap(x) for x in v)
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    return v


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if isinstance(obj, str):
            obj = yaml.safe_load(obj)
        return _wrap(obj if obj is not None else {})

    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f) or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)

    @staticmethod
    def to_yaml(cfg):
        return yaml.safe_dump(_unwrap(cfg), sort_keys=False)

    @staticmethod
    def update(cfg, dotted_key, value):
        node = cfg
        parts = dotted_key.split(".")
        for p in parts[:-1]:
            if p not in node:
                node[p] = DictConfig()
            node = node[p]
        node[parts[-1]] = value


__all__ = ["DictConfig", "ListConfig", "OmegaConf"]
