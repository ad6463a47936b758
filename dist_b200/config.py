"""Hierarchical YAML configuration with attribute access.

Same schema and merge rules as the reference's ``utils/config.py:16-265`` so that an
unchanged ``configs/projects/dist/**.yaml`` selects this implementation:

* ``configs/pool/base.yaml`` holds the defaults every config is merged onto
  (``utils/config.py:79-93``).
* A file may name one parent with ``_BASE`` or a pair ``_BASE_RUN`` + ``_BASE_MODEL``
  (``utils/config.py:111-150``); children override parents key by key, dictionaries merge
  recursively, and keys containing ``BASE`` are not propagated upwards
  (``utils/config.py:154-175``).
* ``KEY.SUB value`` pairs on the command line override existing keys, at most four levels deep
  (``utils/config.py:177-232``).  Deviation: values are parsed as YAML scalars, so ``16`` stays an
  integer (the reference stores the raw string for nested keys).
* Strings of the form ``8e-6`` become floats (``utils/config.py:245-246``).

Unlike the reference the loader has no side effects (no checkpoint directory is created) and
the location of ``base.yaml`` is found relative to the config file or this repository, not
the current working directory.
"""

import argparse
import copy
import json
import os

import yaml

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_BASE_KEYS = ("_BASE", "_BASE_RUN", "_BASE_MODEL")


def _read_yaml(path):
    with open(path, "r") as f:
        data = yaml.load(f.read(), Loader=yaml.SafeLoader)
    return data if data is not None else {}


def _merge(base, new, preserve_base=False):
    """Overlay ``new`` on ``base`` in place (``utils/config.py:154-175``)."""
    for key, val in new.items():
        if key in base:
            if isinstance(val, dict) and isinstance(base[key], dict):
                _merge(base[key], val)
            else:
                base[key] = val
        elif "BASE" not in key or preserve_base:
            base[key] = val
    return base


def _resolve(parent_file, rel):
    if os.path.isabs(rel):
        return rel
    return os.path.normpath(os.path.join(os.path.dirname(parent_file), rel))


def _apply_overrides(cfg, opts):
    """``KEY.SUB value`` pairs; every key on the path must already exist."""
    if not opts:
        return cfg
    assert len(opts) % 2 == 0, "Override list {} has odd length: {}.".format(opts, len(opts))
    for key, raw in zip(opts[0::2], opts[1::2]):
        parts = key.split(".")
        assert len(parts) <= 4, "Key depth error. \nMaximum depth: 3\n Get depth: {}".format(len(parts))
        node = cfg
        for part in parts[:-1]:
            assert isinstance(node, dict) and part in node, "Non-existant key: {}.".format(key)
            node = node[part]
        assert isinstance(node, dict) and parts[-1] in node, "Non-existant key: {}.".format(key)
        node[parts[-1]] = yaml.load(raw, Loader=yaml.SafeLoader) if isinstance(raw, str) else raw
    return cfg


def _load_tree(path, opts):
    cfg = _read_yaml(path)
    if not any(k in cfg for k in _BASE_KEYS):
        return cfg
    if "_BASE" in cfg:
        parent = _load_tree(_resolve(path, cfg["_BASE"]), opts)
        cfg = _merge(parent, cfg)
    else:
        if "_BASE_RUN" in cfg:
            parent = _load_tree(_resolve(path, cfg["_BASE_RUN"]), opts)
            cfg = _merge(parent, cfg, preserve_base=True)
        if "_BASE_MODEL" in cfg:
            parent = _load_tree(_resolve(path, cfg["_BASE_MODEL"]), opts)
            cfg = _merge(parent, cfg)
    # the reference re-applies the overrides at every inheritance level; keys that do not exist
    # yet at this level are applied once the level that introduces them is merged.
    present = []
    for key, raw in zip(opts[0::2], opts[1::2]):
        node, ok = cfg, True
        for part in key.split("."):
            if isinstance(node, dict) and part in node:
                node = node[part]
            else:
                ok = False
                break
        if ok:
            present += [key, raw]
    return _apply_overrides(cfg, present)


def _find_base_yaml(cfg_file):
    probe = os.path.dirname(os.path.abspath(cfg_file))
    while True:
        cand = os.path.join(probe, "pool", "base.yaml")
        if os.path.exists(cand):
            return cand
        cand = os.path.join(probe, "configs", "pool", "base.yaml")
        if os.path.exists(cand):
            return cand
        up = os.path.dirname(probe)
        if up == probe:
            break
        probe = up
    return os.path.join(_REPO_ROOT, "configs", "pool", "base.yaml")


class Config(object):
    """Attribute view of the merged dictionary; nested dictionaries become nested ``Config``s."""

    def __init__(self, load=True, cfg_dict=None, cfg_level=None):
        self._level = "cfg" + ("." + cfg_level if cfg_level is not None else "")
        if load:
            self.args = self._parse_args()
            cfg_dict = load_cfg_dict(self.args.cfg_file, self.args.opts)
            self.cfg_dict = cfg_dict
        self._update_dict(cfg_dict if cfg_dict is not None else {})

    @classmethod
    def from_file(cls, cfg_file, opts=None):
        cfg_dict = load_cfg_dict(cfg_file, list(opts) if opts else [])
        cfg = cls(load=False, cfg_dict=cfg_dict)
        cfg.cfg_dict = cfg_dict
        return cfg

    @classmethod
    def from_dict(cls, cfg_dict):
        cfg_dict = copy.deepcopy(cfg_dict)
        cfg = cls(load=False, cfg_dict=cfg_dict)
        cfg.cfg_dict = cfg_dict
        return cfg

    @staticmethod
    def _parse_args():
        parser = argparse.ArgumentParser(description="dist_b200 configuration")
        parser.add_argument("--cfg", dest="cfg_file", help="Path to the configuration file", default=None)
        parser.add_argument("--init_method", default="tcp://127.0.0.1:9999", type=str,
                            help="Initialization method, includes TCP or shared file-system")
        parser.add_argument("opts", help="other configurations", default=None, nargs=argparse.REMAINDER)
        return parser.parse_args()

    def _update_dict(self, cfg_dict):
        def convert(key, elem):
            if type(elem) is dict:
                return key, Config(load=False, cfg_dict=elem, cfg_level=key)
            if type(elem) is str and elem[1:3] == "e-":
                elem = float(elem)
            return key, elem

        self.__dict__.update(dict(convert(k, v) for k, v in cfg_dict.items()))

    def get_args(self):
        return self.args

    def dump(self):
        return json.dumps(self.cfg_dict, indent=2)

    def __repr__(self):
        return "{}\n".format(self.dump()) if hasattr(self, "cfg_dict") else object.__repr__(self)

    def deep_copy(self):
        return copy.deepcopy(self)


def load_cfg_dict(cfg_file, opts=None):
    assert cfg_file is not None
    opts = list(opts) if opts else []
    base = _read_yaml(_find_base_yaml(cfg_file))
    tree = _load_tree(cfg_file, opts)
    merged = _merge(base, tree)
    # overrides that only exist after the merge with base.yaml
    return _apply_overrides(merged, opts)
