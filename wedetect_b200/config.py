"""A small reader for the reference's mmengine-style python configs (config/wedetect_*.py).

Supports what those files use: `_base_ = [...]` inheritance (dict merge, `_delete_` honoured), plain python
assignments evaluated in order, and `--cfg-options key.sub=value` overrides (infer_wedetect.py:88-97,149-151).
If the real `mmengine` is importable its `Config` is used instead (same call surface).
"""
import ast
import copy
import os


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


def _merge(base, new):
    out = copy.deepcopy(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            v = copy.deepcopy(v)
            if isinstance(v, dict):
                v.pop("_delete_", None)
            out[k] = v
    return out


def _load_file(path):
    ns = {}
    with open(path) as f:
        src = f.read()
    exec(compile(src, path, "exec"), {"__file__": path}, ns)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not callable(v) and type(v).__name__ != "module"}
    bases = cfg.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load_file(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


class Config(ConfigDict):
    @staticmethod
    def fromfile(path):
        try:
            from mmengine.config import Config as MMConfig  # noqa: F401
            return MMConfig.fromfile(path)
        except ImportError:
            pass
        c = Config(_wrap(_load_file(path)))
        object.__setattr__(c, "_filename", path)
        return c

    def merge_from_dict(self, options):
        for key, val in options.items():
            d = self
            parts = key.split(".")
            for p in parts[:-1]:
                d = d.setdefault(p, ConfigDict())
            d[parts[-1]] = _wrap(val)


def parse_cfg_options(items):
    """['a.b=1', 'c=[1,2]'] -> {'a.b': 1, 'c': [1, 2]} (DictAction semantics for the simple cases)."""
    out = {}
    for it in items or []:
        k, v = it.split("=", 1)
        try:
            out[k] = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            out[k] = v
    return out
