"""Minimal stand-ins for mmengine.structures.InstanceData / mmdet DetDataSample.

Only the surface that the reference's entry points touch (infer_wedetect.py:113-126): attribute access,
boolean / index `__getitem__` applied to every field, `.cpu()`, `.numpy()`, `len()`; a data sample exposes
its metainfo keys as attributes (yolo_world.py:94 relies on `hasattr(sample, 'texts')`).
When the caller hands in real mmdet / mmengine samples (the unchanged mmcv test pipeline produces them), predictions are
attached as real `mmengine.structures.InstanceData` (`instances_for`), because `DetDataSample.pred_instances` type-checks.
"""
import torch


class InstanceData:
    def __init__(self, **fields):
        object.__setattr__(self, "_fields", {})
        for k, v in fields.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        self._fields[k] = v

    def __getattr__(self, k):
        f = object.__getattribute__(self, "_fields")
        if k in f:
            return f[k]
        raise AttributeError(k)

    def __contains__(self, k):
        return k in self._fields

    def keys(self):
        return list(self._fields)

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0

    def __getitem__(self, idx):
        if isinstance(idx, str):          # mmengine's InstanceData also answers field names (infer_wedetect.py:126-128: pred_instances['bboxes'])
            return self._fields[idx]
        return InstanceData(**{k: v[idx] for k, v in self._fields.items()})

    def _map(self, fn):
        return InstanceData(**{k: fn(v) if isinstance(v, torch.Tensor) else v for k, v in self._fields.items()})

    def cpu(self):
        return self._map(lambda t: t.cpu())

    def to(self, *a, **kw):
        return self._map(lambda t: t.to(*a, **kw))

    def numpy(self):
        return InstanceData(**{k: v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v for k, v in self._fields.items()})

    def __repr__(self):
        return "InstanceData(" + ", ".join(f"{k}={tuple(v.shape) if hasattr(v, 'shape') else v}" for k, v in self._fields.items()) + ")"


class DetDataSample:
    def __init__(self, metainfo=None):
        object.__setattr__(self, "_meta", dict(metainfo or {}))
        object.__setattr__(self, "_data", {})

    @property
    def metainfo(self):
        return dict(self._meta)

    def set_metainfo(self, meta):
        self._meta.update(meta)

    def __setattr__(self, k, v):
        self._data[k] = v

    def __getattr__(self, k):
        d = object.__getattribute__(self, "_data")
        if k in d:
            return d[k]
        m = object.__getattribute__(self, "_meta")
        if k in m:
            return m[k]
        raise AttributeError(k)

    def get(self, k, default=None):
        try:
            return getattr(self, k)
        except AttributeError:
            return default


def instances_for(sample, **fields):
    """InstanceData of the flavour `sample` expects: mmengine's own class for real mmdet / mmengine samples, ours otherwise."""
    mod = type(sample).__module__ if sample is not None else ""
    if mod.startswith(("mmdet", "mmengine")):
        from mmengine.structures import InstanceData as MMInstanceData
        return MMInstanceData(**fields)
    return InstanceData(**fields)
