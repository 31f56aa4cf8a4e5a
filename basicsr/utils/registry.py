"""Name -> class registries (same contract as the reference's basicsr/utils/registry.py:4-82:
`register()` as decorator or call, `get`, duplicate names rejected, keyed by `__name__`)."""
from __future__ import annotations


class Registry:
    def __init__(self, name: str):
        self._name = name
        self._table = {}

    def register(self, obj=None):
        def _add(o):
            key = o.__name__
            if key in self._table:
                raise AssertionError(f"An object named '{key}' was already registered in '{self._name}' registry!")
            self._table[key] = o
            return o
        return _add if obj is None else (_add(obj), None)[1]

    def get(self, name: str):
        try:
            return self._table[name]
        except KeyError:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!") from None

    def __contains__(self, name) -> bool:
        return name in self._table

    def __iter__(self):
        return iter(self._table.items())

    def keys(self):
        return self._table.keys()


DATASET_REGISTRY = Registry("dataset")
ARCH_REGISTRY = Registry("arch")
MODEL_REGISTRY = Registry("model")
LOSS_REGISTRY = Registry("loss")
METRIC_REGISTRY = Registry("metric")
