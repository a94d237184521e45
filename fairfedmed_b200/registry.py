"""Trainer registry with the Dassl `Registry` contract (Dassl/dassl/utils/registry.py:36-68,
Dassl/dassl/engine/build.py:7-20): `register()` as decorator or call, duplicate names raise KeyError unless
`force=True`, `get(name)` raises KeyError listing the registered names, `build_trainer(cfg)` looks up
`cfg.TRAINER.NAME`."""
from __future__ import annotations


class Registry:
    def __init__(self, name: str):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj, force=False):
        if name in self._obj_map and not force:
            raise KeyError(f'An object named "{name}" was already registered in "{self._name}" registry')
        self._obj_map[name] = obj

    def register(self, obj=None, force=False):
        if obj is None:
            def wrapper(fn_or_class):
                self._do_register(fn_or_class.__name__, fn_or_class, force=force)
                return fn_or_class
            return wrapper
        self._do_register(obj.__name__, obj, force=force)
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f'Object name "{name}" does not exist in "{self._name}" registry')
        return self._obj_map[name]

    def registered_names(self):
        return list(self._obj_map.keys())


TRAINER_REGISTRY = Registry("TRAINER")


def build_trainer(cfg):
    avai = TRAINER_REGISTRY.registered_names()
    if cfg.TRAINER.NAME not in avai:
        raise ValueError(f"TRAINER.NAME must be one of {avai}, got {cfg.TRAINER.NAME!r}")
    return TRAINER_REGISTRY.get(cfg.TRAINER.NAME)(cfg)
