"""String-keyed component registry and config-carrying base object.

The renderer plugin must register and be constructed exactly like the
reference's: ``@RENDERER_REGISTRY.register()`` keyed by class ``__name__``
(pointrix/utils/registry.py:26-51), instantiated as ``cls(cfg, **kwargs)`` where
``cfg`` is parsed against the nested ``Config`` dataclass and ``setup(**kwargs)``
is then called (pointrix/utils/base.py:25-39).  When the real ``pointrix``
package is importable its registry is reused so that ``parse_renderer`` finds
this class; otherwise this OmegaConf-free equivalent is used.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Any, Dict, Iterable, Iterator, Optional, Tuple

import torch


class Registry(Iterable[Tuple[str, Any]]):
    def __init__(self, name: str) -> None:
        self._name = name
        self._obj_map: Dict[str, Any] = {}

    def _do_register(self, name: str, obj: Any, override: bool = False) -> None:
        if not override and name in self._obj_map:
            raise AssertionError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj: Any = None, override: bool = False) -> Any:
        if obj is None:
            def deco(func_or_class: Any) -> Any:
                self._do_register(func_or_class.__name__, func_or_class, override)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj, override)
        return obj

    def get(self, name: str) -> Any:
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name: str) -> bool:
        return name in self._obj_map

    def __iter__(self) -> Iterator[Tuple[str, Any]]:
        return iter(self._obj_map.items())

    def __repr__(self) -> str:
        rows = "\n".join(f"  {k}: {v}" for k, v in self._obj_map.items())
        return f"Registry of {self._name}:\n{rows}"


def get_rank() -> int:
    # pointrix/utils/base.py:11-20
    for key in ("RANK", "LOCAL_RANK", "SLURM_PROCID", "JSM_NAMESPACE_RANK"):
        rank = os.environ.get(key)
        if rank is not None:
            return int(rank)
    return 0


def get_device() -> torch.device:
    return torch.device(f"cuda:{get_rank()}")


def parse_structured(fields: Any, cfg: Optional[Any] = None) -> Any:
    """dict / dataclass / attribute-bag -> instance of the ``fields`` dataclass.
    Unknown keys are an error, as with OmegaConf structured configs
    (pointrix/utils/config.py parse_structured)."""
    if cfg is None:
        return fields()
    if dataclasses.is_dataclass(cfg) and not isinstance(cfg, type):
        cfg = dataclasses.asdict(cfg)
    elif not isinstance(cfg, dict):
        try:  # omegaconf.DictConfig and friends
            cfg = {k: cfg[k] for k in cfg.keys()}
        except Exception as e:  # pragma: no cover
            raise TypeError(f"cannot interpret config of type {type(cfg)}") from e
    names = {f.name: f for f in dataclasses.fields(fields)}
    unknown = set(cfg) - set(names)
    if unknown:
        raise KeyError(f"Key(s) {sorted(unknown)} not in '{fields.__qualname__}'")
    kwargs = {}
    for k, v in cfg.items():
        t = names[k].type
        if t in (int, "int") and not isinstance(v, bool):
            v = int(v)
        elif t in (float, "float"):
            v = float(v)
        elif t in (bool, "bool"):
            v = bool(v)
        kwargs[k] = v
    return fields(**kwargs)


class BaseObject:
    @dataclasses.dataclass
    class Config:
        pass

    cfg: Config

    def __init__(self, cfg: Optional[Any] = None, *args, **kwargs) -> None:
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self.device = get_device()
        self.setup(*args, **kwargs)

    def setup(self, *args, **kwargs) -> None:
        pass


def _renderer_registry() -> Registry:
    try:  # real pointrix present: share its registry so parse_renderer() resolves to us
        from pointrix.model.renderer.msplat import RENDERER_REGISTRY as reg  # type: ignore
        return reg
    except Exception:
        return Registry("RENDERER")


RENDERER_REGISTRY = _renderer_registry()


def register_renderer(cls):
    """Register under ``cls.__name__``; when the real pointrix registry is shared this
    replaces the stock entry of the same name, which is the whole point of a drop-in."""
    RENDERER_REGISTRY._obj_map[cls.__name__] = cls
    return cls


def parse_renderer(cfg: dict, **kwargs):
    """pointrix/model/renderer/__init__.py:4-20 for the msplat backend only."""
    cfg = dict(cfg)
    name = cfg.pop("name")
    return RENDERER_REGISTRY.get(name)(cfg, **kwargs)
