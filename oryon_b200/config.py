"""Configuration in the reference's form without hydra / omegaconf: ``configs/config.yaml`` (or the copy an experiment folder
holds, ``python run_test.py -cp exp_data/baseline/ dataset.test.name=nocs test.mask=oracle``, reference README) read with PyYAML,
``${a.b}`` references resolved (the file uses one: ``test.n_corrs: ${dataset.max_corrs}``), hydra-style dotted overrides applied.
The result is an attribute-style mapping -- what ``FPM_Pipeline`` / the dataset readers accept as ``args``."""
from __future__ import annotations

import os
import re
from typing import Any, Iterable, Mapping, Optional

_REF = re.compile(r"\$\{([A-Za-z0-9_.]+)\}")


class Config(dict):
    """A dict whose keys are also attributes (``args.dataset.test.name``), missing ones reading as ``None`` like the entries the
    reference's YAML leaves empty."""

    def __getattr__(self, key: str) -> Any:
        if key.startswith("__"):
            raise AttributeError(key)
        return self.get(key)

    def __setattr__(self, key: str, value: Any) -> None:
        self[key] = value


def _wrap(node: Any) -> Any:
    if isinstance(node, Mapping):
        return Config({k: _wrap(v) for k, v in node.items()})
    if isinstance(node, list):
        return [_wrap(v) for v in node]
    return node


def select(cfg: Mapping, path: str, default: Any = None) -> Any:
    cur: Any = cfg
    for key in path.split("."):
        if not isinstance(cur, Mapping) or key not in cur:
            return default
        cur = cur[key]
    return default if cur is None else cur


def _resolve(cfg: Config, node: Any, depth: int = 0) -> Any:
    if depth > 16:
        raise ValueError("config: circular ${...} reference")
    if isinstance(node, Mapping):
        for k in list(node):
            node[k] = _resolve(cfg, node[k], depth)
        return node
    if isinstance(node, list):
        return [_resolve(cfg, v, depth) for v in node]
    if isinstance(node, str):
        whole = _REF.fullmatch(node)
        if whole:                                              # a value that IS a reference keeps the referenced type
            target = select(cfg, whole.group(1))
            if target is None:
                raise KeyError(f"config: ${{{whole.group(1)}}} refers to nothing")
            return _resolve(cfg, target, depth + 1)
        if _REF.search(node):
            return _REF.sub(lambda m: str(_resolve(cfg, select(cfg, m.group(1), ""), depth + 1)), node)
    return node


def apply_overrides(cfg: Config, overrides: Iterable[str]) -> Config:
    """``key.sub=value`` items as hydra takes them on the command line; values are parsed as YAML scalars (``32``, ``true``,
    ``null``, ``[192,192]``, plain strings); a leading ``+`` (hydra's "add a new key") is accepted."""
    import yaml
    for item in overrides:
        if "=" not in item:
            raise ValueError(f"config override {item!r} is not of the form key=value")
        path, raw = item.lstrip("+").split("=", 1)
        keys = path.split(".")
        cur = cfg
        for k in keys[:-1]:
            if not isinstance(cur.get(k), Mapping):
                cur[k] = Config()
            cur = cur[k]
        cur[keys[-1]] = _wrap(yaml.safe_load(raw)) if raw != "" else None
    return cfg


def load_config(config_path: str, config_name: str = "config", overrides: Optional[Iterable[str]] = None) -> Config:
    """``config_path``: a YAML file, or a folder holding ``<config_name>.yaml`` (hydra's ``-cp`` / ``-cn``)."""
    import yaml
    path = config_path if os.path.isfile(config_path) else os.path.join(config_path, config_name + ".yaml")
    with open(path) as f:
        cfg = _wrap(yaml.safe_load(f) or {})
    apply_overrides(cfg, overrides or [])
    return _resolve(cfg, cfg)
