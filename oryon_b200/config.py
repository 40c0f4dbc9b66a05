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


_OVERRIDE_LOADER = None


def _override_loader():
    """A YAML loader whose plain scalars resolve the way hydra's override grammar does: ``true`` / ``false`` (any case) are
    booleans, ``null`` is None, ``[+-]?(0|[1-9][0-9_]*)`` an int, decimal / exponent forms (and ``inf`` / ``nan``) a float -- and
    everything else stays a STRING.  PyYAML's own resolvers follow YAML 1.1, where ``yes`` / ``no`` / ``on`` / ``off`` are
    booleans, ``007`` is the integer 7 and ``1:30`` a sexagesimal number; hydra keeps all of those as the strings the reference
    compares against (``test.add_description`` in 'yes' / 'no' / 'wrong' / 'desconly', datasets.py)."""
    global _OVERRIDE_LOADER
    if _OVERRIDE_LOADER is None:
        import yaml

        class Loader(yaml.SafeLoader):
            pass

        Loader.yaml_implicit_resolvers = {}
        Loader.add_implicit_resolver("tag:yaml.org,2002:bool", re.compile(r"^(?:true|false)$", re.I), list("tTfF"))
        Loader.add_implicit_resolver("tag:yaml.org,2002:null", re.compile(r"^(?:null)$", re.I), list("nN"))
        Loader.add_implicit_resolver("tag:yaml.org,2002:int", re.compile(r"^[-+]?(?:0|[1-9][0-9_]*)$"), list("-+0123456789"))
        Loader.add_implicit_resolver(
            "tag:yaml.org,2002:float",
            re.compile(r"^[-+]?(?:(?:[0-9][0-9_]*)?\.[0-9_]+(?:[eE][-+]?[0-9]+)?|[0-9][0-9_]*\.?(?:[eE][-+]?[0-9]+)|[0-9][0-9_]*\.|inf|nan)$", re.I),
            list("-+0123456789.iInN"))
        _OVERRIDE_LOADER = Loader
    return _OVERRIDE_LOADER


def parse_override_value(raw: str) -> Any:
    """The value of a ``key=value`` override with hydra's typing (see ``_override_loader``); lists / dicts in flow syntax and
    quoted strings work as on hydra's command line."""
    import yaml
    if raw == "":
        return None
    value = yaml.load(raw, Loader=_override_loader())
    if isinstance(value, float) and re.fullmatch(r"[-+]?nan", raw.strip(), re.I):
        return float("nan")
    return value


def apply_overrides(cfg: Config, overrides: Iterable[str]) -> Config:
    """``key.sub=value`` items as hydra takes them on the command line (``32``, ``true``, ``null``, ``[192,192]``; ``yes``, ``no``,
    ``007`` stay strings, see ``parse_override_value``); a leading ``+`` (hydra's "add a new key") is accepted."""
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
        cur[keys[-1]] = _wrap(parse_override_value(raw))
    return cfg


def load_config(config_path: str, config_name: str = "config", overrides: Optional[Iterable[str]] = None) -> Config:
    """``config_path``: a YAML file, or a folder holding ``<config_name>.yaml`` (hydra's ``-cp`` / ``-cn``)."""
    import yaml
    path = config_path if os.path.isfile(config_path) else os.path.join(config_path, config_name + ".yaml")
    with open(path) as f:
        cfg = _wrap(yaml.safe_load(f) or {})
    apply_overrides(cfg, overrides or [])
    return _resolve(cfg, cfg)
