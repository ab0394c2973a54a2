"""Minimal Hydra-compatible config composition (hydra-core / omegaconf are not installed here).

Covers what the reference's experiment tree uses (`/root/reference/configs/train/*.yaml`, composed by
`@hydra.main(config_path="configs/train", config_name="v1")`, train.py:164):
  * `defaults:` lists with plain names (`- default`), relative groups with a package
    (`- ../model@model: v1`), the structured-schema entry (`- /train/base@_here_`, a no-op here: the
    schema only type-checks) and `_self_` ordering;
  * command-line overrides `a.b.c=value`, `+a.b=value` (add) and `~a.b` (delete), values parsed as YAML;
  * `-cn/--config-name`, `-cd/--config-dir`, `-cp/--config-path`.
The result is a plain nested `dict` wrapped in `Cfg` for attribute access.
"""
from __future__ import annotations

import copy
import os

import yaml


class Cfg(dict):
    """dict with attribute access (enough of DictConfig for train.py)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return Cfg({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def to_container(x):
    if isinstance(x, dict):
        return {k: to_container(v) for k, v in x.items()}
    if isinstance(x, list):
        return [to_container(v) for v in x]
    return x


def _merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


class _Loader(yaml.SafeLoader):
    pass


# YAML 1.1 (PyYAML) does not read "4e-4" as a float; Hydra/OmegaConf do.
import re  # noqa: E402

_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                   |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                   |\.[0-9_]+(?:[eE][-+]?[0-9]+)?
                   |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$""", re.X),
    list("-+0123456789."))


def _load_yaml(path: str) -> dict:
    with open(path) as f:
        return yaml.load(f, Loader=_Loader) or {}


def _find(base_dir: str, name: str) -> str:
    for ext in (".yaml", ".yml"):
        p = os.path.normpath(os.path.join(base_dir, name + ext))
        if os.path.isfile(p):
            return p
    raise FileNotFoundError(f"config '{name}' not found under {base_dir}")


def load_config(path: str) -> dict:
    """Compose one YAML file with its `defaults` list (depth-first, later entries win, `_self_` placement)."""
    raw = _load_yaml(path)
    defaults = raw.pop("defaults", None) or []
    base_dir = os.path.dirname(path)
    out: dict = {}
    self_done = False
    for entry in defaults:
        if entry == "_self_":
            _merge(out, raw)
            self_done = True
            continue
        if isinstance(entry, str):
            group, package, name = "", None, entry
        else:
            (key, name), = entry.items()
            group, _, package = str(key).partition("@")
        if str(group if not isinstance(entry, str) else name).startswith("/"):
            continue  # structured-config schema (ConfigStore): nothing to merge
        if isinstance(entry, str):
            sub = load_config(_find(base_dir, name))
        else:
            sub = load_config(_find(os.path.join(base_dir, group), str(name)))
        if package and package not in ("_here_", "_global_"):
            node = out
            for part in package.split("."):
                node = node.setdefault(part, {})
            _merge(node, sub)
        else:
            _merge(out, sub)
    if not self_done:
        _merge(out, raw)
    return out


def apply_overrides(cfg: dict, overrides: list[str]) -> dict:
    for ov in overrides:
        if ov.startswith("~"):
            parts = ov[1:].split(".")
            node = cfg
            for p in parts[:-1]:
                node = node[p]
            node.pop(parts[-1], None)
            continue
        key, sep, val = ov.partition("=")
        if not sep:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        add = key.startswith("+")
        parts = key.lstrip("+").split(".")
        node = cfg
        for p in parts[:-1]:
            if p not in node or not isinstance(node[p], dict):
                if not add and p not in node:
                    raise KeyError(f"override '{ov}': '{p}' not in config (use +{key}=... to add)")
                node[p] = {} if not isinstance(node.get(p), dict) else node[p]
            node = node[p]
        if not add and parts[-1] not in node:
            raise KeyError(f"override '{ov}': key '{parts[-1]}' not in config (use +{key}=... to add)")
        node[parts[-1]] = yaml.load(val, Loader=_Loader) if val != "" else ""
    return cfg


def parse_cli(argv: list[str], default_dir: str, default_name: str):
    """-> (config_dir, config_name, overrides)."""
    cfg_dir, name, overrides = default_dir, default_name, []
    it = iter(argv)
    for a in it:
        if a in ("-cn", "--config-name"):
            name = next(it)
        elif a.startswith("--config-name="):
            name = a.split("=", 1)[1]
        elif a in ("-cd", "--config-dir", "-cp", "--config-path"):
            cfg_dir = next(it)
        elif a.startswith(("--config-dir=", "--config-path=")):
            cfg_dir = a.split("=", 1)[1]
        else:
            overrides.append(a)
    return cfg_dir, name, overrides


def compose(config_dir: str, config_name: str, overrides: list[str]) -> Cfg:
    cfg = load_config(_find(config_dir, config_name))
    return _wrap(apply_overrides(cfg, overrides))
