"""Hydra-style config composition without hydra (not installable here).

Reproduces what the reference's ``@hydra.main(config_path="./config",
config_name="config")`` + ``OmegaConf.to_container`` delivers for the command
lines in its README (``+alg=pqn_minatar alg.NUM_ENVS=4096 NUM_SEEDS=8``):

* root ``config.yaml`` composed with the group file ``alg/<name>.yaml`` under
  the ``alg`` key (``+alg=<name>``), group values merged over root's ``alg`` map;
* ``KEY=V`` / ``alg.KEY=V`` dotted overrides (``+`` / ``++`` prefixes accepted);
* OmegaConf scalar typing: ``1e7`` -> float (PyYAML alone would keep a string),
  ``null`` -> None, ``true/false`` -> bool.
"""
from __future__ import annotations

import os
import re

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG_DIR = os.path.join(_HERE, "config")
_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")
_INT = re.compile(r"^[+-]?\d+$")


def _coerce(v):
    if isinstance(v, str):
        s = v.strip()
        if _INT.match(s):
            return int(s)
        if _FLOAT.match(s):
            return float(s)
        low = s.lower()
        if low in ("null", "none", "~"):
            return None
        if low == "true":
            return True
        if low == "false":
            return False
        return v
    if isinstance(v, dict):
        return {k: _coerce(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_coerce(x) for x in v]
    return v


def _load(path):
    with open(path) as f:
        return _coerce(yaml.safe_load(f) or {})


def _set(cfg, dotted, value):
    keys = dotted.split(".")
    d = cfg
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def compose(overrides=(), config_dir=CONFIG_DIR, config_name="config"):
    cfg = _load(os.path.join(config_dir, f"{config_name}.yaml"))
    rest = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        key = key.lstrip("+")
        if key == "alg":
            group = _load(os.path.join(config_dir, "alg", f"{val}.yaml"))
            merged = dict(cfg.get("alg") or {})
            merged.update(group)
            cfg["alg"] = merged
        else:
            rest.append((key, val))
    for key, val in rest:
        _set(cfg, key, _coerce(yaml.safe_load(val) if val != "" else ""))
    return cfg


def save_yaml(config: dict, path: str):
    with open(path, "w") as f:
        yaml.safe_dump(config, f, sort_keys=False)
