"""Checkpoint wire format of purejaxql/utils/save_load.py: safetensors whose
keys are the flax parameter paths joined with ",", one file per seed."""
from __future__ import annotations

import os
from typing import Dict, Union

import torch
from safetensors.torch import load_file, save_file


def _flatten(d: Dict, prefix=()):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + (k,)))
        else:
            out[",".join(prefix + (k,))] = v
    return out


def _unflatten(flat: Dict):
    tree: Dict = {}
    for k, v in flat.items():
        d = tree
        parts = k.split(",")
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return tree


def save_params(params: Dict, filename: Union[str, os.PathLike]) -> None:
    flat = {k: torch.as_tensor(v).detach().cpu().contiguous().clone() for k, v in _flatten(params).items()}
    save_file(flat, str(filename))


def load_params(filename: Union[str, os.PathLike]) -> Dict:
    return _unflatten(load_file(str(filename)))
