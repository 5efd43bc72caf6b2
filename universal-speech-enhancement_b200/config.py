"""Hydra-compatible config surface without Hydra.

The reference composes ``configs/predict.yaml`` with Hydra and instantiates ``_target_`` / ``_partial_`` nodes
(/root/reference/src/predict.py:50-79, configs/model/SGMSE_Large.yaml:1-33).  Hydra / OmegaConf are not in the
offline image, so this module implements the small subset the predict path needs -- defaults-list composition of
config groups, ``key=value`` / ``group=option`` overrides, and recursive ``instantiate`` -- over the same YAML
files and the same keys.  When Hydra is importable the YAML files work with it unchanged.
"""
from __future__ import annotations

import functools
import importlib
import os
import re
from typing import Any, Dict, List, Optional

import yaml

# the reference's class paths resolve to this package's implementations of the same interface
TARGET_ALIASES = {
    "src.models.SGMSE_module.SGMSEModule": "use_b200.sgmse_module.SGMSEModule",
    "src.models.components.sgmse.model_wrapper.ScoreModel": "use_b200.model_wrapper.ScoreModel",
    "src.data.loadwav_datamodule.LoadWavDataModule": "use_b200.predict.LoadWavDataModule",
    "src.models.LSGAN_module.GANModule": "use_b200.gan.GANModule",
    "src.models.components.GAN.generator.ncsnpp.model_wrapper.NCSNPP_Wrapper": "use_b200.gan.NCSNPP_Wrapper",
}


def _locate(path: str):
    path = TARGET_ALIASES.get(path, path)
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(node: Any, **overrides):
    """hydra.utils.instantiate for dict configs: recursive, honours ``_target_`` and ``_partial_``."""
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "_target_" not in node:
        return {k: instantiate(v) for k, v in node.items()}
    kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
    kwargs.update(overrides)
    fn = _locate(node["_target_"])
    if node.get("_partial_", False):
        return functools.partial(fn, **kwargs)
    return fn(**kwargs)


_SCI = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)[eE][+-]?\d+$")


def _fix_scalars(node):
    """PyYAML reads `3e-2` / `5e-4` as strings (YAML 1.1 wants a dot); OmegaConf reads floats.  Follow OmegaConf."""
    if isinstance(node, dict):
        return {k: _fix_scalars(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_fix_scalars(v) for v in node]
    if isinstance(node, str) and _SCI.match(node):
        return float(node)
    return node


def _set(cfg: dict, dotted: str, value):
    keys = dotted.split(".")
    for k in keys[:-1]:
        cfg = cfg.setdefault(k, {})
    cfg[keys[-1]] = value


def compose(config_dir: str, config_name: str = "predict.yaml", overrides: Optional[List[str]] = None) -> Dict:
    """Compose a root config: ``defaults`` list entries ``{group: option}`` load ``<group>/<option>.yaml`` under the
    group key; ``group=option`` overrides swap the option, ``a.b.c=value`` overrides set leaves (YAML-typed)."""
    overrides = list(overrides or [])
    with open(os.path.join(config_dir, config_name)) as f:
        root = yaml.safe_load(f) or {}
    defaults = root.pop("defaults", [])
    groups = {}
    for d in defaults:
        if isinstance(d, dict):
            for g, opt in d.items():
                groups[g] = opt
    leaf = []
    for ov in overrides:
        k, _, v = ov.partition("=")
        k = k.lstrip("+")
        if k in groups or os.path.isdir(os.path.join(config_dir, k)):
            groups[k] = v
        else:
            leaf.append((k, yaml.safe_load(v)))
    cfg: Dict = {}
    for g, opt in groups.items():
        if opt in (None, "null"):
            continue
        p = os.path.join(config_dir, g, f"{opt}.yaml")
        if not os.path.exists(p):
            raise FileNotFoundError(f"config group option {g}={opt} not found ({p})")
        with open(p) as f:
            cfg[g] = yaml.safe_load(f) or {}
    cfg.update(root)
    for k, v in leaf:
        _set(cfg, k, v)
    return _fix_scalars(cfg)
