"""TOML configuration front-end for ensembles.

Mirrors ``rscm.config`` of the reference (python/rscm/config/loader.py:28-128 ``deep_merge`` / ``load_config`` /
``load_config_layers``; registry.py:31-100 ``ComponentRegistry``; builder.py:19-108 ``build_model`` /
``build_two_layer_model``; models/two_layer.py:53-103 parameter metadata) so that the reference's ``configs/two-layer/*.toml``
files work unchanged, and adds what makes an ensemble of a configured model a one-liner:

    cfg = load_config_layers("configs/two-layer/defaults.toml", "configs/two-layer/tuning/high-ecs.toml")
    model = build_model(cfg)                                             # reference behaviour: one member
    ens, names = build_ensemble(cfg, vary=["lambda0", "efficacy"])       # the same model, those parameters per member
    priors = prior_parameter_set("TwoLayer", names)                      # Uniform over the metadata's typical ranges
"""

from __future__ import annotations

import logging
import tomllib
from pathlib import Path
from typing import Any

import numpy as np

from .core import ModelBuilder, TimeAxis

logger = logging.getLogger(__name__)

__all__ = ["deep_merge", "load_config", "load_config_layers", "ComponentRegistry", "component_registry", "PARAMETER_METADATA",
           "model_builder_from_config", "build_model", "build_two_layer_model", "build_ensemble", "prior_parameter_set"]

_KNOWN_TOP_LEVEL = {"schema", "time", "components", "inputs", "outputs", "model", "initial_values"}


def deep_merge(base: dict[str, Any], override: dict[str, Any]) -> dict[str, Any]:
    """Nested dicts merge recursively; every other value (lists included) is replaced by the override."""
    merged = dict(base)
    for key, value in override.items():
        if isinstance(value, dict) and isinstance(merged.get(key), dict):
            merged[key] = deep_merge(merged[key], value)
        else:
            merged[key] = value
    return merged


def load_config(path: str | Path) -> dict[str, Any]:
    with Path(path).open("rb") as f:
        config = tomllib.load(f)
    unknown = sorted(set(config) - _KNOWN_TOP_LEVEL)
    if unknown:
        logger.warning("Unknown configuration keys in %s: %s. These will be ignored.", path, ", ".join(unknown))
    return config


def load_config_layers(*paths: str | Path) -> dict[str, Any]:
    """Later files override earlier ones."""
    merged: dict[str, Any] = {}
    for p in paths:
        merged = deep_merge(merged, load_config(p))
    return merged


class ComponentRegistry:
    """name -> ``<Kind>Builder`` class (anything with ``from_parameters(dict).build()``)."""

    def __init__(self) -> None:
        self._builders: dict[str, type] = {}

    def register(self, name: str, builder_class: type) -> None:
        if name in self._builders:
            raise ValueError(f"Component '{name}' is already registered")
        self._builders[name] = builder_class

    def get(self, name: str) -> type:
        try:
            return self._builders[name]
        except KeyError:
            raise KeyError(f"Unknown component: {name!r}. Available: {', '.join(sorted(self._builders)) or '(none)'}") from None

    def list(self) -> list[str]:
        return sorted(self._builders)

    def __contains__(self, name: str) -> bool:
        return name in self._builders


component_registry = ComponentRegistry()


def _register_defaults() -> None:
    from . import components, magicc, two_layer
    for mod in (two_layer, components, magicc):
        for attr in dir(mod):
            if attr.endswith("Builder") and attr != "ComponentBuilder":
                cls = getattr(mod, attr)
                name = getattr(cls, "TYPE_NAME", "") or attr[: -len("Builder")]
                if name and name not in component_registry:
                    component_registry.register(name, cls)


_register_defaults()

# (default, hard range, typical range) — python/rscm/config/models/two_layer.py:53-103
PARAMETER_METADATA: dict[str, dict[str, dict[str, Any]]] = {
    "TwoLayer": {
        "lambda0": {"default": 1.0, "range": (0.1, 5.0), "typical_range": (0.8, 1.5), "unit": "W/(m² K)"},
        "a": {"default": 0.0, "range": (0.0, 1.0), "typical_range": (0.0, 0.1), "unit": "W/(m² K²)"},
        "efficacy": {"default": 1.0, "range": (0.5, 3.0), "typical_range": (1.0, 1.8), "unit": "dimensionless"},
        "eta": {"default": 0.7, "range": (0.1, 2.0), "typical_range": (0.5, 1.0), "unit": "W/(m² K)"},
        "heat_capacity_surface": {"default": 8.0, "range": (1.0, 50.0), "typical_range": (5.0, 15.0), "unit": "W yr/(m² K)"},
        "heat_capacity_deep": {"default": 100.0, "range": (10.0, 500.0), "typical_range": (50.0, 200.0), "unit": "W yr/(m² K)"},
    },
}


def _validate(type_name: str, params: dict[str, float]) -> None:
    """validate_parameters (python/rscm/config/validation.py): every value inside its metadata range."""
    errors = []
    for name, value in params.items():
        meta = PARAMETER_METADATA.get(type_name, {}).get(name)
        if meta and not (meta["range"][0] <= float(value) <= meta["range"][1]):
            errors.append(f"{name}={value} outside [{meta['range'][0]}, {meta['range'][1]}]")
    if errors:
        raise ValueError(f"Invalid parameters: {errors}")


def model_builder_from_config(config: dict[str, Any]) -> tuple[ModelBuilder, list[tuple[str, str]]]:
    """``ModelBuilder`` for a TOML dict: time axis (annual, ``[time] start..end``), every ``[components.<key>]`` in file order
    (``type`` looked up in the registry), ``[initial_values]``.  Returns the builder and the (key, type) list."""
    builder = ModelBuilder()
    time_cfg = config.get("time") or {}
    if time_cfg:
        start, end = time_cfg.get("start", 1750), time_cfg.get("end", 2100)
        builder = builder.with_time_axis(TimeAxis.from_values(np.arange(start, end + 1, dtype=float)))
    added = []
    for key, comp in (config.get("components") or {}).items():
        type_name = comp.get("type")
        if not type_name:
            raise ValueError(f"components.{key}: missing 'type'")
        params = dict(comp.get("parameters") or {})
        _validate(type_name, params)
        builder = builder.with_rust_component(component_registry.get(type_name).from_parameters(params).build())
        added.append((key, type_name))
    initial = dict(config.get("initial_values") or {})
    if initial:
        builder = builder.with_initial_values(initial)
    return builder, added


def build_two_layer_model(config: dict[str, Any]):
    """builder.py:50-108: the ``climate`` component's parameters, annual axis, zero initial temperatures."""
    return _two_layer_builder(config).build()


def _two_layer_builder(config: dict[str, Any]) -> ModelBuilder:
    params = dict(((config.get("components") or {}).get("climate") or {}).get("parameters") or {})
    cfg = {"time": config.get("time") or {}, "components": {"climate": {"type": "TwoLayer", "parameters": params}},
           "initial_values": {"Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0, **(config.get("initial_values") or {})}}
    return model_builder_from_config(cfg)[0]


def build_model(config: dict[str, Any]):
    """builder.py:19-47: dispatch on ``[model] type``; only ``two-layer`` is defined by the reference."""
    model_type = (config.get("model") or {}).get("type", "")
    if model_type == "two-layer":
        return build_two_layer_model(config)
    raise ValueError(f"Unknown model type: {model_type!r}")


def build_ensemble(config: dict[str, Any], vary: list[str] | dict[str, str], *, dtype: str = "f64", device: int = -1):
    """The configured model as a GPU ensemble.  ``vary`` names the parameters that differ per member: either bare field
    names of a single-component model (``["lambda0", "efficacy"]``) or ``{column: "Type.field" | "initial:Variable"}``.
    Returns ``(Ensemble, column names in parameter-matrix order)``."""
    model_type = (config.get("model") or {}).get("type", "")
    if model_type == "two-layer":
        builder, comps = _two_layer_builder(config), [("climate", "TwoLayer")]
    else:
        builder, comps = model_builder_from_config(config)
    if isinstance(vary, dict):
        bindings = dict(vary)
    else:
        if len(comps) != 1:
            raise ValueError("bare parameter names need a single-component model; give {'column': 'Type.field'} bindings")
        bindings = {name: f"{comps[0][1]}.{name}" for name in vary}
    ens = builder.build_ensemble(dtype=dtype, device=device).bind_parameters(bindings)
    return ens, list(bindings)


def prior_parameter_set(type_name: str, names: list[str], which: str = "typical_range"):
    """``ParameterSet`` of Uniform priors over the metadata ranges (``typical_range`` or ``range``) of ``names``."""
    from .calibrate import ParameterSet, Uniform
    ps = ParameterSet()
    for n in names:
        lo, hi = PARAMETER_METADATA[type_name][n][which]
        ps.add(n, Uniform(lo, hi))
    return ps
