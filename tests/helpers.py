"""Test helpers: build the CPU oracle's model from the same ModelBuilder description the
product is given, so every parity test feeds both sides identical inputs."""

from __future__ import annotations

import numpy as np

from oracle import oracle as orc
from rscm_b200.core import GridType, ModelBuilder, VariableSchema

_OPS = {"Sum": orc.AGG_SUM, "Mean": orc.AGG_MEAN, "Weighted": orc.AGG_WEIGHTED}


def oracle_from_builder(b: ModelBuilder, scenario: dict | None = None) -> orc.OracleModel:
    m = orc.OracleModel()
    for c in b._components:
        m.add_component(c.kind, c.params)
    m.set_time_bounds(b._time_axis.bounds())
    if b._schema is not None:
        for v in b._schema.variables.values():
            m.add_schema_variable(v["name"], v["grid_type"].value)
        for name in b._schema._topological_order():
            a = b._schema.aggregates[name]
            m.add_aggregate(name, _OPS[a["operation"]], a["contributors"], a["weights"], a["grid_type"].value)
    for k, v in b._initial_values.items():
        m.set_initial_value(k, v)
    for ci, var, f in b._unit_factors:
        m.set_unit_factor(ci, var, f)
    for g, w in b._grid_weights.items():
        m.set_grid_weights(g.value, w)
    for name, (ts, _) in b._exogenous._items.items():
        vals = ts.interpolate_into(b._time_axis)._values
        m.set_exogenous(name, vals, {1: 0, 4: 1, 2: 2}[vals.shape[1]])
    if scenario:
        for name, vals in scenario.items():
            v = np.asarray(vals, dtype=float)
            m.set_exogenous(name, v, {1: 0, 4: 1, 2: 2}[1 if v.ndim == 1 else v.shape[1]])
    return m.build()


def oracle_bindings(b: ModelBuilder, bindings: dict) -> list:
    """{'col': 'Type.field' | 'initial:Var' | [..]} -> per column list of (component idx, param idx) / (-1, var)."""
    out = []
    for _, target in bindings.items():
        targets = [target] if isinstance(target, str) else list(target)
        col = []
        for t in targets:
            if t.startswith("initial:"):
                col.append((-1, t[8:]))
                continue
            typ, field = t.rsplit(".", 1)
            want = None
            if "#" in typ:
                typ, idx = typ.split("#")
                want = int(idx)
            for ci, c in enumerate(b._components):
                if c.type_name == typ and (want is None or want == ci):
                    col.append((ci, c.param_names.index(field)))
                    break
            else:
                raise KeyError(t)
        out.append(col)
    return out


def rel_err(actual: np.ndarray, expected: np.ndarray) -> float:
    """max |a - e| / max(|e|) over one output series block; NaN positions must match exactly."""
    a, e = np.asarray(actual), np.asarray(expected)
    assert a.shape == e.shape, (a.shape, e.shape)
    assert np.array_equal(np.isnan(a), np.isnan(e)), "NaN positions differ"
    ok = ~np.isnan(e)
    if not ok.any():
        return 0.0
    scale = max(np.max(np.abs(e[ok])), 1e-300)
    return float(np.max(np.abs(a[ok] - e[ok])) / scale)


def elementwise_rel_err(actual, expected, floor: float) -> float:
    """max |a - e| / max(|e|, floor) element by element."""
    a, e = np.asarray(actual), np.asarray(expected)
    assert np.array_equal(np.isnan(a), np.isnan(e)), "NaN positions differ"
    ok = ~np.isnan(e)
    return float(np.max(np.abs(a[ok] - e[ok]) / np.maximum(np.abs(e[ok]), floor))) if ok.any() else 0.0
