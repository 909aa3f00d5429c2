"""Host-side mirror of ``rscm._lib.core`` for the ensemble hot path.

Same class and method names as the reference's pyo3 surface
(``python/rscm/_lib/core/__init__.pyi``): ``TimeAxis``, ``Timeseries``,
``TimeseriesCollection``, ``VariableSchema``, ``ModelBuilder``, ``Model``.  What
differs is underneath: ``ModelBuilder.build()`` compiles the component graph
ONCE into a fused sm_100a kernel (through the C ABI, ``_ffi``) and
``ModelBuilder.build_ensemble()`` exposes the many-members entry point the
reference reaches by looping ``Model.run`` (``ModelRunner::run_batch``,
``crates/rscm-calibrate/src/model_runner.rs:261-266``).

Scenario ingestion (``Timeseries.interpolate_into``) restates
``crates/rscm-core/src/timeseries.rs:586-611`` and
``crates/rscm-core/src/interpolate/strategies/*.rs`` in numpy: it runs once per
scenario before staging to the GPU, not inside the time loop.
"""

from __future__ import annotations

import ctypes as C
import math
from enum import Enum, auto
from typing import Iterable, Sequence

import numpy as np

from . import _ffi

__all__ = [
    "TimeAxis",
    "InterpolationStrategy",
    "Timeseries",
    "FourBoxTimeseries",
    "HemisphericTimeseries",
    "VariableType",
    "TimeseriesCollection",
    "RequirementType",
    "GridType",
    "VariableSchema",
    "Component",
    "ModelBuilder",
    "Model",
    "Ensemble",
]


# ----------------------------------------------------------------------------
# time axis — crates/rscm-core/src/timeseries.rs:24-211
# ----------------------------------------------------------------------------
class TimeAxis:
    def __init__(self, bounds: np.ndarray):
        b = np.ascontiguousarray(bounds, dtype=np.float64)
        if b.ndim != 1 or b.size < 2:
            raise ValueError("TimeAxis needs at least two bounds")
        if not np.all(np.diff(b) > 0):
            raise ValueError("TimeAxis bounds must be strictly increasing")
        self._bounds = b

    @staticmethod
    def from_values(values) -> "TimeAxis":
        # timeseries.rs:66-79: one extra bound = last + (last - previous)
        v = np.asarray(values, dtype=np.float64)
        if v.size < 2:
            raise ValueError("from_values needs at least two values")
        step = v[-1] - v[-2]
        return TimeAxis(np.concatenate([v, [v[-1] + step]]))

    @staticmethod
    def from_bounds(bounds) -> "TimeAxis":
        return TimeAxis(np.asarray(bounds, dtype=np.float64))

    def values(self) -> np.ndarray:
        return self._bounds[:-1].copy()

    def bounds(self) -> np.ndarray:
        return self._bounds.copy()

    def __len__(self) -> int:
        return self._bounds.size - 1

    def at(self, index: int) -> float:
        if not 0 <= index < len(self):
            raise IndexError(index)
        return float(self._bounds[index])

    def at_bounds(self, index: int) -> tuple[float, float]:
        if not 0 <= index < len(self):
            raise IndexError(index)
        return float(self._bounds[index]), float(self._bounds[index + 1])

    def __eq__(self, other) -> bool:
        return isinstance(other, TimeAxis) and np.array_equal(self._bounds, other._bounds)

    def __repr__(self) -> str:
        return f"TimeAxis({self._bounds[0]}..{self._bounds[-2]}, n={len(self)})"


class InterpolationStrategy(Enum):
    Linear = auto()
    Next = auto()
    Previous = auto()


def _is_close(a: float, b: float) -> bool:
    # is_close! with the `is_close` crate defaults (third-party, not vendored in the
    # reference tree): relative tolerance 1e-8, absolute tolerance 0.
    return abs(a - b) <= 1e-8 * max(abs(a), abs(b))


def _find_segment(target: float, tb: np.ndarray, extrapolate: bool):
    # interpolate/strategies/mod.rs:24-68 ; returns (option, end_segment_idx)
    idx = int(np.searchsorted(tb, target, side="left"))  # binary_search: Ok(i) | Err(insertion)
    fwd = idx == tb.size
    bwd = (not fwd) and idx == 0
    if not fwd and _is_close(float(tb[idx]), target):
        return "boundary", idx
    if (fwd or bwd) and not extrapolate:
        raise RuntimeError(f"extrapolation not allowed for target {target}")
    if bwd:
        return "backward", 0
    if fwd:
        return "forward", tb.size
    return "in", idx


def _interp(strategy: InterpolationStrategy, time: np.ndarray, y: np.ndarray, target: float, extrapolate=True) -> float:
    n = y.size
    if strategy is InterpolationStrategy.Linear:
        # linear_spline.rs:33-95 (the last time value is trimmed before the search)
        opt, e = _find_segment(target, time[:-1], extrapolate)
        e = min(e, n - 1)
        if opt == "boundary":
            return float(y[e])
        if opt == "backward":
            t1, y1, t2, y2 = time[0], y[0], time[1], y[1]
        elif opt == "forward":
            t1, y1, t2, y2 = time[n - 2], y[n - 2], time[n - 1], y[n - 1]
        else:
            t1, y1, t2, y2 = time[e - 1], y[e - 1], time[e], y[e]
        m = (y2 - y1) / (t2 - t1)
        return float(m * (target - t1) + y1)
    if strategy is InterpolationStrategy.Previous:
        # previous.rs
        opt, e = _find_segment(target, time, extrapolate)
        if opt == "boundary":
            return float(y[e])
        if opt == "backward":
            return float(y[0])
        if opt == "forward":
            return float(y[n - 1])
        return float(y[e - 1])
    # next.rs
    opt, e = _find_segment(target, time, extrapolate)
    e = min(e, n - 1)
    if opt == "backward":
        return float(y[0])
    if opt == "forward":
        return float(y[n - 1])
    return float(y[e])


_STRATEGY_ABI = {InterpolationStrategy.Linear: 0, InterpolationStrategy.Next: 1, InterpolationStrategy.Previous: 2}


def interpolate_device(src_times, src_values, dst_times, strategy: InterpolationStrategy = InterpolationStrategy.Linear, out=None,
                       stream: int = 0):
    """``Timeseries::interpolate_into`` for a batch of series on the GPU (``rscm_b200_interpolate_device``).

    ``src_times`` [K] and ``dst_times`` [T] are axis VALUES (timeseries.rs:586-611); ``src_values`` is a CUDA tensor
    [..., K] or [..., K, R] (R = 2 or 4 for grid series).  Returns a CUDA tensor [..., T] / [..., T, R]; results equal the
    host ``interpolate_into`` bit for bit.  Host arrays are uploaded; the values themselves never visit the host."""
    import torch

    def dev(x):
        return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    st, dt, sv = dev(src_times), dev(dst_times), dev(src_values).contiguous()
    K, T = st.numel(), dt.numel()
    if sv.shape[-1] == K:
        R, lead = 1, tuple(sv.shape[:-1])
    elif sv.dim() >= 2 and sv.shape[-2] == K and sv.shape[-1] in (2, 4):
        R, lead = sv.shape[-1], tuple(sv.shape[:-2])
    else:
        raise ValueError(f"src_values {tuple(sv.shape)} does not end in [K={K}] or [K, R]")
    n_series = int(np.prod(lead)) if lead else 1
    shape = lead + ((T,) if R == 1 else (T, R))
    if out is None:
        out = torch.empty(shape, dtype=torch.float64, device=sv.device)
    _ffi.check(_ffi.lib.rscm_b200_interpolate_device(st.data_ptr(), K, sv.data_ptr(), n_series, R, dt.data_ptr(), T, _STRATEGY_ABI[strategy],
                                                     out.data_ptr(), stream))
    return out


class _GridTimeseries:
    """GridTimeseries<T, G> — values [T][R], NaN-initialised, `latest` tracking
    (crates/rscm-core/src/timeseries.rs:261-275, 334-345, 387-397)."""

    _regions = 1

    def __init__(self, values, time_axis: TimeAxis, units: str, interpolation_strategy=InterpolationStrategy.Linear):
        v = np.array(values, dtype=np.float64)
        if v.ndim == 1:
            v = v.reshape(-1, 1)
        if v.shape != (len(time_axis), self._regions):
            raise ValueError(f"values shape {v.shape} does not match ({len(time_axis)}, {self._regions})")
        self._values = v
        self._time_axis = time_axis
        self._units = units
        self._strategy = interpolation_strategy
        self._latest = self._compute_latest()

    def _compute_latest(self) -> int:
        ok = np.where(~np.isnan(self._values).any(axis=1))[0]
        return int(ok[-1]) if ok.size else 0

    def with_interpolation_strategy(self, interpolation_strategy):
        out = type(self).__new__(type(self))
        out.__dict__.update(self.__dict__)
        out._values = self._values.copy()
        out._strategy = interpolation_strategy
        return out

    def __len__(self) -> int:
        return self._values.shape[0]

    @property
    def latest(self) -> int:
        return self._latest

    @property
    def units(self) -> str:
        return self._units

    @property
    def time_axis(self) -> TimeAxis:
        return self._time_axis

    def interpolate_into(self, new_time_axis: TimeAxis):
        # timeseries.rs:586-611: Interp1d over time_axis.values() (NOT bounds)
        t_old = self._time_axis.values()
        new = np.empty((len(new_time_axis), self._regions))
        if np.array_equal(t_old, new_time_axis.values()):
            new[:] = self._values
        else:
            for r in range(self._regions):
                y = self._values[:, r]
                for i, t in enumerate(new_time_axis.values()):
                    new[i, r] = _interp(self._strategy, t_old, y, float(t))
        return type(self)(new if self._regions > 1 else new[:, 0], new_time_axis, self._units, self._strategy)


class Timeseries(_GridTimeseries):
    """Scalar timeseries (``Timeseries`` in the reference stubs)."""

    _regions = 1

    @staticmethod
    def from_values(values, time) -> "Timeseries":
        ta = time if isinstance(time, TimeAxis) else TimeAxis.from_values(time)
        return Timeseries(values, ta, "", InterpolationStrategy.Linear)

    def set(self, time_index: int, value: float) -> None:
        self._values[time_index, 0] = value
        self._latest = self._compute_latest()

    def values(self) -> np.ndarray:
        return self._values[:, 0].copy()

    def latest_value(self):
        v = self._values[self._latest, 0]
        return None if np.isnan(v) else float(v)

    def at(self, time_index: int) -> float:
        return float(self._values[time_index, 0])

    def at_time(self, time: float) -> float:
        return _interp(self._strategy, self._time_axis.values(), self._values[:, 0], float(time))


class FourBoxTimeseries(_GridTimeseries):
    _regions = 4

    def values(self) -> np.ndarray:
        return self._values.copy()


class HemisphericTimeseries(_GridTimeseries):
    _regions = 2

    def values(self) -> np.ndarray:
        return self._values.copy()


class VariableType(Enum):
    Exogenous = auto()
    Endogenous = auto()


class RequirementType(Enum):
    Input = auto()
    Output = auto()
    State = auto()
    EmptyLink = auto()


class GridType(Enum):
    Scalar = 0
    FourBox = 1
    Hemispheric = 2


_GRID_CLS = {GridType.Scalar: Timeseries, GridType.FourBox: FourBoxTimeseries, GridType.Hemispheric: HemisphericTimeseries}


def _grid_of(ts) -> GridType:
    return {1: GridType.Scalar, 4: GridType.FourBox, 2: GridType.Hemispheric}[ts._regions]


class TimeseriesCollection:
    """crates/rscm-core/src/timeseries_collection.rs:318-438"""

    def __init__(self) -> None:
        self._items: dict[str, tuple[object, VariableType]] = {}

    def add_timeseries(self, name: str, timeseries, variable_type: VariableType = VariableType.Exogenous) -> None:
        if name in self._items:
            raise ValueError(f"timeseries {name!r} already exists")
        self._items[name] = (timeseries, variable_type)

    def _get(self, name, cls):
        it = self._items.get(name)
        if it is None or not isinstance(it[0], cls):
            return None
        ts = it[0]
        return ts.with_interpolation_strategy(ts._strategy)  # clone

    def get_timeseries_by_name(self, name: str):
        return self._get(name, Timeseries)

    def get_fourbox_timeseries_by_name(self, name: str):
        return self._get(name, FourBoxTimeseries)

    def get_hemispheric_timeseries_by_name(self, name: str):
        return self._get(name, HemisphericTimeseries)

    def names(self) -> list[str]:
        return list(self._items)

    def timeseries(self) -> list[Timeseries]:
        return [self._get(n, Timeseries) for n, (t, _) in self._items.items() if isinstance(t, Timeseries)]

    def variable_type(self, name: str) -> VariableType:
        return self._items[name][1]


class VariableSchema:
    """crates/rscm-core/src/schema.rs — variables + aggregates."""

    _OPS = {"Sum": _ffi.AGG_SUM, "Mean": _ffi.AGG_MEAN, "Weighted": _ffi.AGG_WEIGHTED}

    def __init__(self) -> None:
        self.variables: dict[str, dict] = {}
        self.aggregates: dict[str, dict] = {}

    def add_variable(self, name: str, unit: str, grid_type: GridType | None = None) -> None:
        self.variables[name] = {"name": name, "unit": unit, "grid_type": grid_type or GridType.Scalar}

    def add_aggregate(self, name, unit, operation, contributors, weights=None, grid_type=None) -> None:
        if operation not in self._OPS:
            raise ValueError(f"operation must be one of Sum, Mean, Weighted; got {operation!r}")
        if operation == "Weighted" and weights is None:
            raise ValueError("Weighted aggregate requires weights")
        self.aggregates[name] = {
            "name": name,
            "unit": unit,
            "operation": operation,
            "contributors": list(contributors),
            "weights": None if weights is None else [float(w) for w in weights],
            "grid_type": grid_type or GridType.Scalar,
        }

    def contains(self, name: str) -> bool:
        return name in self.variables or name in self.aggregates

    def validate(self) -> None:
        # schema.rs validate(): contributors exist, weights match, no cycles
        for a in self.aggregates.values():
            for c in a["contributors"]:
                if not self.contains(c):
                    raise ValueError(f"aggregate {a['name']!r}: unknown contributor {c!r}")
                g = (self.variables.get(c) or self.aggregates.get(c))["grid_type"]
                if g != a["grid_type"]:
                    raise ValueError(f"aggregate {a['name']!r}: grid type mismatch with contributor {c!r}")
            if a["operation"] == "Weighted" and len(a["weights"]) != len(a["contributors"]):
                raise ValueError(f"aggregate {a['name']!r}: weight count does not match contributors")
        self._topological_order()

    def _topological_order(self) -> list[str]:
        # schema.topological_order_aggregates: contributors that are aggregates first
        order: list[str] = []
        state: dict[str, int] = {}

        def visit(n: str) -> None:
            if state.get(n) == 2:
                return
            if state.get(n) == 1:
                raise ValueError(f"circular dependency between aggregates at {n!r}")
            state[n] = 1
            for c in self.aggregates[n]["contributors"]:
                if c in self.aggregates:
                    visit(c)
            state[n] = 2
            order.append(n)

        for n in self.aggregates:
            visit(n)
        return order


class Component:
    """A component instance produced by ``<Kind>Builder.from_parameters(...).build()``:
    a component kind plus its parameter block in the C ABI's documented order."""

    def __init__(self, kind: int, type_name: str, param_names: Sequence[str], params: Sequence[float]):
        self.kind = kind
        self.type_name = type_name
        self.param_names = list(param_names)
        self.params = [float(p) for p in params]

    def __repr__(self) -> str:
        body = ", ".join(f"{n}: {v}" for n, v in zip(self.param_names, self.params))
        return f"{self.type_name} {{ {body} }}"


# ----------------------------------------------------------------------------
# ModelBuilder / Model / Ensemble
# ----------------------------------------------------------------------------
class ModelBuilder:
    """crates/rscm-core/src/model/builder.rs:29-204 (Python names: python/model.rs:24-151)."""

    def __init__(self) -> None:
        self._components: list[Component] = []
        self._initial_values: dict[str, float] = {}
        self._exogenous = TimeseriesCollection()
        self._time_axis: TimeAxis | None = None
        self._schema: VariableSchema | None = None
        self._grid_weights: dict[GridType, list[float]] = {}
        self._unit_factors: list[tuple[int, str, float]] = []

    def with_time_axis(self, time_axis: TimeAxis) -> "ModelBuilder":
        self._time_axis = time_axis
        return self

    def with_rust_component(self, component: Component) -> "ModelBuilder":
        self._components.append(component)
        return self

    with_component = with_rust_component

    def with_py_component(self, component) -> "ModelBuilder":
        raise NotImplementedError(
            "user-defined Python components cannot run inside the fused GPU kernel; "
            "the engine has no CPU fallback (out of scope: SURVEY.md §8)"
        )

    def with_initial_values(self, initial_values: dict) -> "ModelBuilder":
        for k, v in initial_values.items():
            self._initial_values[k] = float(v)
        return self

    def with_exogenous_variable(self, name: str, timeseries) -> "ModelBuilder":
        self._exogenous.add_timeseries(name, timeseries, VariableType.Exogenous)
        return self

    def with_exogenous_collection(self, timeseries: TimeseriesCollection) -> "ModelBuilder":
        for n, (t, _) in timeseries._items.items():
            self._exogenous.add_timeseries(n, t, VariableType.Exogenous)
        return self

    def with_schema(self, schema: VariableSchema) -> "ModelBuilder":
        self._schema = schema
        return self

    def with_grid_weights(self, grid_type: GridType, weights) -> "ModelBuilder":
        w = [float(x) for x in weights]
        need = {GridType.FourBox: 4, GridType.Hemispheric: 2}.get(grid_type)
        if need is None:
            raise ValueError("grid weights only apply to FourBox / Hemispheric")
        if len(w) != need:
            raise ValueError(f"expected {need} weights")
        if abs(sum(w) - 1.0) > 1e-6:
            raise ValueError("weights must sum to 1.0")
        self._grid_weights[grid_type] = w
        return self

    def with_unit_factor(self, component_index: int, variable: str, factor: float) -> "ModelBuilder":
        """Pre-computed unit conversion factor for one (variable, consuming component)
        pair (model/runtime.rs:385-389); the reference derives it from its unit registry,
        which is outside the hot path."""
        self._unit_factors.append((int(component_index), variable, float(factor)))
        return self

    # -- lowering to the C ABI ------------------------------------------------
    def _create_handle(self, dtype: str = "f64", device: int = -1):
        if self._time_axis is None:
            raise ValueError("time axis required")
        keep = []  # keep ctypes buffers alive across the call
        comps = (_ffi.ComponentDesc * max(1, len(self._components)))()
        for i, c in enumerate(self._components):
            arr = (C.c_double * len(c.params))(*c.params)
            keep.append(arr)
            comps[i] = _ffi.ComponentDesc(c.kind, len(c.params), arr)
        d = _ffi.GraphDesc()
        d.abi_version = _ffi.ABI_VERSION
        d.n_components = len(self._components)
        d.components = comps
        if self._schema is not None:
            self._schema.validate()
            d.has_schema = 1
            svars = list(self._schema.variables.values())
            sv = (_ffi.SchemaVariable * max(1, len(svars)))()
            for i, v in enumerate(svars):
                sv[i] = _ffi.SchemaVariable(v["name"].encode(), v["grid_type"].value)
            d.n_schema_variables = len(svars)
            d.schema_variables = sv
            order = self._schema._topological_order()
            ag = (_ffi.AggregateDesc * max(1, len(order)))()
            for i, name in enumerate(order):
                a = self._schema.aggregates[name]
                names = (C.c_char_p * len(a["contributors"]))(*[c.encode() for c in a["contributors"]])
                keep.append(names)
                w = None
                if a["weights"] is not None:
                    w = (C.c_double * len(a["weights"]))(*a["weights"])
                    keep.append(w)
                ag[i] = _ffi.AggregateDesc(
                    name.encode(), VariableSchema._OPS[a["operation"]], a["grid_type"].value, len(a["contributors"]),
                    names, w if w is not None else C.POINTER(C.c_double)(),
                )
            d.n_aggregates = len(order)
            d.aggregates = ag
            keep += [sv, ag]
        iv = (_ffi.InitialValue * max(1, len(self._initial_values)))()
        for i, (k, v) in enumerate(self._initial_values.items()):
            iv[i] = _ffi.InitialValue(k.encode(), v)
        d.n_initial_values = len(self._initial_values)
        d.initial_values = iv
        uf = (_ffi.UnitFactor * max(1, len(self._unit_factors)))()
        for i, (ci, var, f) in enumerate(self._unit_factors):
            uf[i] = _ffi.UnitFactor(ci, var.encode(), f)
        d.n_unit_factors = len(self._unit_factors)
        d.unit_factors = uf
        if GridType.FourBox in self._grid_weights:
            w4 = (C.c_double * 4)(*self._grid_weights[GridType.FourBox])
            keep.append(w4)
            d.four_box_weights = w4
        if GridType.Hemispheric in self._grid_weights:
            w2 = (C.c_double * 2)(*self._grid_weights[GridType.Hemispheric])
            keep.append(w2)
            d.hemispheric_weights = w2
        b = self._time_axis.bounds()
        tb = (C.c_double * b.size)(*b.tolist())
        d.n_times = len(self._time_axis)
        d.time_bounds = tb
        d.compute_dtype = {"f64": 0, "f32": 1}[dtype]
        d.device = device
        h = C.c_void_p()
        _ffi.check(_ffi.lib.rscm_b200_ensemble_create(C.byref(d), C.byref(h)))
        return h

    def build_ensemble(self, dtype: str = "f64", device: int = -1) -> "Ensemble":
        """Compile the component graph once for many members (the engine's replacement
        for rebuilding a Model per member, model_runner.rs:233-235)."""
        return Ensemble(self, dtype=dtype, device=device)

    def build(self) -> "Model":
        return Model(self)


class Ensemble:
    """The compiled component graph on one GPU: `run` = ``ModelRunner::run_batch``."""

    def __init__(self, builder: ModelBuilder, dtype: str = "f64", device: int = -1):
        self._builder = builder
        self._time_axis = builder._time_axis
        self._h = builder._create_handle(dtype=dtype, device=device)
        self.dtype = dtype
        L = _ffi.lib
        self.variable_names = [L.rscm_b200_variable_name(self._h, i).decode() for i in range(L.rscm_b200_n_variables(self._h))]
        self.variable_grids = [GridType(L.rscm_b200_variable_grid(self._h, i)) for i in range(len(self.variable_names))]
        self.exogenous_names = [
            self.variable_names[L.rscm_b200_exogenous_variable(self._h, i)] for i in range(L.rscm_b200_n_exogenous(self._h))
        ]
        self.param_names: list[str] = []
        self._selected = list(range(len(self.variable_names)))
        self._tsel = (0, len(self._time_axis), 1)

    def close(self) -> None:
        if getattr(self, "_h", None):
            _ffi.lib.rscm_b200_ensemble_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection ---------------------------------------------------------
    @property
    def n_times(self) -> int:
        return len(self._time_axis)

    def regions(self, name: str) -> int:
        return {GridType.Scalar: 1, GridType.FourBox: 4, GridType.Hemispheric: 2}[self.variable_grids[self.variable_names.index(name)]]

    def is_endogenous(self, name: str) -> bool:
        return bool(_ffi.lib.rscm_b200_variable_is_endogenous(self._h, self.variable_names.index(name)))

    def execution_order(self) -> list[int]:
        n = _ffi.lib.rscm_b200_n_nodes(self._h)
        buf = (C.c_int * max(1, n))()
        k = _ffi.lib.rscm_b200_execution_order(self._h, buf, n)
        return [buf[i] for i in range(k)]

    def variable_source(self, component: int, variable: str) -> int:
        return _ffi.lib.rscm_b200_variable_source(self._h, component, variable.encode())

    def program_signature(self) -> str:
        return _ffi.lib.rscm_b200_program_signature(self._h).decode()

    def program_is_jit(self) -> bool:
        return bool(_ffi.lib.rscm_b200_program_is_jit(self._h))

    def time_index(self, time: float) -> int:
        return _ffi.lib.rscm_b200_time_index(self._h, float(time))

    def launch_count(self) -> int:
        return int(_ffi.lib.rscm_b200_launch_count(self._h))

    def shared_bytes(self, log_posterior: bool = False) -> int:
        """Dynamic shared memory per CTA of the fused kernel."""
        return int(_ffi.lib.rscm_b200_shared_bytes(self._h, 1 if log_posterior else 0))

    def kernel_ms(self, reset: bool = False) -> float:
        return float(_ffi.lib.rscm_b200_kernel_ms(self._h, 1 if reset else 0))

    # -- configuration -----------------------------------------------------------
    def bind_parameters(self, bindings: dict[str, str | Sequence[str]] | Sequence[str]) -> "Ensemble":
        """``bindings`` maps each parameter column name to the slot(s) it feeds, e.g.
        ``{"lambda0": "TwoLayer.lambda0", "conc_pi": ["CarbonCycle.conc_pi", "CO2ERF.conc_pi"],
        "T0": "initial:Surface Temperature"}``; a plain list of slot names binds column i
        to slot i.  Column order = the order of the mapping."""
        if not isinstance(bindings, dict):
            bindings = {s: s for s in bindings}
        slots, cols = [], []
        for j, (_, target) in enumerate(bindings.items()):
            for s in [target] if isinstance(target, str) else list(target):
                slots.append(s.encode())
                cols.append(j)
        arr = (C.c_char_p * max(1, len(slots)))(*slots)
        ca = (C.c_int32 * max(1, len(cols)))(*cols)
        _ffi.check(_ffi.lib.rscm_b200_bind_parameters(self._h, len(slots), arr, ca, len(bindings)), self._h)
        self.param_names = list(bindings)
        return self

    def select_outputs(self, variables: Iterable[str] | None = None, t_start: int = 0, t_stop: int | None = None, t_step: int = 1) -> "Ensemble":
        names = list(self.variable_names if variables is None else variables)
        idx = [self.variable_names.index(n) for n in names]
        t_stop = self.n_times if t_stop is None else t_stop
        arr = (C.c_int32 * max(1, len(idx)))(*idx)
        _ffi.check(_ffi.lib.rscm_b200_select_outputs(self._h, len(idx), arr, t_start, t_stop, t_step), self._h)
        self._selected = idx
        self._tsel = (t_start, t_stop, t_step)
        return self

    @property
    def output_rows(self) -> int:
        return int(_ffi.lib.rscm_b200_output_rows(self._h))

    def selected_times(self) -> np.ndarray:
        return self._time_axis.values()[self._tsel[0]:self._tsel[1]:self._tsel[2]]

    def output_layout(self) -> dict[str, tuple[int, int, int]]:
        """name -> (first row, n_selected_times, n_regions); rows are [time][region]."""
        nt = len(range(*self._tsel))
        out, row = {}, 0
        for v in self._selected:
            r = self.regions(self.variable_names[v])
            out[self.variable_names[v]] = (row, nt, r)
            row += nt * r
        return out

    def split_outputs(self, out: np.ndarray) -> dict[str, np.ndarray]:
        """[rows][runs] -> {name: [T_sel, R, runs] (R squeezed for scalars)} views."""
        res = {}
        for name, (row, nt, r) in self.output_layout().items():
            blk = out[row:row + nt * r].reshape(nt, r, out.shape[1])
            res[name] = blk[:, 0, :] if r == 1 else blk
        return res

    # -- scenarios -----------------------------------------------------------------
    def scenario_shape(self) -> tuple[int, ...]:
        return (sum(self.regions(n) * self.n_times for n in self.exogenous_names),)

    def pack_scenarios(self, scenarios: Sequence[dict[str, np.ndarray]]) -> np.ndarray:
        """list of {exogenous name: [T] or [T,R]} -> [S][n_exo][T][R] flat array.
        A missing exogenous variable is an all-NaN series with the builder's initial
        value (if any) at index 0 — what the reference's collection holds
        (model/builder.rs:771-781)."""
        S = len(scenarios)
        width = self.scenario_shape()[0]
        arr = np.empty((max(S, 0), width))
        for s, sc in enumerate(scenarios):
            off = 0
            for n in self.exogenous_names:
                r = self.regions(n)
                if n in sc:
                    v = np.asarray(sc[n], dtype=np.float64).reshape(self.n_times, r)
                else:
                    v = np.full((self.n_times, r), np.nan)
                    if n in self._builder._initial_values:
                        v[0, :] = self._builder._initial_values[n]
                arr[s, off:off + v.size] = v.ravel()
                off += v.size
        return arr

    def default_scenarios(self) -> np.ndarray:
        """The builder's own exogenous timeseries, interpolated onto the model axis
        (model/builder.rs:768, timeseries.rs:586-611), as one scenario."""
        sc = {}
        for n in self.exogenous_names:
            it = self._builder._exogenous._items.get(n)
            if it is not None and _GRID_CLS[self.variable_grids[self.variable_names.index(n)]] is type(it[0]):
                sc[n] = it[0].interpolate_into(self._time_axis)._values
        return self.pack_scenarios([sc])

    # -- execution ---------------------------------------------------------------------
    @staticmethod
    def _ptr(x) -> int:
        if x is None:
            return 0
        if isinstance(x, np.ndarray):
            return x.ctypes.data
        return int(x.data_ptr())  # torch tensor

    def run_device(self, params, scenarios, out, status=None, *, M: int | None = None, S: int | None = None, layout: int = 0, stream: int = 0) -> None:
        """All arguments are CUDA tensors (torch) or raw device pointers; asynchronous."""
        if M is None:
            M = params.shape[1] if layout == 0 else params.shape[0]
        if S is None:
            S = 0 if scenarios is None else scenarios.shape[0]
        _ffi.check(
            _ffi.lib.rscm_b200_run_device(self._h, self._ptr(params), M, layout, self._ptr(scenarios), S, self._ptr(out), self._ptr(status), stream),
            self._h,
        )

    # -- summaries across members ------------------------------------
    def member_quantiles_device(self, out, q: Sequence[float], result, *, M: int, S: int = 1, stream: int = 0) -> None:
        """Quantiles across members of a device output block ``out`` [rows][S*M] into ``result`` [len(q)][rows][S]
        (CUDA tensors or raw pointers; asynchronous; at most 5 quantiles per call)."""
        qs = (C.c_double * len(q))(*[float(x) for x in q])
        _ffi.check(_ffi.lib.rscm_b200_member_quantiles(self._ptr(out), self.output_rows, max(S, 1), M, qs, len(q), self._ptr(result), stream))

    def run_quantiles(self, params: np.ndarray, scenarios: np.ndarray | None, q: Sequence[float], *, layout: int = 1) -> dict[str, np.ndarray]:
        """Run the ensemble and return, instead of every member's series, their across-member quantiles per scenario:
        ``{variable: [len(q), T_sel, R, S]}`` (R squeezed for scalars).  The member outputs never leave the GPU — what a
        notebook of the reference gets from ``np.nanquantile`` over looped ``Model.run()`` results, without the loop or the copy.
        Needs the whole output block in device memory at once."""
        import torch

        p = np.ascontiguousarray(params, dtype=np.float64)
        M = p.shape[0] if layout == 1 else p.shape[1]
        if scenarios is None and self.exogenous_names:
            scenarios = self.default_scenarios()
        S = 1 if scenarios is None else scenarios.shape[0]
        rows = self.output_rows
        d_p = torch.from_numpy(np.ascontiguousarray(p.T) if layout == 1 else p).cuda()
        d_s = None if scenarios is None else torch.from_numpy(np.ascontiguousarray(scenarios, dtype=np.float64)).cuda()
        d_o = torch.empty((rows, S * M), dtype=torch.float64, device="cuda")
        self.run_device(d_p, d_s, d_o, layout=0, M=M, S=0 if d_s is None else S)
        res = torch.empty((len(q), rows, S), dtype=torch.float64, device="cuda")
        for k0 in range(0, len(q), 5):
            self.member_quantiles_device(d_o, q[k0:k0 + 5], res[k0:k0 + 5], M=M, S=S)
        host = res.cpu().numpy()
        out = {}
        for name, (row, nt, r) in self.output_layout().items():
            blk = host[:, row:row + nt * r, :].reshape(len(q), nt, r, S)
            out[name] = blk[:, :, 0, :] if r == 1 else blk
        return out

    def run(self, params: np.ndarray | None, scenarios: np.ndarray | None = None, *, layout: int = 1, out: np.ndarray | None = None,
            status: np.ndarray | None = None) -> np.ndarray:
        """Host entry point: ``params`` [M, n_cols] (layout 1, one row per member like the
        reference's ``&[Vec<f64>]``) or [n_cols, M] (layout 0); ``scenarios`` [S, ...] from
        :meth:`pack_scenarios`.  Returns out [rows, S*M] (host)."""
        if scenarios is None:
            scenarios = self.default_scenarios() if self.exogenous_names else None
        S = 0 if scenarios is None else scenarios.shape[0]
        if params is None:
            M = 1
            p = None
        else:
            p = np.ascontiguousarray(params, dtype=np.float64)
            if p.ndim == 1:
                p = p.reshape(1, -1) if layout == 1 else p.reshape(-1, 1)
            M = p.shape[0] if layout == 1 else p.shape[1]
            if (p.shape[1] if layout == 1 else p.shape[0]) != len(self.param_names):
                raise ValueError(f"expected {len(self.param_names)} parameter columns")
        sc = None if scenarios is None else np.ascontiguousarray(scenarios, dtype=np.float64)
        runs = max(S, 1) * M
        if out is None:
            out = np.empty((self.output_rows, runs))
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous and out.flags.writeable
                  and out.shape == (self.output_rows, runs)):
            raise ValueError(f"out must be a writable C-contiguous float64 array of shape ({self.output_rows}, {runs})")
        if status is not None and not (isinstance(status, np.ndarray) and status.dtype == np.uint8 and status.flags.c_contiguous
                                       and status.flags.writeable and status.size == runs):
            raise ValueError(f"status must be a writable C-contiguous uint8 array of {runs} elements")
        _ffi.check(
            _ffi.lib.rscm_b200_run_host(self._h, self._ptr(p), M, layout, self._ptr(sc), S, self._ptr(out), self._ptr(status)),
            self._h,
        )
        return out

    # -- calibration ---------------------------------------------------------------------
    def set_target(self, observations: Sequence[tuple[str, float, float, float]], normalize: bool = False) -> "Ensemble":
        """observations: (variable name, time, value, sigma) grouped by variable in Target order."""
        obs = (_ffi.Obs * max(1, len(observations)))()
        for i, (name, time, value, sigma) in enumerate(observations):
            obs[i] = _ffi.Obs(self.variable_names.index(name), self.time_index(time), value, sigma)
        _ffi.check(_ffi.lib.rscm_b200_set_target(self._h, obs, len(observations), 1 if normalize else 0), self._h)
        return self

    def set_priors(self, priors: Sequence[tuple]) -> "Ensemble":
        """priors: one (kind, a, b[, low, high]) per parameter column (kind = _ffi.PRIOR_*)."""
        arr = (_ffi.Prior * max(1, len(priors)))()
        for i, p in enumerate(priors):
            p = tuple(p) + (0.0,) * (5 - len(p))
            arr[i] = _ffi.Prior(int(p[0]), 0, float(p[1]), float(p[2]), float(p[3]), float(p[4]))
        _ffi.check(_ffi.lib.rscm_b200_set_priors(self._h, arr, len(priors)), self._h)
        return self

    def log_posterior_device(self, params, scenarios, logpost, summary=None, *, M=None, S=None, layout: int = 0, stream: int = 0) -> None:
        if M is None:
            M = params.shape[1] if layout == 0 else params.shape[0]
        if S is None:
            S = 0 if scenarios is None else scenarios.shape[0]
        _ffi.check(
            _ffi.lib.rscm_b200_logpost_device(self._h, self._ptr(params), M, layout, self._ptr(scenarios), S, self._ptr(logpost), self._ptr(summary), stream),
            self._h,
        )

    def log_posterior(self, params: np.ndarray, scenarios: np.ndarray | None = None, *, layout: int = 1, with_summary: bool = False):
        if scenarios is None:
            scenarios = self.default_scenarios() if self.exogenous_names else None
        S = 0 if scenarios is None else scenarios.shape[0]
        p = np.ascontiguousarray(params, dtype=np.float64)
        M = p.shape[0] if layout == 1 else p.shape[1]
        sc = None if scenarios is None else np.ascontiguousarray(scenarios, dtype=np.float64)
        lp = np.empty(max(S, 1) * M)
        summ = _ffi.LogpostSummary() if with_summary else None
        _ffi.check(
            _ffi.lib.rscm_b200_logpost_host(self._h, self._ptr(p), M, layout, self._ptr(sc), S, self._ptr(lp), C.byref(summ) if summ is not None else None),
            self._h,
        )
        if with_summary:
            return lp, {"max_logpost": summ.max_logpost, "argmax": summ.argmax, "sum_finite": summ.sum_finite, "n_finite": summ.n_finite, "n_runs": summ.n_runs}
        return lp


class Model:
    """Single-member view with the reference's ``Model`` methods
    (crates/rscm-core/src/model/runtime.rs:353-360,515-554; python/model.rs:162-240).
    The run itself is one member of the fused GPU kernel."""

    def __init__(self, builder: ModelBuilder):
        self._ens = builder.build_ensemble()
        self._builder = builder
        self._time_axis = builder._time_axis
        self._time_index = 0
        self._full: dict[str, np.ndarray] | None = None

    def current_time(self) -> float:
        return self._time_axis.at(self._time_index)

    def current_time_bounds(self) -> tuple[float, float]:
        return self._time_axis.at_bounds(self._time_index)

    def _ensure(self) -> None:
        if self._full is None:
            out = self._ens.run(None)
            self._full = {k: np.array(v[..., 0]) for k, v in self._ens.split_outputs(out).items()}

    def step(self) -> None:
        assert self._time_index < len(self._time_axis) - 1
        self._ensure()
        self._time_index += 1

    def run(self) -> None:
        self._ensure()
        self._time_index = len(self._time_axis) - 1

    def finished(self) -> bool:
        return self._time_index == len(self._time_axis) - 1

    def execution_order(self) -> list[int]:
        return self._ens.execution_order()

    def timeseries(self) -> TimeseriesCollection:
        self._ensure()
        coll = TimeseriesCollection()
        for name, grid in zip(self._ens.variable_names, self._ens.variable_grids):
            vals = self._full[name].copy()
            endo = self._ens.is_endogenous(name)
            if endo:  # values beyond the current step have not been "computed" yet
                vals[self._time_index + 1:] = np.nan
            ts = _GRID_CLS[grid](vals, self._time_axis, "", InterpolationStrategy.Linear)
            coll.add_timeseries(name, ts, VariableType.Endogenous if endo else VariableType.Exogenous)
        return coll
