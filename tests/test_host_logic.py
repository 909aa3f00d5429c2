"""CPU-side tests of the product's host logic: the C-ABI library loads and exports every
symbol the header declares, the C++ graph compiler reproduces ModelBuilder::build, the
scenario-ingestion (interpolation) code follows the reference's strategies, and every run
entry point fails loudly without a CUDA device (no CPU fallback).  No compute calls."""

import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from rscm_b200 import _ffi, calibrate as cal
from rscm_b200 import synthetic as syn
from rscm_b200.components import CarbonCycleBuilder, CO2ERFBuilder
from rscm_b200.core import (GridType, InterpolationStrategy, ModelBuilder, TimeAxis, Timeseries, TimeseriesCollection,
                            VariableSchema, _interp)
from rscm_b200.magicc import GhgForcingBuilder
from rscm_b200.two_layer import TwoLayerBuilder

from .helpers import oracle_from_builder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_only(builder):
    return builder.build_ensemble(device=-2)


# ---- the boundary -------------------------------------------------------------------------------------
def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "rscm_b200.h")).read()
    declared = set(re.findall(r"\b(rscm_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(_ffi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/rscm_b200.h but not exported"
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    assert lib.rscm_b200_abi_version() == 1


def test_no_cpu_fallback_without_device():
    if _ffi.lib.rscm_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_ffi.EngineError) as e:
        syn.two_layer_builder().build_ensemble()
    assert e.value.code == _ffi.ENODEVICE and "no CPU fallback" in str(e.value)


def test_host_only_handle_cannot_run():
    ens = host_only(syn.two_layer_builder())
    with pytest.raises(_ffi.EngineError) as e:
        ens.run(None, ens.pack_scenarios([{"Effective Radiative Forcing": np.zeros(351)}]))
    assert e.value.code == _ffi.ENODEVICE


def test_unknown_component_kind_is_rejected_not_emulated():
    from rscm_b200.core import Component

    b = ModelBuilder().with_time_axis(syn.time_axis()).with_rust_component(Component(99, "Mystery", ["x"], [1.0]))
    with pytest.raises(_ffi.EngineError, match="no device implementation"):
        host_only(b)


def test_graph_outside_the_aot_registry_is_compiled_at_run_time(tmp_path, monkeypatch):
    """A graph shape that is not one of the ahead-of-time programs goes through the emitter + NVRTC
    (sm_100a cubin, no GPU needed to compile); the named graphs do not."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = (ModelBuilder().with_time_axis(syn.time_axis())
         .with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": 3.7, "conc_pi": 278.0}).build())
         .with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
         .with_rust_component(GhgForcingBuilder.from_parameters({}).build())
         .with_initial_values({"Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0}))
    ens = host_only(b)
    assert ens.program_is_jit()
    assert any(f.endswith(".cubin") for f in os.listdir(tmp_path))  # disk cache written
    assert host_only(b).program_is_jit()                            # second build served from the cache
    assert not host_only(syn.coupled_builder()).program_is_jit()


# ---- graph compiler == ModelBuilder::build (checked against the independent oracle implementation) ---------
NAMED = {
    "two_layer": lambda: syn.two_layer_builder(),
    "coupled": lambda: syn.coupled_builder(),
    "carbon_cycle": lambda: (ModelBuilder().with_time_axis(syn.time_axis())
                             .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": 20.3, "conc_pi": 280.0, "alpha_temperature": 0.0}).build())
                             .with_initial_values({"Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": 280.0})),
    "carbon_cycle_co2erf": lambda: (ModelBuilder().with_time_axis(syn.time_axis())
                                    .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": 20.3, "conc_pi": 280.0, "alpha_temperature": 0.0}).build())
                                    .with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": 4.0, "conc_pi": 280.0}).build())
                                    .with_initial_values({"Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": 300.0})),
    "ghg_forcing": lambda: ModelBuilder().with_time_axis(syn.time_axis()).with_rust_component(GhgForcingBuilder.from_parameters({}).build()),
}


@pytest.mark.parametrize("name", sorted(NAMED))
def test_graph_compiler_matches_oracle_builder(name):
    b = NAMED[name]()
    ens = host_only(b)
    m = oracle_from_builder(b)
    assert ens.variable_names == m.names
    assert [g.value for g in ens.variable_grids] == m.grids
    assert ens.execution_order() == m.execution_order()
    assert [ens.is_endogenous(n) for n in ens.variable_names] == [m.is_endogenous(n) for n in m.names]
    for ci, comp in enumerate(b._components):
        for v in ens.variable_names:
            assert ens.variable_source(ci, v) == m.variable_source(ci, v), (ci, v)
    assert "rscm_dev::" in ens.program_signature()


def test_coupled_graph_is_the_reference_chain():
    ens = host_only(syn.coupled_builder())
    assert ens.execution_order() == [0, 1, 3, 2]
    assert ens.exogenous_names == ["Emissions|CO2|Anthropogenic"]
    assert ens.variable_source(0, "Surface Temperature") == _ffi.SRC_EXOGENOUS  # lagged feedback (builder.rs:465-485)
    assert ens.variable_source(1, "Atmospheric Concentration|CO2") == _ffi.SRC_UPSTREAM
    assert ens.variable_source(2, "Effective Radiative Forcing") == _ffi.SRC_UPSTREAM
    assert ens.variable_source(2, "Surface Temperature") == _ffi.SRC_OWN_STATE


def test_cycle_is_rejected():
    schema = VariableSchema()
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", ["Effective Radiative Forcing|CO2"])
    schema.add_variable("Effective Radiative Forcing|CO2", "W/m^2")
    b = (ModelBuilder().with_time_axis(syn.time_axis()).with_schema(schema)
         .with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
         .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": 20.0, "conc_pi": 278.0, "alpha_temperature": 0.0}).build())
         .with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": 3.7, "conc_pi": 278.0}).build())
         .with_initial_values({"Surface Temperature": 0, "Deep Ocean Temperature": 0, "Atmospheric Concentration|CO2": 278,
                               "Cumulative Emissions|CO2": 0, "Cumulative Land Uptake": 0}))
    with pytest.raises(_ffi.EngineError, match="cycle"):
        host_only(b)


def test_missing_initial_value_and_bad_bindings():
    b = ModelBuilder().with_time_axis(syn.time_axis()).with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
    with pytest.raises(_ffi.EngineError, match="initial value"):
        host_only(b)
    ens = host_only(syn.coupled_builder())
    ens.bind_parameters(syn.COUPLED_BINDINGS)
    assert ens.param_names == list(syn.COUPLED_BINDINGS)
    with pytest.raises(_ffi.EngineError, match="cannot vary per member"):
        ens.bind_parameters({"h": "CarbonCycle.step_size"})
    with pytest.raises(_ffi.EngineError, match="no component matches"):
        ens.bind_parameters({"x": "Nope.x"})
    ens.bind_parameters({"conc_pi": ["CarbonCycle.conc_pi", "CO2ERF.conc_pi"], "T0": "initial:Surface Temperature"})


def test_output_selection_rows_and_time_index():
    ens = host_only(syn.coupled_builder())
    assert ens.output_rows == 8 * 351
    ens.select_outputs(["Surface Temperature"], t_start=100, t_stop=351, t_step=10)
    assert ens.output_rows == len(range(100, 351, 10))
    assert ens.time_index(1850.0) == 100 and ens.time_index(1850.0000004) == 100 and ens.time_index(1850.5) == -1
    with pytest.raises(_ffi.EngineError):
        ens.select_outputs(["Surface Temperature"], t_start=5, t_stop=400)


def test_target_validation_mirrors_reference_errors():
    ens = host_only(syn.coupled_builder())
    ens.set_target([("Surface Temperature", 1900.0, 0.3, 0.1)])
    with pytest.raises(_ffi.EngineError, match="time index"):
        ens.set_target([("Surface Temperature", 1900.5, 0.3, 0.1)])
    with pytest.raises(_ffi.EngineError, match="positive"):
        ens.set_target([("Surface Temperature", 1900.0, 0.3, 0.0)])


# ---- scenario ingestion: interpolate strategies (crates/rscm-core/src/interpolate/strategies/*.rs tests) ------
TIME = np.array([0.0, 0.5, 1.0, 1.5])
Y = np.array([5.0, 8.0, 9.0])


@pytest.mark.parametrize("strategy,targets,expected", [
    (InterpolationStrategy.Linear, [0.0, 0.25, 0.5, 0.75, 1.0], [5.0, 6.5, 8.0, 8.5, 9.0]),
    (InterpolationStrategy.Previous, [0.0, 0.25, 0.5, 0.75, 1.0], [5.0, 5.0, 8.0, 8.0, 9.0]),
    (InterpolationStrategy.Next, [0.0, 0.25, 0.5, 0.75, 1.0], [5.0, 8.0, 8.0, 9.0, 9.0]),
])
def test_interpolation_reference_vectors(strategy, targets, expected):
    for t, e in zip(targets, expected):
        assert _interp(strategy, TIME, Y, t, extrapolate=False) == pytest.approx(e, rel=1e-12)


def test_interpolation_extrapolation_rules():
    with pytest.raises(RuntimeError):
        _interp(InterpolationStrategy.Previous, np.array([0.0, 1.0]), np.array([5.0]), -1.0, extrapolate=False)
    with pytest.raises(RuntimeError):
        _interp(InterpolationStrategy.Previous, np.array([0.0, 1.0]), np.array([5.0]), 1.2, extrapolate=False)
    # linear forward extrapolation uses the last two points, backward the first two
    assert _interp(InterpolationStrategy.Linear, TIME, Y, 2.0) == pytest.approx(9.0 + 2.0 * 1.0)
    assert _interp(InterpolationStrategy.Linear, TIME, Y, -0.5) == pytest.approx(5.0 - 3.0)


def test_interpolate_into_uses_time_values_like_the_reference():
    # docs/notebooks/coupled_model.py:397-409: knots -> annual axis
    years = np.array([1750.0, 1850.0, 1950.0, 2000.0, 2020.0, 2050.0, 2100.0])
    vals = np.array([0.0, 0.5, 3.0, 7.0, 10.0, 5.0, 1.0])
    ts = Timeseries(vals, TimeAxis.from_bounds(np.concatenate([years, [2101.0]])), "GtC / yr", InterpolationStrategy.Linear)
    annual = ts.interpolate_into(syn.time_axis()).values()
    np.testing.assert_allclose(annual, np.interp(np.arange(1750.0, 2101.0), years, vals), rtol=1e-13, atol=1e-13)
    assert annual[0] == 0.0 and annual[-1] == 1.0


def test_time_axis_and_collection_api():
    ax = TimeAxis.from_values(np.array([2000.0, 2010.0, 2020.0]))
    assert len(ax) == 3 and ax.bounds()[-1] == 2030.0 and ax.at_bounds(2) == (2020.0, 2030.0)
    ts = Timeseries.from_values(np.array([1.0, 2.0, np.nan]), ax)
    assert ts.latest == 1 and ts.latest_value() == 2.0 and ts.at(0) == 1.0
    coll = TimeseriesCollection()
    coll.add_timeseries("x", ts)
    assert coll.names() == ["x"] and coll.get_fourbox_timeseries_by_name("x") is None
    got = coll.get_timeseries_by_name("x")
    got.set(2, 5.0)
    assert math.isnan(coll.get_timeseries_by_name("x").at(2))  # clone semantics


# ---- calibrate host classes ------------------------------------------------------------------------------------
def test_distributions_match_reference_formulas():
    assert cal.Uniform(0.0, 2.0).ln_pdf(1.0) == pytest.approx(-math.log(2.0))
    assert cal.Uniform(0.0, 2.0).ln_pdf(3.0) == -math.inf
    assert cal.Normal(0.0, 1.0).ln_pdf(0.0) == pytest.approx(-0.5 * math.log(2 * math.pi))
    assert cal.LogNormal(0.0, 1.0).ln_pdf(-1.0) == -math.inf
    assert cal.Bound(cal.Normal(0.0, 1.0), -1.0, 1.0).ln_pdf(2.0) == -math.inf
    ps = cal.ParameterSet().add("x", cal.Uniform(0.0, 1.0)).add("y", cal.Uniform(0.0, 1.0))
    assert math.isfinite(ps.log_prior([0.5, 0.5])) and ps.log_prior([1.5, 0.5]) == -math.inf
    lhs = ps.sample_lhs(16, np.random.default_rng(0))
    assert lhs.shape == (16, 2) and sorted(np.floor(lhs[:, 0] * 16).astype(int)) == list(range(16))
    assert ps._abi()[0] == (_ffi.PRIOR_UNIFORM, 0.0, 1.0, 0.0, 0.0)


def test_target_flattening_keeps_variable_grouping():
    t = cal.Target()
    t.add_observation("A", 2000.0, 1.0, 0.1).add_observation("B", 2000.0, 2.0, 0.2).add_observation("A", 2001.0, 1.5, 0.1)
    assert [o[0] for o in t._flat()] == ["A", "A", "B"]
    assert t.total_observations() == 3 and t.time_range() == (2000.0, 2001.0)
