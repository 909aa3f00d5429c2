"""Pins the CPU oracle against every golden vector / known-answer test the reference holds
for the hot path.  Runs without a GPU."""

import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as orc
from rscm_b200 import synthetic as syn
from rscm_b200.components import CarbonCycleBuilder, CO2ERFBuilder
from rscm_b200.core import InterpolationStrategy, ModelBuilder, TimeAxis, Timeseries, VariableSchema
from rscm_b200.magicc import GhgForcingBuilder
from rscm_b200.two_layer import TwoLayerBuilder

from .helpers import oracle_from_builder

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---- CO2ERF: crates/rscm-components/src/components/co2_erf.rs:94-113 ---------------
def test_co2_erf_zero_at_preindustrial():
    assert abs(orc.lib().orc_co2_erf(3.7, 278.0, 278.0)) < 1e-10


def test_co2_erf_at_2x():
    assert abs(orc.lib().orc_co2_erf(3.7, 278.0, 556.0) - 3.7) < 1e-10


# ---- compute_aggregate doctests: crates/rscm-core/src/schema.rs:741-759 -------------
def _agg(vals, op, w=None):
    v = np.array(vals, dtype=float)
    ww = None if w is None else np.array(w, dtype=float)
    return orc.lib().orc_compute_aggregate(orc._dp(v), orc._dp(ww), v.size, op)


def test_compute_aggregate_doctest_values():
    assert _agg([1.0, 2.0, 3.0], orc.AGG_SUM) == 6.0
    assert _agg([1.0, np.nan, 3.0], orc.AGG_SUM) == 4.0
    assert _agg([1.0, 2.0, 3.0], orc.AGG_MEAN) == 2.0
    assert _agg([1.0, np.nan, 3.0], orc.AGG_MEAN) == 2.0
    assert _agg([10.0, 20.0], orc.AGG_WEIGHTED, [0.3, 0.7]) == 17.0
    assert math.isnan(_agg([np.nan, np.nan], orc.AGG_SUM))


# ---- RK4 step count: ode_solvers ceil rule at the reference's call sites ------------
def test_rk4_step_counts():
    L = orc.lib()
    assert L.orc_rk4_steps(1750.0, 1751.0, 0.1) == 10      # annual axis, h = 0.1
    assert L.orc_rk4_steps(2000.0, 2005.0, 0.1) == 50      # tests/test_model.py 5-year steps
    assert L.orc_rk4_steps(1800.0, 1801.0, 1.0 / 120.0) == 120  # coupled_models.rs step_size
    assert L.orc_rk4_steps(2000.0, 2010.0, 0.1) == 100     # tests/test_two_layer.py 10-year step


# ---- Gaussian likelihood: crates/rscm-calibrate/src/likelihood.rs:296-380 ------------
def _toy_model(values_by_time):
    """A CO2ERF-only model is used as a carrier: we only need a variable table + time axis."""
    years = np.array(sorted(values_by_time))
    axis = TimeAxis.from_values(np.concatenate([years, [years[-1] + 1.0]]) if years.size == 1 else years)
    b = ModelBuilder().with_time_axis(axis).with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": 3.7, "conc_pi": 278.0}).build())
    m = oracle_from_builder(b)
    run = np.full(m.L.orc_output_size(m.h), np.nan)
    off = m.L.orc_variable_offset(m.h, m.names.index("Effective Radiative Forcing|CO2"))
    for i, t in enumerate(axis.values()):
        if t in values_by_time:
            run[off + i] = values_by_time[t]
    return m, run


def _lnl(m, run, obs, normalize=False):
    arr = m.make_obs(obs)
    return m.L.orc_ln_likelihood(m.h, orc._dp(run), arr, len(obs), 1 if normalize else 0)


V = "Effective Radiative Forcing|CO2"


def test_gaussian_likelihood_perfect_match():
    m, run = _toy_model({2020.0: 1.2, 2021.0: 1.3})
    assert _lnl(m, run, [(V, 2020.0, 1.2, 0.1), (V, 2021.0, 1.3, 0.1)]) == 0.0


def test_gaussian_likelihood_with_residuals():
    m, run = _toy_model({2020.0: 1.1, 2021.0: 0.0})
    assert abs(_lnl(m, run, [(V, 2020.0, 1.0, 0.1)]) - (-0.5)) < 1e-10


def test_gaussian_likelihood_multiple_observations():
    m, run = _toy_model({2020.0: 1.1, 2021.0: 1.2})
    assert abs(_lnl(m, run, [(V, 2020.0, 1.0, 0.1), (V, 2021.0, 1.1, 0.1)]) - (-1.0)) < 1e-10


def test_gaussian_likelihood_normalized_and_failures():
    m, run = _toy_model({2020.0: 1.1, 2021.0: np.inf})
    want = -0.5 - 0.5 * math.log(2 * math.pi) - math.log(0.1)
    assert abs(_lnl(m, run, [(V, 2020.0, 1.0, 0.1)], normalize=True) - want) < 1e-12
    assert math.isnan(_lnl(m, run, [(V, 2021.0, 1.0, 0.1)]))  # non-finite model value => Err
    m2, run2 = _toy_model({2020.0: 1.0, 2021.0: np.nan})
    assert math.isnan(_lnl(m2, run2, [(V, 2021.0, 1.0, 0.1)]))  # NaN is never extracted => missing time => Err


# ---- priors: crates/rscm-calibrate/src/distribution.rs -------------------------------------
def test_prior_ln_pdf_known_values():
    L = orc.lib()
    P = orc.OracleModel.make_priors
    assert L.orc_ln_pdf(P([(1, 0.0, 2.0)]), C_double(1.0)) == pytest.approx(-math.log(2.0), abs=1e-15)
    assert L.orc_ln_pdf(P([(1, 0.0, 2.0)]), C_double(2.5)) == -math.inf
    assert L.orc_ln_pdf(P([(2, 0.0, 1.0)]), C_double(0.0)) == pytest.approx(-0.5 * math.log(2 * math.pi), abs=1e-15)
    assert L.orc_ln_pdf(P([(3, 0.0, 1.0)]), C_double(1.0)) == pytest.approx(-0.5 * math.log(2 * math.pi), abs=1e-15)
    assert L.orc_ln_pdf(P([(3, 0.0, 1.0)]), C_double(-1.0)) == -math.inf
    assert L.orc_ln_pdf(P([(4, 0.0, 1.0, -1.0, 1.0)]), C_double(2.0)) == -math.inf
    assert L.orc_ln_pdf(P([(4, 0.0, 1.0, -1.0, 1.0)]), C_double(0.5)) == pytest.approx(-0.125 - 0.5 * math.log(2 * math.pi), abs=1e-15)


def C_double(x):
    import ctypes

    return ctypes.c_double(x)


# ---- CarbonCycle vs analytical solution: crates/rscm-components/tests/coupled_models.rs:13-141
def test_carbon_cycle_matches_reference_analytical_test():
    tau, conc_pi, conc_initial, t_initial = 20.3, 280.0, 280.0, 1800.0
    emissions_level, step_year, gtc = 10.0, 1850.0, 2.13
    axis = TimeAxis.from_values(np.arange(t_initial, 2100.0, 1.0))
    emissions = Timeseries(
        np.array([0.0, 0.0, emissions_level, emissions_level]),
        TimeAxis.from_bounds(np.array([t_initial, (t_initial + step_year) / 2.0, step_year, step_year + 50.0, 2100.0])),
        "GtC / yr", InterpolationStrategy.Previous)
    temperature = Timeseries(np.array([1.0]), TimeAxis.from_bounds(np.array([t_initial, 2100.0])), "K", InterpolationStrategy.Next)
    b = (
        ModelBuilder()
        .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": tau, "conc_pi": conc_pi, "alpha_temperature": 0.0})
                             .with_solver_options(1.0 / 120.0).build())
        .with_initial_values({"Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": conc_initial})
        .with_time_axis(axis)
        .with_exogenous_variable("Emissions|CO2|Anthropogenic", emissions)
        .with_exogenous_variable("Surface Temperature", temperature)
    )
    res = oracle_from_builder(b).run()
    t = axis.values()
    expected_em = np.where(t < step_year, 0.0, emissions_level)
    assert np.array_equal(res["Emissions|CO2|Anthropogenic"], expected_em)  # assert_eq! in the reference
    before = (conc_initial - conc_pi) * np.exp(-(t - t_initial) / tau) + conc_pi
    after = emissions_level / gtc * tau * (1.0 - np.exp(-(t - step_year) / tau)) + before
    expected = np.where(t < step_year, before, after)
    rel = np.abs(res["Atmospheric Concentration|CO2"] - expected) / np.abs(expected)
    assert rel.max() < 0.01  # the reference's bound
    assert rel.max() < 1e-12  # what the restated index convention actually achieves


# ---- TwoLayer qualitative tests: crates/rscm-two-layer/src/component.rs:300-406 ---------------
def _two_layer_one_step(erf):
    axis = TimeAxis.from_bounds(np.array([2000.0, 2001.0, 2002.0]))
    b = syn.two_layer_builder(axis=axis).with_exogenous_variable(
        "Effective Radiative Forcing", Timeseries(np.array([erf, erf]), axis, "W/m^2", InterpolationStrategy.Linear))
    return oracle_from_builder(b).run()["Surface Temperature"][1]


def test_two_layer_reference_qualitative_cases():
    t4 = _two_layer_one_step(4.0)
    assert 0.0 < t4 < 4.0
    assert abs(_two_layer_one_step(0.0)) < 1e-10
    assert _two_layer_one_step(-2.0) < 0.0
    t2 = _two_layer_one_step(2.0)
    assert t4 > t2 and abs(t4 / t2 - 2.0) < 0.1


def test_two_layer_rk4_is_fourth_order():
    """Not in the reference: the restated RK4 converges at order 4 to the exact linear solution."""
    lam, eps, eta, cs, cd, F = 1.0, 1.0, 0.7, 8.0, 100.0, 4.0
    A = np.array([[-(lam + eps * eta) / cs, eps * eta / cs], [eta / cd, -eta / cd]])
    w, Vv = np.linalg.eig(A)
    y_eq = np.linalg.solve(A, -np.array([F / cs, 0.0]))
    exact = y_eq + Vv @ (np.exp(w * 1.0) * np.linalg.solve(Vv, -y_eq))
    got = _two_layer_one_step(F)
    # RK4 h = 0.1 over one year: global error ~ C h^4
    assert abs(got - exact[0].real) < 5e-7


# ---- GhgForcing vs MAGICC7 golden vectors: tests/regression/test_ghg_forcing.py:237-331 ---------
@pytest.mark.parametrize("name", ["01", "02"])
def test_ghg_forcing_golden(name):
    d = np.load(os.path.join(GOLDEN, f"ghg_forcing_{name}.npz"))
    cfg = json.loads(str(d["config"]))
    method = {"IPCCTAR": "Ipcctar", "OLBL": "Olbl"}[cfg["core_co2ch4n2o_rfmethod"]]
    dflt = (1.0, 1.0, 1.0) if method == "Ipcctar" else (1.05, 0.86, 1.0)
    years = d["years"]
    params = {
        "method": method, "delq2xco2": cfg.get("core_delq2xco2", 3.71),
        "co2_pi": float(d["co2"][0]), "ch4_pi": float(d["ch4"][0]), "n2o_pi": float(d["n2o"][0]),
        "adjust_co2": cfg.get("core_rfrapidadjust_co2", dflt[0]), "adjust_ch4": cfg.get("core_rfrapidadjust_ch4", dflt[1]),
        "adjust_n2o": cfg.get("core_rfrapidadjust_n2o", dflt[2]),
    }
    axis = TimeAxis.from_bounds(np.concatenate([years, [years[-1] + 1.0]]))
    b = ModelBuilder().with_time_axis(axis).with_rust_component(GhgForcingBuilder.from_parameters(params).build())
    for var, key, unit in [("CO2", "co2", "ppm"), ("CH4", "ch4", "ppb"), ("N2O", "n2o", "ppb")]:
        b.with_exogenous_variable(f"Atmospheric Concentration|{var}", Timeseries(d[key], axis, unit, InterpolationStrategy.Linear))
    res = oracle_from_builder(b).run()
    for var, key in [("CO2", "erf_co2"), ("CH4", "erf_ch4"), ("N2O", "erf_n2o")]:
        actual = res[f"Effective Radiative Forcing|{var}"]
        assert math.isnan(actual[0])  # pure outputs keep NaN at index 0
        np.testing.assert_allclose(actual[1:], d[key][:-1], rtol=1e-5, atol=1e-6)


# ---- framework semantics -----------------------------------------------------------------------------
def test_coupled_graph_classification_and_order():
    """insertion-order classification, aggregator placement, BFS chain."""
    b, *_ = syn.config3(M=1, S=1)
    m = oracle_from_builder(b, {"Emissions|CO2|Anthropogenic": np.zeros(351)})
    assert m.execution_order() == [0, 1, 3, 2]  # CarbonCycle, CO2ERF, Aggregator, TwoLayer
    assert m.variable_source(0, "Surface Temperature") == orc.lib().orc_variable_source(m.h, 0, b"Surface Temperature") == 0  # Exogenous => lagged T[N]
    assert m.variable_source(0, "Atmospheric Concentration|CO2") == 1  # OwnState
    assert m.variable_source(1, "Atmospheric Concentration|CO2") == 2  # UpstreamOutput => C[N+1]
    assert m.variable_source(2, "Effective Radiative Forcing") == 2   # aggregate name => ERF[N+1]
    assert m.is_endogenous("Surface Temperature") and not m.is_endogenous("Emissions|CO2|Anthropogenic")


def test_coupled_semantics_nan_and_lag():
    b, _, params, scen = syn.config3(M=1, S=8)
    m = oracle_from_builder(b, scen[4])
    r = m.run()
    # pure outputs stay NaN at index 0; states carry their initial value
    assert math.isnan(r["Effective Radiative Forcing|CO2"][0]) and math.isnan(r["Effective Radiative Forcing"][0])
    assert r["Atmospheric Concentration|CO2"][0] == 278.0 and r["Surface Temperature"][0] == 0.0
    assert not np.isnan(r["Surface Temperature"][1:]).any()
    # ERF[N+1] is computed from C[N+1] (same index), aggregate of one contributor is the identity
    np.testing.assert_array_equal(r["Effective Radiative Forcing"][1:], r["Effective Radiative Forcing|CO2"][1:])
    c = r["Atmospheric Concentration|CO2"]
    np.testing.assert_allclose(r["Effective Radiative Forcing|CO2"][1:], 3.7 / math.log(2) * np.log(1 + (c[1:] - 278.0) / 278.0), rtol=1e-14)
    # cumulative emissions integrate E[N] held over the step (emissions read at index N)
    e = scen[4]["Emissions|CO2|Anthropogenic"]
    np.testing.assert_allclose(r["Cumulative Emissions|CO2"][1:], np.cumsum(e[:-1]), rtol=1e-12)


def test_two_layer_reads_exogenous_erf_at_start_index():
    """configs 1-2: exogenous ERF => F = ERF[N]."""
    axis = TimeAxis.from_values(np.array([2000.0, 2001.0, 2002.0, 2003.0]))
    erf = np.array([0.0, 4.0, 0.0, 0.0])
    b = syn.two_layer_builder(axis=axis).with_exogenous_variable("Effective Radiative Forcing", Timeseries(erf, axis, "W/m^2", InterpolationStrategy.Linear))
    ts = oracle_from_builder(b).run()["Surface Temperature"]
    assert ts[1] == 0.0 and ts[2] > 0.0


def test_missing_state_initial_value_is_an_error():
    b = ModelBuilder().with_time_axis(syn.time_axis()).with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
    with pytest.raises(RuntimeError, match="initial value"):
        oracle_from_builder(b)


def test_rk4_assertion_failure_leaves_nan():
    """A step the RK4 grid cannot land on (|t_last - t_next| >= 5e-3) makes the reference panic; the
    oracle reports it as a failed component: outputs stay NaN and the status flag is set."""
    axis = TimeAxis.from_bounds(np.array([2000.0, 2001.0, 2002.05, 2003.05]))
    b = syn.two_layer_builder(axis=axis)
    m = oracle_from_builder(b, {"Effective Radiative Forcing": np.array([1.0, 1.0, 1.0])})
    out, status = m.run_batch([], np.zeros((1, 0)), ["Effective Radiative Forcing"], np.array([[1.0, 1.0, 1.0]]), ["Surface Temperature"], want_status=True)
    assert status[0] == 1 and not math.isnan(out[1, 0]) and math.isnan(out[2, 0])
