"""GPU parity tests: the fused sm_100a kernels (through the C ABI) against the CPU oracle on the
same seeded inputs, against the committed golden fixtures, and — at BASELINE.json's full sizes —
through size-independent properties.

Tolerances: fp64 path 1e-9 relative on every output series (north_star); fp32 path 1e-4.
NaN positions must match exactly."""

import json
import math
import os

import numpy as np
import pytest

from rscm_b200 import _ffi, calibrate as cal
from rscm_b200 import synthetic as syn
from rscm_b200.components import CarbonCycleBuilder, CO2ERFBuilder
from rscm_b200.core import InterpolationStrategy, ModelBuilder, TimeAxis, Timeseries, VariableSchema
from rscm_b200.magicc import GhgForcingBuilder

from .helpers import elementwise_rel_err, oracle_bindings, oracle_from_builder, rel_err

pytestmark = pytest.mark.gpu

TOL64 = 1e-9
TOL32 = 1e-4
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def gpu_vs_oracle(builder, bindings, params, scenarios, outputs=None, dtype="f64", tol=TOL64, layout=1):
    ens = builder.build_ensemble(dtype=dtype)
    ens.bind_parameters(bindings)
    names = outputs or ens.variable_names
    ens.select_outputs(names)
    sc = ens.pack_scenarios(scenarios)
    p = params if layout == 1 else np.ascontiguousarray(params.T)
    status = np.zeros(len(scenarios) * params.shape[0], dtype=np.uint8)
    out = ens.run(p, sc, layout=layout, status=status)
    got = ens.split_outputs(out)
    m = oracle_from_builder(builder)
    ref_out, ref_status = m.run_batch(oracle_bindings(builder, bindings), params, ens.exogenous_names, sc, names, want_status=True)
    ref = m.split(ref_out, names)
    worst = 0.0
    for n in names:
        e = rel_err(got[n], ref[n])
        worst = max(worst, e)
        assert e <= tol, f"{n}: relative error {e:.3e} > {tol}"
    assert np.array_equal(status & 1, ref_status & 1)
    return got, ref, worst, ens


# ---- config 1: single two-layer run (configs/two-layer defaults) ----------------------------------------
def test_config1_single_two_layer_run_model_api():
    axis = syn.time_axis()
    forcing = syn.ssp_like_forcing(axis.values())
    b = syn.two_layer_builder(axis=axis).with_exogenous_variable(
        "Effective Radiative Forcing", Timeseries(forcing, axis, "W/m^2", InterpolationStrategy.Linear))
    model = b.build()
    model.run()
    assert model.finished()
    res = model.timeseries()
    ref = oracle_from_builder(b).run()
    for n in ("Surface Temperature", "Deep Ocean Temperature", "Effective Radiative Forcing"):
        assert rel_err(res.get_timeseries_by_name(n).values(), ref[n]) <= TOL64
    ts = res.get_timeseries_by_name("Surface Temperature").values()
    assert ts[0] == 0.0 and 1.5 < ts[-1] < 6.0


def test_model_step_reveals_one_index_at_a_time():
    axis = TimeAxis.from_values(np.arange(2000.0, 2006.0))
    b = syn.two_layer_builder(axis=axis).with_exogenous_variable(
        "Effective Radiative Forcing", Timeseries(np.full(6, 2.0), axis, "W/m^2", InterpolationStrategy.Linear))
    model = b.build()
    model.step()
    model.step()
    v = model.timeseries().get_timeseries_by_name("Surface Temperature").values()
    assert not np.isnan(v[:3]).any() and np.isnan(v[3:]).all() and model.current_time() == 2002.0


# ---- config 2: two-layer ensemble -------------------------------------------------------------------------
@pytest.mark.parametrize("M,layout", [(1, 1), (127, 1), (128, 0), (1000, 1), (4097, 0)])
def test_two_layer_ensemble_parity_ragged_sizes(M, layout):
    b, binds, params, scen = syn.config2(M=M)
    gpu_vs_oracle(b, binds, params, scen, layout=layout)


def test_two_layer_nonlinear_feedback_and_initial_value_binding():
    b, binds, params, scen = syn.config2(M=300)
    binds = dict(binds)
    binds["T0"] = "initial:Surface Temperature"
    rng = np.random.default_rng(5)
    params = np.column_stack([params, rng.uniform(-0.5, 0.5, 300)])
    got, ref, _, _ = gpu_vs_oracle(b, binds, params, scen)
    np.testing.assert_array_equal(got["Surface Temperature"][0], params[:, -1])


# ---- config 3: coupled carbon cycle + CO2 ERF + two-layer ---------------------------------------------------
def test_coupled_parity_all_series():
    b, binds, params, scen = syn.config3(M=1500, S=8)
    got, ref, worst, ens = gpu_vs_oracle(b, binds, params, scen)
    assert ens.execution_order() == [0, 1, 3, 2]
    # NaN at index 0 for pure outputs, for every run
    assert np.isnan(got["Effective Radiative Forcing"][0]).all() and not np.isnan(got["Effective Radiative Forcing"][1:]).any()
    # element-wise check as well (floor = 1e-6 of each series' scale)
    for n in syn.COUPLED_OUTPUTS:
        floor = 1e-6 * np.nanmax(np.abs(ref[n]))
        assert elementwise_rel_err(got[n], ref[n], floor) <= 1e-7, n


def test_coupled_conc_pi_bound_to_two_components():
    b, binds, params, scen = syn.config3(M=257, S=2)
    binds = dict(binds)
    binds["conc_pi"] = ["CarbonCycle.conc_pi", "CO2ERF.conc_pi"]
    binds["C0"] = "initial:Atmospheric Concentration|CO2"
    cpi = np.random.default_rng(3).uniform(270.0, 290.0, 257)
    params = np.column_stack([params, cpi, cpi])
    gpu_vs_oracle(b, binds, params, scen)


def test_output_and_time_subselection_matches_full_run():
    b, binds, params, scen = syn.config3(M=300, S=3)
    full, _, _, _ = gpu_vs_oracle(b, binds, params, scen)
    ens = b.build_ensemble().bind_parameters(binds)
    ens.select_outputs(["Surface Temperature", "Atmospheric Concentration|CO2"], t_start=100, t_stop=351, t_step=25)
    out = ens.split_outputs(ens.run(params, ens.pack_scenarios(scen)))
    idx = list(range(100, 351, 25))
    np.testing.assert_array_equal(out["Surface Temperature"], full["Surface Temperature"][idx])
    np.testing.assert_array_equal(out["Atmospheric Concentration|CO2"], full["Atmospheric Concentration|CO2"][idx])


def test_members_are_independent_and_deterministic():
    """Per-member results are bit-identical whatever the batch they run in."""
    b, binds, params, scen = syn.config3(M=1024, S=2)
    ens = b.build_ensemble().bind_parameters(binds)
    ens.select_outputs(["Surface Temperature"])
    sc = ens.pack_scenarios(scen)
    a = ens.run(params, sc)
    a2 = ens.run(params, sc)
    np.testing.assert_array_equal(a, a2)
    part = ens.run(params[300:700], sc)
    np.testing.assert_array_equal(part.reshape(351, 2, 400), a.reshape(351, 2, 1024)[:, :, 300:700])


# ---- other named graphs ----------------------------------------------------------------------------------------
def test_carbon_cycle_reference_analytical_case_on_gpu():
    """crates/rscm-components/tests/coupled_models.rs:13-141 through the product API."""
    tau, conc_pi, t0, level, step_year = 20.3, 280.0, 1800.0, 10.0, 1850.0
    axis = TimeAxis.from_values(np.arange(t0, 2100.0, 1.0))
    em = Timeseries(np.array([0.0, 0.0, level, level]), TimeAxis.from_bounds(np.array([t0, (t0 + step_year) / 2, step_year, step_year + 50.0, 2100.0])),
                    "GtC / yr", InterpolationStrategy.Previous)
    temp = Timeseries(np.array([1.0]), TimeAxis.from_bounds(np.array([t0, 2100.0])), "K", InterpolationStrategy.Next)
    b = (ModelBuilder()
         .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": tau, "conc_pi": conc_pi, "alpha_temperature": 0.0}).with_solver_options(1.0 / 120.0).build())
         .with_initial_values({"Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": conc_pi})
         .with_time_axis(axis).with_exogenous_variable("Emissions|CO2|Anthropogenic", em).with_exogenous_variable("Surface Temperature", temp))
    model = b.build()
    model.run()
    conc = model.timeseries().get_timeseries_by_name("Atmospheric Concentration|CO2").values()
    t = axis.values()
    before = (conc_pi - conc_pi) * np.exp(-(t - t0) / tau) + conc_pi
    after = level / 2.13 * tau * (1.0 - np.exp(-(t - step_year) / tau)) + before
    expected = np.where(t < step_year, before, after)
    assert np.max(np.abs(conc - expected) / expected) < 0.01   # reference bound
    assert np.max(np.abs(conc - expected) / expected) < 1e-11
    ref = oracle_from_builder(b).run()
    for n in ref:
        assert rel_err(model.timeseries().get_timeseries_by_name(n).values(), ref[n]) <= TOL64, n


def test_carbon_cycle_with_co2erf_exogenous_temperature():
    """crates/rscm-components/tests/coupled_models.rs:143-224 graph, numerically against the oracle."""
    axis = syn.time_axis()
    b = (ModelBuilder().with_time_axis(axis)
         .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": 20.3, "conc_pi": 280.0, "alpha_temperature": 0.05}).build())
         .with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": 4.0, "conc_pi": 280.0}).build())
         .with_initial_values({"Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": 300.0}))
    years = axis.values()
    scen = [{"Emissions|CO2|Anthropogenic": syn.emission_scenarios(years, 3)[s], "Surface Temperature": 0.01 * s * (years - 1750.0) / 10.0} for s in range(3)]
    binds = {"tau": "CarbonCycle.tau", "alpha": "CarbonCycle.alpha_temperature"}
    params = syn.uniform_params({"tau": (15.0, 40.0), "alpha": (0.0, 0.15)}, 200, 11)
    got, _, _, ens = gpu_vs_oracle(b, binds, params, scen)
    assert sorted(ens.variable_names) == sorted([
        "Atmospheric Concentration|CO2", "Cumulative Emissions|CO2", "Cumulative Land Uptake",
        "Effective Radiative Forcing|CO2", "Emissions|CO2|Anthropogenic", "Surface Temperature"])


@pytest.mark.parametrize("name", ["01", "02"])
def test_ghg_forcing_golden_on_gpu(name):
    """MAGICC7 golden vectors (tests/regression/test_ghg_forcing.py:237-331) through the CUDA path."""
    d = np.load(os.path.join(GOLDEN, f"ghg_forcing_{name}.npz"))
    cfg = json.loads(str(d["config"]))
    method = {"IPCCTAR": "Ipcctar", "OLBL": "Olbl"}[cfg["core_co2ch4n2o_rfmethod"]]
    dflt = (1.0, 1.0, 1.0) if method == "Ipcctar" else (1.05, 0.86, 1.0)
    years = d["years"]
    params = {"method": method, "delq2xco2": cfg.get("core_delq2xco2", 3.71), "co2_pi": float(d["co2"][0]), "ch4_pi": float(d["ch4"][0]),
              "n2o_pi": float(d["n2o"][0]), "adjust_co2": cfg.get("core_rfrapidadjust_co2", dflt[0]),
              "adjust_ch4": cfg.get("core_rfrapidadjust_ch4", dflt[1]), "adjust_n2o": cfg.get("core_rfrapidadjust_n2o", dflt[2])}
    axis = TimeAxis.from_bounds(np.concatenate([years, [years[-1] + 1.0]]))
    b = ModelBuilder().with_time_axis(axis).with_rust_component(GhgForcingBuilder.from_parameters(params).build())
    for var, key, unit in [("CO2", "co2", "ppm"), ("CH4", "ch4", "ppb"), ("N2O", "n2o", "ppb")]:
        b.with_exogenous_variable(f"Atmospheric Concentration|{var}", Timeseries(d[key], axis, unit, InterpolationStrategy.Linear))
    model = b.build()
    model.run()
    res = model.timeseries()
    ref = oracle_from_builder(b).run()
    for var, key in [("CO2", "erf_co2"), ("CH4", "erf_ch4"), ("N2O", "erf_n2o")]:
        actual = res.get_timeseries_by_name(f"Effective Radiative Forcing|{var}").values()
        assert math.isnan(actual[0])
        np.testing.assert_allclose(actual[1:], d[key][:-1], rtol=1e-5, atol=1e-6)
        assert rel_err(actual, ref[f"Effective Radiative Forcing|{var}"]) <= TOL64


def test_ghg_forcing_into_two_layer_with_sum_aggregate():
    d = np.load(os.path.join(GOLDEN, "ghg_forcing_02.npz"))
    axis = syn.time_axis()
    schema = VariableSchema()
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "ppm")
        schema.add_variable(f"Effective Radiative Forcing|{n}", "W/m^2")
    schema.add_variable("Surface Temperature", "K")
    schema.add_variable("Deep Ocean Temperature", "K")
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", [f"Effective Radiative Forcing|{n}" for n in ("CO2", "CH4", "N2O")])
    from rscm_b200.two_layer import TwoLayerBuilder
    b = (ModelBuilder().with_time_axis(axis).with_schema(schema)
         .with_rust_component(GhgForcingBuilder.from_parameters({}).build())
         .with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
         .with_initial_values({"Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0}))
    scen = [{f"Atmospheric Concentration|{n}": d[k] * f for n, k in (("CO2", "co2"), ("CH4", "ch4"), ("N2O", "n2o"))} for f in (1.0, 1.1)]
    binds = {**syn.TWO_LAYER_BINDINGS, "adjust_co2": "GhgForcing.adjust_co2", "co2_pi": "GhgForcing.co2_pi"}
    p = np.column_stack([syn.uniform_params(syn.TWO_LAYER_RANGES, 130, 2), np.random.default_rng(1).uniform(0.9, 1.1, 130),
                         np.random.default_rng(2).uniform(270, 285, 130)])
    gpu_vs_oracle(b, binds, p, scen)


def test_run_time_compiled_graph_parity(tmp_path, monkeypatch):
    """A graph outside the AOT registry (CO2ERF on exogenous concentrations -> Sum aggregate -> TwoLayer):
    emitted, compiled by NVRTC for sm_100a, loaded through the driver API, and checked like the others."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    from rscm_b200.two_layer import TwoLayerBuilder
    axis = syn.time_axis()
    schema = VariableSchema()
    schema.add_variable("Atmospheric Concentration|CO2", "ppm")
    schema.add_variable("Effective Radiative Forcing|CO2", "W/m^2")
    schema.add_variable("Effective Radiative Forcing|Other", "W/m^2")
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Weighted", ["Effective Radiative Forcing|CO2", "Effective Radiative Forcing|Other"],
                         weights=[1.0, 0.5])
    b = (ModelBuilder().with_time_axis(axis).with_schema(schema)
         .with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": 3.7, "conc_pi": 278.0}).build())
         .with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
         .with_initial_values({"Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0}))
    years = axis.values()
    conc = 278.0 * np.exp(0.004 * np.maximum(0.0, years - 1850.0))
    other = np.where(years > 1990, -0.3, np.nan)  # exogenous contributor with NaNs: dropped by the aggregate while NaN
    scen = [{"Atmospheric Concentration|CO2": conc * f, "Effective Radiative Forcing|Other": other} for f in (1.0, 1.05)]
    binds = {**syn.TWO_LAYER_BINDINGS, "erf_2xco2": "CO2ERF.erf_2xco2"}
    p = np.column_stack([syn.uniform_params(syn.TWO_LAYER_RANGES, 200, 9), np.random.default_rng(4).uniform(3.4, 4.0, 200)])
    got, ref, _, ens = gpu_vs_oracle(b, binds, p, scen)
    assert ens.program_is_jit()
    assert not np.isnan(got["Effective Radiative Forcing"][1:]).any()


# ---- failure semantics -------------------------------------------------------------------------------------------
def test_rk4_assertion_failure_is_data_not_an_error():
    axis = TimeAxis.from_bounds(np.array([2000.0, 2001.0, 2002.05, 2003.05, 2004.05]))
    b = syn.two_layer_builder(axis=axis)
    params = syn.uniform_params(syn.TWO_LAYER_RANGES, 64, 1)
    got, ref, _, ens = gpu_vs_oracle(b, syn.TWO_LAYER_BINDINGS, params, [{"Effective Radiative Forcing": np.ones(4)}])
    assert np.isnan(got["Surface Temperature"][2:]).all() and not np.isnan(got["Surface Temperature"][:2]).any()


def test_non_finite_member_does_not_affect_others():
    b, binds, params, scen = syn.config2(M=256)
    params = params.copy()
    params[7, 4] = 0.0   # heat_capacity_surface = 0 -> division by zero -> inf/NaN for that member only
    ens = b.build_ensemble().bind_parameters(binds)
    ens.select_outputs(["Surface Temperature"])
    status = np.zeros(256, dtype=np.uint8)
    out = ens.run(params, ens.pack_scenarios(scen), status=status)
    assert status[7] & 2 and not (np.delete(status, 7) & 2).any()
    assert not np.isfinite(out[-1, 7]) and np.isfinite(np.delete(out[-1], 7)).all()


# ---- calibration: log-posterior ---------------------------------------------------------------------------------------
def _calibration_case(M=500, normalize=False):
    b, binds, params, scen = syn.config2(M=M)
    ens = b.build_ensemble().bind_parameters(binds)
    sc = ens.pack_scenarios(scen)
    truth = dict(syn.TWO_LAYER_DEFAULTS, lambda0=1.1, efficacy=1.3, a=0.05)  # tests/test_calibration_integration.py:37-44
    ens.select_outputs(["Surface Temperature"])
    t_true = ens.run(np.array([[truth[k] for k in syn.TWO_LAYER_RANGES]]), sc)[:, 0]
    obs = syn.config5_observations(t_true, ens._time_axis.values())
    priors = [(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.TWO_LAYER_RANGES.values()]
    ens.set_target(obs, normalize=normalize).set_priors(priors)
    return b, binds, params, scen, sc, ens, obs, priors


@pytest.mark.parametrize("normalize", [False, True])
def test_log_posterior_matches_oracle(normalize):
    b, binds, params, scen, sc, ens, obs, priors = _calibration_case(normalize=normalize)
    params = params.copy()
    params[3, 0] = 5.0                    # outside the uniform prior -> -inf
    params[4, 4] = 0.0                    # non-finite model output -> -inf
    lp, summ = ens.log_posterior(params, sc, with_summary=True)
    m = oracle_from_builder(b)
    ref = m.log_posterior_batch(oracle_bindings(b, binds), params, ens.exogenous_names, sc, priors, obs, normalize=normalize)
    assert len(obs) == 171
    assert lp[3] == -np.inf and lp[4] == -np.inf and np.array_equal(np.isinf(lp), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.max(np.abs(lp[fin] - ref[fin]) / np.abs(ref[fin])) <= TOL64
    # block-reduced summary (warp shuffle + block + last-block pass)
    assert summ["n_runs"] == params.shape[0] and summ["n_finite"] == fin.sum()
    assert summ["argmax"] == int(np.argmax(np.where(fin, lp, -np.inf))) and summ["max_logpost"] == lp[summ["argmax"]]
    assert abs(summ["sum_finite"] - lp[fin].sum()) <= 1e-9 * abs(lp[fin].sum())


def test_log_posterior_multiple_variables_and_mixed_priors():
    b, binds, params, scen = syn.config3(M=200, S=2)
    ens = b.build_ensemble().bind_parameters(binds)
    sc = ens.pack_scenarios(scen)
    obs = [("Surface Temperature", 1900.0, 0.1, 0.1), ("Surface Temperature", 2000.0, 0.8, 0.15), ("Surface Temperature", 1750.0, 0.0, 0.5),
           ("Atmospheric Concentration|CO2", 2000.0, 370.0, 5.0), ("Atmospheric Concentration|CO2", 2020.0, 410.0, 5.0)]
    priors = [(_ffi.PRIOR_NORMAL, 25.0, 10.0), (_ffi.PRIOR_BOUND_NORMAL, 0.05, 0.05, 0.0, 0.12), (_ffi.PRIOR_LOGNORMAL, 1.3, 0.2)] + \
             [(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.TWO_LAYER_RANGES.values()]
    ens.set_target(obs).set_priors(priors)
    lp = ens.log_posterior(params, sc)
    ref = oracle_from_builder(b).log_posterior_batch(oracle_bindings(b, binds), params, ens.exogenous_names, sc, priors, obs)
    assert np.array_equal(np.isinf(lp), np.isinf(ref)) and np.isinf(ref).any() and np.isfinite(ref).any()
    fin = np.isfinite(ref)
    assert np.max(np.abs(lp[fin] - ref[fin]) / np.abs(ref[fin])) <= TOL64


def test_model_runner_and_sampler_recover_parameters():
    """tests/test_calibration_integration.py-style statistical check through the reference-shaped API."""
    b, _, _, scen = syn.config2(M=1)
    runner = cal.ModelRunner(b, {"lambda0": "TwoLayer.lambda0", "efficacy": "TwoLayer.efficacy"}, ["Surface Temperature"],
                             scenarios=None)
    ens = runner.ensemble
    runner._scenarios = ens.pack_scenarios(scen)
    single = runner.run([1.1, 1.3])
    assert set(single) == {"Surface Temperature"} and len(single["Surface Temperature"]) == 351 and single["Surface Temperature"][1750.0] == 0.0
    with pytest.raises(ValueError):
        runner.run([1.0])
    truth = np.array([single["Surface Temperature"][float(y)] for y in range(1850, 2021)])
    target = cal.Target()
    rng = np.random.default_rng(42)
    for y, v in zip(range(1850, 2021), truth):
        target.add_observation("Surface Temperature", float(y), float(v + 0.05 * rng.standard_normal()), 0.05)
    ps = cal.ParameterSet().add("lambda0", cal.Uniform(0.5, 2.0)).add("efficacy", cal.Uniform(0.5, 2.5))
    sampler = cal.EnsembleSampler(ps, runner, cal.GaussianLikelihood(), target, seed=7)
    chain = sampler.run(300, cal.WalkerInit.from_prior(), thin=1)
    flat = chain.flat_samples(discard=150)
    assert abs(flat[:, 0].mean() - 1.1) < 0.1 and 0.1 < sampler.acceptance_rate < 0.9
    assert chain.total_iterations == 300 and flat.shape[1] == 2


# ---- fp32 path -----------------------------------------------------------------------------------------------------
def test_fp32_path_within_1e4():
    b, binds, params, scen = syn.config3(M=512, S=2)
    names = ["Atmospheric Concentration|CO2", "Effective Radiative Forcing", "Surface Temperature", "Deep Ocean Temperature",
             "Cumulative Emissions|CO2"]
    _, _, worst, _ = gpu_vs_oracle(b, binds, params, scen, outputs=names, dtype="f32", tol=TOL32)
    assert worst > 1e-9  # it really is a different (single-precision) path


# ---- full-size properties (BASELINE config sizes; no oracle) ----------------------------------------------------------
def test_full_size_two_layer_linearity_and_subsample_parity():
    """1M two-layer members: with a = 0 the model is linear in the forcing, so doubling the scenario doubles
    every temperature (1e-12); a strided subsample is checked against the oracle."""
    torch = pytest.importorskip("torch")
    M = 1 << 20
    b, binds, params, scen = syn.config2(M=M)
    params = params.copy()
    params[:, 1] = 0.0
    f = scen[0]["Effective Radiative Forcing"]
    ens = b.build_ensemble().bind_parameters(binds)
    ens.select_outputs(["Surface Temperature", "Deep Ocean Temperature"])
    sc = torch.from_numpy(ens.pack_scenarios([{"Effective Radiative Forcing": f}, {"Effective Radiative Forcing": 2.0 * f}])).cuda()
    p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    out = torch.empty((ens.output_rows, 2 * M), dtype=torch.float64, device="cuda")
    ens.run_device(p, sc, out, layout=0)
    torch.cuda.synchronize()
    a, c = out[:, :M], out[:, M:]
    assert torch.isfinite(out).all()
    err = ((c - 2.0 * a).abs().max() / c.abs().max()).item()
    assert err < 1e-12
    idx = np.arange(0, M, 4099)
    sub = out[:, torch.from_numpy(idx).cuda()].cpu().numpy()
    m = oracle_from_builder(b)
    ref = m.run_batch(oracle_bindings(b, binds), params[idx], ens.exogenous_names, ens.pack_scenarios(scen),
                      ["Surface Temperature", "Deep Ocean Temperature"])
    assert rel_err(sub, ref) <= TOL64


def test_full_size_coupled_invariants():
    """256k x 8 coupled runs: cumulative emissions are member-independent and equal the scenario's running sum;
    ERF == ERF|CO2; uptake + atmospheric burden closes the carbon budget to round-off."""
    torch = pytest.importorskip("torch")
    M, S = 1 << 18, 8
    b, binds, params, scen = syn.config3(M=M, S=S)
    ens = b.build_ensemble().bind_parameters(binds)
    names = ["Cumulative Emissions|CO2", "Cumulative Land Uptake", "Atmospheric Concentration|CO2",
             "Effective Radiative Forcing|CO2", "Effective Radiative Forcing"]
    ens.select_outputs(names, t_start=0, t_stop=351, t_step=50)
    sc = torch.from_numpy(ens.pack_scenarios(scen)).cuda()
    p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    out = torch.empty((ens.output_rows, S * M), dtype=torch.float64, device="cuda")
    status = torch.zeros(S * M, dtype=torch.uint8, device="cuda")
    ens.run_device(p, sc, out, status, layout=0)
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    nt = len(range(0, 351, 50))
    o = out.view(len(names), nt, S, M)
    years = ens._time_axis.values()
    em = syn.emission_scenarios(years, S)
    for s in range(S):
        want = torch.from_numpy(np.concatenate([[0.0], np.cumsum(em[s][:-1])])[::50]).cuda()
        got = o[0, :, s, :]
        assert ((got - want[:, None]).abs().max() / max(want.abs().max().item(), 1.0)) < 1e-12
    assert torch.equal(o[3, 1:], o[4, 1:]) and torch.isnan(o[3, 0]).all()
    # budget: GTC_PER_PPM * (C - C0) + uptake == cumulative emissions
    closure = 2.13 * (o[2] - 278.0) + o[1] - o[0]
    assert (closure.abs().max() / o[0].abs().max()).item() < 1e-10


# ---- irregular time axes and missing data ----------------------------------------------------------------------------
def test_irregular_time_axis_parity():
    """Annual, then 5-year, then 10-year steps (tests/test_model.py:9-46 and tests/test_two_layer.py:12-56 run such axes):
    the RK4 sub-step count follows each step's length (50 and 100 sub-steps), outputs land on the coarse points."""
    values = np.concatenate([np.arange(1750.0, 1800.0), np.arange(1800.0, 1900.0, 5.0), np.arange(1900.0, 2101.0, 10.0)])
    axis = TimeAxis.from_values(values)
    b = syn.coupled_builder(axis=axis)
    rng = np.random.default_rng(12)
    params = syn.uniform_params(syn.COUPLED_RANGES, 130, 21)
    scen = [{"Emissions|CO2|Anthropogenic": np.clip(0.02 * (values - 1750.0) * f + rng.normal(0, 0.1, values.size), 0.0, None)} for f in (1.0, 2.0)]
    got, _, worst, ens = gpu_vs_oracle(b, syn.COUPLED_BINDINGS, params, scen)
    assert ens.n_times == values.size and np.isfinite(got["Surface Temperature"]).all() and worst < 1e-10


def test_nan_in_the_scenario_propagates_like_the_reference():
    """A missing (NaN) emission value poisons the carbon cycle from that step on and everything downstream of it, for
    that scenario only; positions must match the oracle exactly (rel_err asserts it)."""
    b, binds, params, scen = syn.config3(M=64, S=2)
    scen[1]["Emissions|CO2|Anthropogenic"] = scen[1]["Emissions|CO2|Anthropogenic"].copy()
    scen[1]["Emissions|CO2|Anthropogenic"][200] = np.nan
    got, _, _, _ = gpu_vs_oracle(b, binds, params, scen)
    conc = got["Atmospheric Concentration|CO2"]
    assert np.isfinite(conc[:, :64]).all() and np.isfinite(conc[:201, 64:]).all() and np.isnan(conc[201:, 64:]).all()
    assert np.isnan(got["Surface Temperature"][202:, 64:]).all()


def test_two_layer_near_runaway_members_keep_parity():
    """lambda0 - a*T reaches zero near T = lambda0/a: with a close to 0.1 and a weak lambda0 the feedback collapses inside the
    run and the temperatures blow up (ill-conditioned: rounding is amplified ~1e5 on the way).  These are the members that
    decide the worst-case parity of BASELINE config 2 — and the ones a cancellation-prone reformulation would break."""
    b, binds, _, scen = syn.config2(M=4)
    lam, a = np.meshgrid(np.linspace(0.8, 1.0, 16), np.linspace(0.07, 0.1, 16))
    params = np.tile(np.array([syn.TWO_LAYER_DEFAULTS[k] for k in syn.TWO_LAYER_RANGES]), (256, 1))
    params[:, 0], params[:, 1] = lam.ravel(), a.ravel()
    got, ref, worst, _ = gpu_vs_oracle(b, binds, params, scen)
    t_end = ref["Surface Temperature"][-1]
    assert (np.abs(t_end) > 50.0).any() or np.isnan(t_end).any(), "the grid must contain members that run away"
    assert (np.abs(t_end) < 10.0).any() and worst <= 1e-9


# ---- full-size properties of configs 4 and 5 ---------------------------------------------------------------------------
def test_full_size_config4_determinism_heat_budget_and_subsample_parity(tmp_path, monkeypatch):
    """100 000 MAGICC-box members (BASELINE config 4): members with identical parameters give bit-identical series wherever
    they sit in the grid; the area-weighted four-box temperature warms under the positive forcing; a strided subsample
    matches the oracle to 1e-9."""
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    M = 100_000
    b, binds, params, scen = syn.config4(M=M)
    params = params.copy()
    params[M // 2:] = params[:M // 2]          # second half repeats the first
    ens = b.build_ensemble().bind_parameters(binds)
    ens.select_outputs(["Surface Temperature", "Heat Uptake"], t_start=0, t_stop=351, t_step=25)
    sc = torch.from_numpy(ens.pack_scenarios(scen)).cuda()
    p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    out = torch.empty((ens.output_rows, M), dtype=torch.float64, device="cuda")
    status = torch.zeros(M, dtype=torch.uint8, device="cuda")
    ens.run_device(p, sc, out, status, layout=0)
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    assert torch.allclose(out[:, : M // 2], out[:, M // 2:], rtol=0.0, atol=0.0, equal_nan=True)   # NaN at index 0 of pure outputs
    host = ens.split_outputs(out.cpu().numpy())
    w = np.array([0.5 * 0.58, 0.5 * 0.42, 0.5 * 0.79, 0.5 * 0.21])
    t_global = np.einsum("trm,r->tm", host["Surface Temperature"], w)
    assert np.all(t_global[-1] > t_global[4]) and np.all(t_global[-1] > 0.5) and np.all(t_global[-1] < 12.0)
    idx = np.arange(0, M, 997)
    m = oracle_from_builder(b)
    ref = m.split(m.run_batch(oracle_bindings(b, binds), params[idx], ens.exogenous_names, sc.cpu().numpy(),
                              ["Surface Temperature", "Heat Uptake"]), ["Surface Temperature", "Heat Uptake"])
    sel = slice(0, 351, 25)
    assert rel_err(host["Surface Temperature"][..., idx], ref["Surface Temperature"][sel]) <= TOL64
    assert rel_err(host["Heat Uptake"][..., idx], ref["Heat Uptake"][sel]) <= TOL64


def test_full_size_config5_log_posterior_equals_the_formula_on_the_run_output():
    """1M-member log-posterior (BASELINE config 5): for a subsample, the fused kernel's value equals the Gaussian
    log-likelihood (likelihood.rs:209-221) evaluated in numpy on the same members' run output plus the uniform log-prior;
    the device summary (max, argmax, finite count) matches a host reduction of the full vector."""
    torch = pytest.importorskip("torch")
    M = 1 << 20
    b, binds, params, scen = syn.config2(M=M)
    ens = b.build_ensemble().bind_parameters(binds)
    sc_host = ens.pack_scenarios(scen)
    ens.select_outputs(["Surface Temperature"])
    truth = np.array([[1.1, 0.05, 1.3, 0.7, 8.0, 100.0]])
    t_true = ens.run(truth, sc_host)[:, 0]
    obs = syn.config5_observations(t_true, syn.time_axis().values())
    priors = [(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.TWO_LAYER_RANGES.values()]
    ens.set_target(obs).set_priors(priors)
    d_p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    d_s = torch.from_numpy(sc_host).cuda()
    d_lp = torch.empty(M, dtype=torch.float64, device="cuda")
    d_sum = torch.zeros(5, dtype=torch.float64, device="cuda")
    ens.log_posterior_device(d_p, d_s, d_lp, d_sum, layout=0)
    torch.cuda.synchronize()
    lp = d_lp.cpu().numpy()
    idx = np.arange(0, M, 4099)
    series = ens.run(params[idx], sc_host)                                  # [351, n]
    years = syn.time_axis().values()
    ti = np.array([int(np.where(years == y)[0][0]) for _, y, _, _ in obs])
    val = np.array([v for _, _, v, _ in obs])[:, None]
    sig = np.array([s for _, _, _, s in obs])[:, None]
    loglik = np.sum(-0.5 * ((series[ti] - val) / sig) ** 2, axis=0)
    logprior = sum(-np.log(hi - lo) for lo, hi in syn.TWO_LAYER_RANGES.values())
    want = loglik + logprior
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(lp[idx]), fin) and np.max(np.abs(lp[idx][fin] - want[fin]) / np.abs(want[fin])) <= 1e-9
    summ = np.frombuffer(d_sum.cpu().numpy().tobytes(), dtype=np.dtype([("max", "f8"), ("argmax", "i8"), ("sum", "f8"), ("nfin", "i8"), ("n", "i8")]))[0]
    finite = np.isfinite(lp)
    assert summ["n"] == M and summ["nfin"] == finite.sum() and summ["max"] == lp[finite].max() and lp[summ["argmax"]] == summ["max"]
    assert abs(summ["sum"] - lp[finite].sum()) <= 1e-9 * abs(lp[finite].sum())


@pytest.mark.parametrize("step_size", [1.0 / 120.0, 0.05, 0.1, 0.25, 1.0])
def test_carbon_cycle_collapsed_substeps_match_stepping_over_wide_ranges(step_size):
    """The device composes the carbon cycle's n RK4 sub-steps analytically (components.cuh); the oracle steps them one by
    one as the reference does.  Wide ranges: lifetimes from 2 to 200 years, warm and cold feedback (negative alpha*T
    included), sub-step sizes from 1/120 to 1 year (z = h/lifetime up to ~0.8), emissions with sign changes, non-annual
    steps (odd sub-step counts)."""
    values = np.concatenate([np.arange(1750.0, 1900.0), np.arange(1900.0, 2101.0, 3.0)])
    axis = TimeAxis.from_values(values)
    b = (ModelBuilder().with_time_axis(axis)
         .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": 20.3, "conc_pi": 280.0, "alpha_temperature": 0.05})
                              .with_solver_options(step_size).build())
         .with_initial_values({"Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": 280.0}))
    t = values - 1750.0
    scen = [{"Emissions|CO2|Anthropogenic": 0.03 * t * np.cos(t / 40.0) * f, "Surface Temperature": 0.012 * t * f - 1.0} for f in (1.0, 2.5)]
    binds = {"tau": "CarbonCycle.tau", "alpha": "CarbonCycle.alpha_temperature", "c0": "initial:Atmospheric Concentration|CO2"}
    params = syn.uniform_params({"tau": (2.0, 200.0), "alpha": (-0.2, 0.4), "c0": (200.0, 500.0)}, 192, 17)
    got, _, worst, _ = gpu_vs_oracle(b, binds, params, scen)
    assert worst <= 1e-11 and np.isfinite(got["Atmospheric Concentration|CO2"]).all()
