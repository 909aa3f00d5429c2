"""ClimateUDEB (MAGICC upwelling-diffusion energy-balance model): the oracle against the MAGICC7 golden
vectors the reference's regression suite holds (tests/regression/test_ocean_udeb.py, phased 1-5 % tolerances),
LAMCALC known answers, and GPU parity against the oracle (1e-9)."""

import json
import os

import numpy as np
import pytest

from rscm_b200 import synthetic as syn
from rscm_b200.core import GridType, InterpolationStrategy, ModelBuilder, TimeAxis, Timeseries, VariableSchema
from rscm_b200.magicc import ClimateUDEBBuilder

from .helpers import oracle_bindings, oracle_from_builder, rel_err

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "ocean_udeb.npz"))
AREA_W = np.array([0.5 * 0.58, 0.5 * 0.42, 0.5 * 0.79, 0.5 * 0.21])  # tests/regression/helpers.py:94-102


def udeb_params(config):
    # build_ocean_model — tests/regression/test_ocean_udeb.py:60-110
    return {
        "ecs": config.get("core_climatesensitivity", 3.0),
        "rf_2xco2": config.get("core_delq2xco2", 3.71),
        "w_initial": config.get("core_initial_upwelling_rate", 3.5),
        "w_variable_fraction": config.get("core_upwelling_variable_part", 0.7),
        "depth_dependent_area": float(config.get("core_ocn_depthdependent", 1)),
        "kappa_dkdt": config.get("core_verticaldiff_top_dkdt", -0.191),
        "land_heat_capacity_enabled": float(bool(config.get("core_landheatcapacity_apply", 1))),
        "land_hc_eff_thickness": config.get("core_landhc_effthickness", 300.0),
        "k_lg": config.get("core_heatxchange_landground", 0.1),
        "k_ns": config.get("core_heatxchange_northsouth", 0.31),
        "feedback_cumt_sensitivity": config.get("core_feedback_cumtsensitivity", 0.08),
        "feedback_q_sensitivity": config.get("core_feedback_qsensitivity", 7.84e-9),
        "efficacy_apply": config.get("rf_efficacy_apply", 0),
        "prescribed_efficacy_co2": config.get("rf_efficacy_co2", 1.0),
    }


def udeb_builder(params, years, erf=None):
    axis = TimeAxis.from_bounds(np.concatenate([years, [years[-1] + 1.0]]))
    schema = VariableSchema()
    schema.add_variable("Effective Radiative Forcing", "W/m^2")
    schema.add_variable("Surface Temperature", "K", GridType.FourBox)
    schema.add_variable("Heat Uptake", "W/m^2")
    schema.add_variable("Ocean Heat Content", "J/m^2")
    schema.add_variable("Sea Surface Temperature", "K")
    b = (ModelBuilder().with_time_axis(axis).with_schema(schema)
         .with_rust_component(ClimateUDEBBuilder.from_parameters(params).build())
         .with_initial_values({"Surface Temperature": 0.0}))
    if erf is not None:
        b.with_exogenous_variable("Effective Radiative Forcing", Timeseries(erf, axis, "W/m^2", InterpolationStrategy.Linear))
    return b


def phased(actual, expected, skip=5, shock_end=25, converge_start=55, shock_rtol=3e-2, converge_rtol=2e-2, final_rtol=2e-2,
           final_years=20, atol=1e-6):
    # compute_phased_metrics — tests/regression/helpers.py:175-270
    n = len(actual)
    rel = np.where(np.abs(expected) > atol, (actual - expected) / np.where(expected == 0, 1, expected), 0.0)
    f_start = max(skip, n - final_years)
    res = {}
    for label, a, b, tol in (("shock", skip, min(shock_end, n), shock_rtol), ("transition", min(shock_end, n), min(converge_start, n), shock_rtol),
                             ("converge", min(converge_start, n), f_start, converge_rtol), ("final", f_start, n, final_rtol)):
        if a < b:
            res[label] = (float(np.max(np.abs(rel[a:b]))), tol)
    return res


SCENARIOS = {  # tolerances exactly as in tests/regression/test_ocean_udeb.py:222-488
    "01_diffusion_only": dict(shock_rtol=1.5e-2, converge_rtol=1.5e-2, final_rtol=1.5e-2),
    "02_constant_upwelling": dict(shock_rtol=1.5e-2, converge_rtol=1.5e-2, final_rtol=1.5e-2),
    "03_depth_dependent_area": dict(final_rtol=1e-2),
    "04_variable_upwelling": dict(),
    "05_temp_dependent_diffusivity": dict(converge_rtol=1.5e-2, final_rtol=1.5e-2),
    "06_ground_heat": dict(shock_rtol=5e-2, skip=15, final_rtol=1.5e-2),
    "07_interhemispheric_exchange": dict(shock_rtol=1.5e-2, converge_rtol=1.5e-2, final_rtol=1.5e-2),
    "09_time_varying_ecs": dict(final_rtol=1e-2),
    "11_efficacy_ar6": dict(final_rtol=1e-2),
}


# scenarios 08, 10, 12 use the full default configuration and a single tolerance (assert_allclose_recorded, rtol 0.1,
# atol 1e-6): tests/regression/test_ocean_udeb.py:316-562.  08 is the ABRUPT-2XCO2 step over ten years, 10 and 12 ramp
# the forcing along the 1pctCO2 pathway; 12 adds efficacy_apply = 2.
FULL_DEFAULT_SCENARIOS = {"08_sst_to_sat": "step", "10_full_default": "1pctco2", "12_efficacy_ar6_1pctco2": "1pctco2"}


def full_default_case(name):
    years, expected = GOLDEN[name + "/years"], GOLDEN[name + "/temp"]
    config = json.loads(str(GOLDEN[name + "/config"]))
    rf = config.get("core_delq2xco2", 3.71)
    if FULL_DEFAULT_SCENARIOS[name] == "step":
        erf = np.where(years >= 1851.0, rf, 0.0)   # construct_step_forcing, test_ocean_udeb.py:113-130
    else:
        dt = years - config.get("startyear", 1850)
        erf = rf * np.log(np.where(dt > 0, 1.01 ** dt, 1.0)) / np.log(2.0)
    params = {"ecs": config.get("core_climatesensitivity", 3.0), "rf_2xco2": rf}
    if name.startswith("12"):
        params["efficacy_apply"] = config.get("rf_efficacy_apply", 2)
    return years, erf, params, expected


def recorded_rel_err(actual, expected, atol=1e-6):
    # assert_allclose_recorded — tests/regression/helpers.py:338-365
    return float(np.max(np.abs(np.where(np.abs(expected) > atol, (actual - expected) / np.where(expected == 0, 1, expected), 0.0))))


@pytest.mark.parametrize("name", sorted(FULL_DEFAULT_SCENARIOS))
def test_oracle_matches_magicc7_golden_full_default(name):
    years, erf, params, expected = full_default_case(name)
    actual = oracle_from_builder(udeb_builder(params, years, erf)).run()["Surface Temperature"] @ AREA_W
    np.testing.assert_allclose(actual, expected, rtol=0.1, atol=1e-6, err_msg=name)
    print(name, "max relative error vs MAGICC7: %.3f" % recorded_rel_err(actual, expected))


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_matches_magicc7_golden(name):
    years, expected = GOLDEN[name + "/years"], GOLDEN[name + "/temp"]
    config = json.loads(str(GOLDEN[name + "/config"]))
    erf = np.where(years >= 1851.0, config.get("core_delq2xco2", 3.71), 0.0)
    res = oracle_from_builder(udeb_builder(udeb_params(config), years, erf)).run()
    actual = res["Surface Temperature"] @ AREA_W
    for phase, (err, tol) in phased(actual, expected, **SCENARIOS[name]).items():
        assert err <= tol, f"{name} {phase}: {err:.4f} > {tol}"


def test_udeb_outputs_and_state_semantics():
    years = np.arange(1850.0, 1901.0)
    erf = np.where(years >= 1851.0, 3.71, 0.0)
    r = oracle_from_builder(udeb_builder({}, years, erf)).run()
    st = r["Surface Temperature"]
    assert st.shape == (51, 4) and np.all(st[0] == 0.0)           # state: scalar initial value broadcast to 4 boxes
    assert np.isnan(r["Heat Uptake"][0]) and np.isnan(r["Sea Surface Temperature"][0])  # pure outputs NaN at index 0
    assert np.all(st[1] > 0.0)     # forcing is interpolated start->end inside the year (erf_end = ERF[N+1] = 3.71)
    assert np.all(np.diff(st[2:, 0]) > 0) and st[-1, 1] > st[-1, 0]  # warming; land warms more than ocean
    assert r["Heat Uptake"][-1] > 0 and r["Ocean Heat Content"][-1] > 0


# ---- GPU parity --------------------------------------------------------------------------------------------
UDEB_BINDS = {"ecs": "ClimateUDEB.ecs", "kappa": "ClimateUDEB.kappa", "rlo": "ClimateUDEB.rlo", "k_ns": "ClimateUDEB.k_ns"}


def _udeb_params(M, seed=5):
    return syn.uniform_params({"ecs": (1.5, 4.5), "kappa": (0.5, 1.5), "rlo": (1.1, 1.5), "k_ns": (0.1, 0.5)}, M, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-9)])
def test_udeb_gpu_parity(dtype, tol, tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    years = np.arange(1850.0, 1951.0)
    b = udeb_builder({}, years)
    ens = b.build_ensemble(dtype=dtype).bind_parameters(UDEB_BINDS)
    ramp = 3.71 * np.log2(np.exp(0.006 * (years - 1850.0)))
    scen = [{"Effective Radiative Forcing": np.where(years >= 1851.0, 3.71, 0.0)}, {"Effective Radiative Forcing": ramp}]
    sc = ens.pack_scenarios(scen)
    p = _udeb_params(96)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, UDEB_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= tol, n
    assert got["Surface Temperature"].shape == (101, 4, 192)


@pytest.mark.gpu
def test_config4_magicc_boxes_parity_and_fp32(tmp_path, monkeypatch):
    """BASELINE config 4 graph (forcing boxes -> Sum aggregate -> ClimateUDEB on the four-box grid) at a size the
    oracle finishes in seconds: fp64 parity 1e-9 on every series; fp32 path measured against the 1e-4 bar."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b, binds, params, scen = syn.config4(M=160)
    ens = b.build_ensemble().bind_parameters(binds)
    sc = ens.pack_scenarios(scen)
    got = ens.split_outputs(ens.run(params, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, binds), params, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    st = got["Surface Temperature"]
    assert st.shape == (351, 4, 160) and np.all(st[0] == 0.0) and np.isfinite(st).all()
    assert ens.execution_order()[-1] == 4  # ClimateUDEB runs after the ERF aggregator (node 5)
    # fp32 path
    e32 = b.build_ensemble(dtype="f32").bind_parameters(binds)
    g32 = e32.split_outputs(e32.run(params, sc))
    errs = {n: rel_err(g32[n], ref[n]) for n in syn.CONFIG4_OUTPUTS}
    print("config4 fp32 relative errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["Effective Radiative Forcing"] <= 1e-4
    assert errs["Surface Temperature"] <= 5e-3  # documented divergence: LAMCALC's 1e-3 tolerance iteration + 50-layer Thomas in fp32


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_gpu_matches_magicc7_golden(name, tmp_path, monkeypatch):
    """The same golden comparison as test_oracle_matches_magicc7_golden, but with the fused kernel producing the series
    (single member through the reference-shaped Model API): the GPU path is pinned to MAGICC7 directly, not only via the oracle."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    years, expected = GOLDEN[name + "/years"], GOLDEN[name + "/temp"]
    config = json.loads(str(GOLDEN[name + "/config"]))
    erf = np.where(years >= 1851.0, config.get("core_delq2xco2", 3.71), 0.0)
    model = udeb_builder(udeb_params(config), years, erf).build()
    model.run()
    t4 = np.asarray(model.timeseries().get_fourbox_timeseries_by_name("Surface Temperature").values())
    actual = t4 @ AREA_W
    for phase, (err, tol) in phased(actual, expected, **SCENARIOS[name]).items():
        assert err <= tol, f"{name} {phase}: {err:.4f} > {tol}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FULL_DEFAULT_SCENARIOS))
def test_gpu_matches_magicc7_golden_full_default(name, tmp_path, monkeypatch):
    """Scenarios 08 / 10 / 12 (full default MAGICC7 configuration) with the fused kernel producing the series."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    years, erf, params, expected = full_default_case(name)
    model = udeb_builder(params, years, erf).build()
    model.run()
    t4 = np.asarray(model.timeseries().get_fourbox_timeseries_by_name("Surface Temperature").values())
    np.testing.assert_allclose(t4 @ AREA_W, expected, rtol=0.1, atol=1e-6, err_msg=name)


@pytest.mark.gpu
@pytest.mark.parametrize("M", [1, 7, 33, 130])
def test_udeb_log_posterior_and_ragged_member_counts(M, tmp_path, monkeypatch):
    """Four lanes work on one member (climate_udeb.cuh): the log-posterior variants (padding lanes stay for the block
    reduction), the summary and ragged member counts — 1 member = one lane quad of a warp, 33 = one quad into a second CTA."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    from rscm_b200 import _ffi
    years = np.arange(1850.0, 1921.0)
    b = udeb_builder({}, years)
    ens = b.build_ensemble().bind_parameters(UDEB_BINDS)
    ramp = 3.71 * np.log2(np.exp(0.006 * (years - 1850.0)))
    sc = ens.pack_scenarios([{"Effective Radiative Forcing": np.where(years >= 1851.0, 3.71, 0.0)}, {"Effective Radiative Forcing": ramp}])
    p = _udeb_params(M, seed=9)
    m = oracle_from_builder(b)
    ob = oracle_bindings(b, UDEB_BINDS)
    names = ["Sea Surface Temperature", "Heat Uptake", "Surface Temperature"]
    ens.select_outputs(names)
    got = ens.split_outputs(ens.run(p, sc))
    ref = m.split(m.run_batch(ob, p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    obs = [("Sea Surface Temperature", float(y), 0.8, 0.3) for y in range(1860, 1921, 10)]
    priors = [(_ffi.PRIOR_UNIFORM, 1.0, 5.0), (_ffi.PRIOR_NORMAL, 1.0, 0.5), (_ffi.PRIOR_UNIFORM, 1.0, 1.6), (_ffi.PRIOR_UNIFORM, 0.0, 0.45)]
    ens.set_target(obs).set_priors(priors)
    lp, summ = ens.log_posterior(p, sc, with_summary=True)
    want = m.log_posterior_batch(ob, p, ens.exogenous_names, sc, priors, obs)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(lp), fin)
    assert np.max(np.abs(lp[fin] - want[fin]) / np.abs(want[fin])) <= 1e-9 if fin.any() else True
    assert summ["n_runs"] == 2 * M and summ["n_finite"] == int(fin.sum())
    if fin.any():
        assert summ["argmax"] == int(np.argmax(np.where(fin, lp, -np.inf))) and summ["max_logpost"] == lp[summ["argmax"]]
