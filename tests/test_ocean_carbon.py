"""OceanCarbon (impulse-response ocean carbon cycle with an O(T^2) flux-history convolution) and the full
emissions-driven MAGICC graph of the reference's regression suite (tests/regression/test_ghg_forcing.py:470-620,
minus HalocarbonChemistry): oracle known answers on CPU, GPU parity at 1e-9."""

import numpy as np
import pytest

from oracle import oracle as orc
from rscm_b200 import synthetic as syn
from rscm_b200.core import ModelBuilder
from rscm_b200.magicc import OceanCarbonBuilder

from .helpers import oracle_bindings, oracle_from_builder, rel_err


def ocean_builder(model="3D-GFDL", start=1850, end=1950):
    return (ModelBuilder().with_time_axis(syn.time_axis(start, end))
            .with_rust_component(OceanCarbonBuilder.from_parameters({"model": model}).build())
            .with_initial_values({"Ocean Surface pCO2": 278.0, "Cumulative Ocean Uptake": 0.0}))


def ocean_scenario(start=1850, end=1950, scale=1.0):
    years = syn.time_axis(start, end).values()
    return {"Atmospheric Concentration|CO2": 278.0 * np.exp(0.004 * scale * (years - start)), "Sea Surface Temperature": 0.01 * (years - start)}


# ---- oracle known answers: crates/rscm-magicc/src/parameters/ocean_carbon.rs and carbon/ocean.rs unit tests ----
def _irf(model, t):
    import ctypes
    comp = OceanCarbonBuilder.from_parameters({"model": model}).build()
    p = np.array(comp.params)
    L = orc.lib()
    L.orc_ocean_irf.restype = ctypes.c_double
    L.orc_ocean_irf.argtypes = [ctypes.c_void_p, ctypes.c_double]
    return L.orc_ocean_irf(orc._dp(p), t)


def test_irf_presets_behave_like_the_reference_tests():
    assert _irf("HILDA", 0.0) > 0.9 and _irf("HILDA", 100.0) < _irf("HILDA", 10.0) < _irf("HILDA", 0.0) and _irf("HILDA", 100.0) > 0.0
    for model, sw in (("3D-GFDL", 1.0), ("2D-BERN", 9.9), ("HILDA", 2.0)):
        before, after = _irf(model, sw - 1e-6), _irf(model, sw + 1e-6)
        assert 0.0 < before < 1.5 and 0.0 < after < 1.5 and 0.1 < before / after < 10.0
    assert _irf("3D-GFDL", 0.0) == pytest.approx(1.0)  # polynomial IRF starts at 1; scaling keeps 1 -> 1


@pytest.mark.parametrize("model", ["3D-GFDL", "2D-BERN", "HILDA"])
def test_ocean_uptake_follows_atmospheric_co2(model):
    b = ocean_builder(model, end=1900)
    r = oracle_from_builder(b, ocean_scenario(end=1900)).run()
    flux, pco2, cum = r["Carbon Flux|Ocean"], r["Ocean Surface pCO2"], r["Cumulative Ocean Uptake"]
    assert np.isnan(flux[0]) and pco2[0] == 278.0 and cum[0] == 0.0
    assert flux[1] == 0.0                        # equilibrium in the first year: atm = ocean pCO2 (test_zero_flux_at_equilibrium)
    assert np.all(flux[3:] > 0.0) and np.all(np.diff(cum[2:]) > 0.0)   # atm > ocean drives uptake
    assert np.all(pco2[3:] > 278.0) and np.all(pco2[3:] < r["Atmospheric Concentration|CO2"][3:])
    np.testing.assert_allclose(np.diff(cum)[1:], flux[2:], rtol=1e-12)  # cumulative integrates the annual-mean flux (dt = 1)


def test_equilibrium_stays_at_rest():
    b = ocean_builder(end=1870)
    r = oracle_from_builder(b, {"Atmospheric Concentration|CO2": np.full(21, 278.0), "Sea Surface Temperature": np.zeros(21)}).run()
    assert np.all(r["Carbon Flux|Ocean"][1:] == 0.0) and np.all(r["Ocean Surface pCO2"] == 278.0)


def test_history_limit_truncates_on_the_oracle():
    """OceanCarbonState::flux_history is a deque bounded by max_history_months (ocean.rs:224-228): once it is full the
    oldest flux leaves the convolution.  With a 10-year bound the run differs from the unbounded one exactly from the year
    the bound starts to bite."""
    def run(max_hist):
        b = (ModelBuilder().with_time_axis(syn.time_axis(1850, 1900))
             .with_rust_component(OceanCarbonBuilder.from_parameters({"max_history_months": max_hist}).build())
             .with_initial_values({"Ocean Surface pCO2": 278.0, "Cumulative Ocean Uptake": 0.0}))
        sc = {"Atmospheric Concentration|CO2": 278.0 + 2.0 * np.arange(51.0), "Sea Surface Temperature": 0.01 * np.arange(51.0)}
        return oracle_from_builder(b, sc).run()["Ocean Surface pCO2"]
    full, cut = run(6000), run(120)
    assert np.array_equal(full[:11], cut[:11]) and not np.array_equal(full[11:], cut[11:])   # 120 months = 10 years
    assert np.all(np.isfinite(cut))


# ---- GPU parity ---------------------------------------------------------------------------------------------------
OCEAN_BINDS = {"tau": "OceanCarbon.gas_exchange_tau", "sst_pi": "OceanCarbon.sst_pi", "ts": "OceanCarbon.temp_sensitivity",
               "mld": "OceanCarbon.mixed_layer_depth"}


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["3D-GFDL", "HILDA"])
def test_ocean_carbon_gpu_parity(model, tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = ocean_builder(model, end=1950)
    ens = b.build_ensemble().bind_parameters(OCEAN_BINDS)
    sc = ens.pack_scenarios([ocean_scenario(), ocean_scenario(scale=1.5)])
    p = syn.uniform_params({"tau": (6.0, 10.0), "sst_pi": (17.0, 19.0), "ts": (0.03, 0.045), "mld": (45.0, 80.0)}, 96, 31)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, OCEAN_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n


# the emissions-driven chain itself lives in rscm_b200/synthetic.py (bench.py times it as BASELINE config 4's widest graph)
full_magicc_builder, full_magicc_scenario, FULL_BINDS = syn.full_chain_builder, syn.full_chain_scenario, syn.FULL_CHAIN_BINDINGS


def test_full_magicc_graph_on_the_oracle():
    b = full_magicc_builder(end=1900)
    m = oracle_from_builder(b, full_magicc_scenario(end=1900))
    ens = b.build_ensemble(device=-2)
    assert ens.execution_order() == m.execution_order() and ens.variable_names == m.names and ens.program_is_jit()
    r = m.run()
    assert r["Atmospheric Concentration|CO2"][-1] > 285.0 and r["Atmospheric Concentration|CH4"][-1] > 900.0
    assert np.all(np.isfinite(r["Surface Temperature"])) and abs(r["Surface Temperature"][-1].mean()) > 0.01
    assert r["Cumulative Ocean Uptake"][-1] > 0.0


@pytest.mark.gpu
def test_full_magicc_gpu_parity(tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = full_magicc_builder()
    ens = b.build_ensemble().bind_parameters(FULL_BINDS)
    sc = ens.pack_scenarios([full_magicc_scenario(), full_magicc_scenario(f=1.3)])
    p = syn.uniform_params({"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}, 64, 41)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, FULL_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    assert np.isfinite(got["Surface Temperature"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("max_hist", [120, 30, 7])
def test_ocean_carbon_history_truncation_gpu_parity(max_hist, tmp_path, monkeypatch):
    """History truncation on the device (zero IRF weights from lag max_history_months on, older fluxes not loaded) against
    the oracle's bounded history, for bounds above and below one year of months."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = (ModelBuilder().with_time_axis(syn.time_axis(1850, 1930))
         .with_rust_component(OceanCarbonBuilder.from_parameters({"max_history_months": max_hist}).build())
         .with_initial_values({"Ocean Surface pCO2": 278.0, "Cumulative Ocean Uptake": 0.0}))
    ens = b.build_ensemble().bind_parameters(OCEAN_BINDS)
    years = np.arange(81.0)
    sc = ens.pack_scenarios([{"Atmospheric Concentration|CO2": 278.0 + 1.5 * years + 0.01 * years ** 2, "Sea Surface Temperature": 0.012 * years}])
    p = syn.uniform_params({"tau": (6.0, 10.0), "sst_pi": (17.0, 19.0), "ts": (0.03, 0.045), "mld": (45.0, 80.0)}, 40, 33)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, OCEAN_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n


@pytest.mark.gpu
def test_full_magicc_truncated_history_in_lane_quads(tmp_path, monkeypatch):
    """The emissions-driven chain runs as lane quads (ClimateUDEB): OceanCarbon takes years in blocks of four with the
    history before the block summed by the four lanes — here with a history bound that cuts inside the blocks."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    from rscm_b200.magicc import OceanCarbonBuilder as OCB
    b = full_magicc_builder(end=1950)
    for i, c in enumerate(b._components):
        if c.type_name == "OceanCarbon":
            b._components[i] = OCB.from_parameters({"max_history_months": 250}).build()
    ens = b.build_ensemble().bind_parameters(FULL_BINDS)
    sc = ens.pack_scenarios([full_magicc_scenario()])
    p = syn.uniform_params({"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}, 37, 45)
    names = ["Atmospheric Concentration|CO2", "Carbon Flux|Ocean", "Ocean Surface pCO2", "Surface Temperature"]
    ens.select_outputs(names)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    ref = m.split(m.run_batch(oracle_bindings(b, FULL_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n


@pytest.mark.gpu
def test_eleven_box_chain_ragged_members_and_log_posterior(tmp_path, monkeypatch):
    """The widest graph (HalocarbonChemistry included: 124 variables, cells swapped by pointer, OceanCarbon's history staged by
    bulk copies) with a member count that leaves padding lanes in the second CTA, two scenarios, 150 years (37 blocks of four
    years, up to 113 staged tiles per block), and the log-posterior variant of the same program."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    from rscm_b200 import _ffi
    b, binds, p, scen = syn.full_chain(M=45, end=2000)
    scen2 = dict(scen[0])
    scen2["Emissions|CO2|Fossil"] = scen2["Emissions|CO2|Fossil"] * 1.4
    scen2["Emissions|CFC-12"] = scen2["Emissions|CFC-12"] * 0.5
    ens = b.build_ensemble().bind_parameters(binds)
    assert ens.program_is_jit() and "SWAP_CELLS = true" in ens.program_signature() and "LANES = 4" in ens.program_signature()
    sc = ens.pack_scenarios([scen[0], scen2])
    names = syn.FULL_CHAIN_OUTPUTS + ["Ocean Surface pCO2", "Atmospheric Concentration|CFC-12", "Forcing|Halocarbons", "Ocean Heat Content"]
    ens.select_outputs(names)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    ob = oracle_bindings(b, binds)
    ref = m.split(m.run_batch(ob, p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    obs = [("Atmospheric Concentration|CO2", float(y), 300.0 + 0.5 * (y - 1900), 8.0) for y in range(1900, 2001, 20)] + \
          [("Sea Surface Temperature", float(y), -0.2, 0.3) for y in range(1950, 2001, 25)]
    priors = [(_ffi.PRIOR_UNIFORM, 1.5, 5.0), (_ffi.PRIOR_UNIFORM, 0.3, 1.0), (_ffi.PRIOR_NORMAL, 8.0, 2.0), (_ffi.PRIOR_UNIFORM, 8.0, 11.0)]
    ens.set_target(obs).set_priors(priors)
    lp, summ = ens.log_posterior(p, sc, with_summary=True)
    want = m.log_posterior_batch(ob, p, ens.exogenous_names, sc, priors, obs)
    fin = np.isfinite(want)
    assert fin.any() and np.array_equal(np.isfinite(lp), fin)
    assert np.max(np.abs(lp[fin] - want[fin]) / np.abs(want[fin])) <= 1e-9
    assert summ["n_runs"] == 90 and summ["n_finite"] == int(fin.sum())


def test_lane_group_programs_are_sized_for_four_ctas_per_sm():
    """graph.cpp picks the CTAs per SM a lane-group program is compiled for from its shared-memory footprint: four where four
    CTAs fit the SM's 227 KB (then the exogenous rows stay in global memory), else three."""
    import re
    b4, binds4, _, _ = syn.config4(M=4)
    e4 = b4.build_ensemble(device=-2).bind_parameters(binds4)
    sig4 = e4.program_signature()
    assert "LANES = 4" in sig4 and "MIN_BLOCKS = 4" in sig4 and "STAGE_EXO = false" in sig4
    assert 4 * (e4.shared_bytes() + 1024) <= 232448
    bf, bindsf, _, _ = syn.full_chain(M=4)
    ef = bf.build_ensemble(device=-2).bind_parameters(bindsf)
    sigf = ef.program_signature()
    assert "MIN_BLOCKS = 4" in sigf and "SWAP_CELLS = true" in sigf and 4 * (ef.shared_bytes() + 1024) <= 232448
    assert ef.shared_bytes(log_posterior=True) >= ef.shared_bytes()
    # thread-per-member programs keep their exogenous rows in shared memory
    sig2 = syn.config2(M=4)[0].build_ensemble(device=-2).program_signature()
    assert "LANES = 1" in sig2 and "STAGE_EXO = true" in sig2 and int(re.search(r"MIN_BLOCKS = (\d+)", sig2).group(1)) == 8


@pytest.mark.gpu
@pytest.mark.parametrize("per_member_initial", [False, True])
def test_eleven_box_chain_halocarbon_species_shared_among_lanes(per_member_initial, tmp_path, monkeypatch):
    """In a lane-group program whose halocarbon emissions are exogenous, the 41 species are shared among the 32 lanes of a warp
    (all members of a scenario compute the same concentrations) — unless an initial concentration is bound per member, which
    the kernel finds out at the first step and then runs member by member.  Both against the oracle, with a NaN emission in
    one scenario (latest_value fall-back) and a member count that leaves padding lanes."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b, binds, p, scen = syn.full_chain(M=41, end=1920)
    binds = dict(binds)
    if per_member_initial:
        binds["cfc11"] = "initial:Atmospheric Concentration|CFC-11"
        p = np.concatenate([p, syn.uniform_params({"cfc11": (0.0, 300.0)}, 41, 5)], axis=1)
    scen2 = {k: np.array(v, dtype=float) for k, v in scen[0].items()}
    scen2["Emissions|CFC-12"][30] = np.nan
    scen2["Emissions|SF6"] *= 3.0
    ens = b.build_ensemble().bind_parameters(binds)
    sc = ens.pack_scenarios([scen[0], scen2])
    names = ["EESC", "Forcing|Halocarbons", "Forcing|F-gases", "Forcing|Montreal Gases", "Atmospheric Concentration|CFC-11",
             "Atmospheric Concentration|CFC-12", "Atmospheric Concentration|SF6", "Atmospheric Concentration|CH3Cl", "Effective Radiative Forcing",
             "Surface Temperature"]
    ens.select_outputs(names)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    ref = m.split(m.run_batch(oracle_bindings(b, binds), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    if per_member_initial:
        assert np.ptp(got["Atmospheric Concentration|CFC-11"][0]) > 100.0      # the bound initial values differ by member


@pytest.mark.gpu
def test_eleven_box_chain_fp32_path_within_1e4(tmp_path, monkeypatch):
    """The optional fp32 path (bar: 1e-4) on the widest graph over a century: the history tiles, prefix sums and exchange slots
    stay fp64, the arithmetic is fp32."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b, binds, p, scen = syn.full_chain(M=70, end=1950)
    names = syn.FULL_CHAIN_OUTPUTS + ["Ocean Surface pCO2", "Ocean Heat Content"]
    ens = b.build_ensemble(dtype="f32").bind_parameters(binds)
    ens.select_outputs(names)
    sc = ens.pack_scenarios(scen)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    ref = m.split(m.run_batch(oracle_bindings(b, binds), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-4, n


@pytest.mark.gpu
@pytest.mark.parametrize("steps", [7, 16])
def test_eleven_box_chain_with_other_ocean_step_counts(steps, tmp_path, monkeypatch):
    """OceanCarbon's staged tiles and IRF windows with an odd number of months per year (the bulk copies of the table window
    start on even entries: the offset inside the window changes from block to block) and with the largest one."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    from rscm_b200.magicc import OceanCarbonBuilder as OCB
    b = full_magicc_builder(end=1930)
    for i, c in enumerate(b._components):
        if c.type_name == "OceanCarbon":
            b._components[i] = OCB.from_parameters({"steps_per_year": steps}).build()
    ens = b.build_ensemble().bind_parameters(FULL_BINDS)
    sc = ens.pack_scenarios([full_magicc_scenario(end=1930)])
    p = syn.uniform_params({"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}, 33, 47)
    names = ["Atmospheric Concentration|CO2", "Carbon Flux|Ocean", "Ocean Surface pCO2", "Cumulative Ocean Uptake", "Surface Temperature"]
    ens.select_outputs(names)
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    ref = m.split(m.run_batch(oracle_bindings(b, FULL_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
