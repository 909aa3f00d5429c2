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
