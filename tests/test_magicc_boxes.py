"""Scalar box components: FourBoxOceanHeatUptake, OceanSurfacePartialPressure (rscm-components) and CO2Budget,
TerrestrialCarbon, CH4Chemistry, N2OChemistry (rscm-magicc).  CPU: the oracle against the reference's unit-test
known answers.  GPU: parity on a coupled gas-cycle graph (chemistry + terrestrial + budget + GHG forcing + two-layer)
that exercises `previous()` / `at_offset()` history and lagged feedbacks."""

import math

import numpy as np
import pytest

from rscm_b200 import synthetic as syn
from rscm_b200.components import FourBoxOceanHeatUptakeBuilder, OceanSurfacePartialPressureBuilder
from rscm_b200.core import GridType, ModelBuilder, VariableSchema
from rscm_b200.magicc import (CH4ChemistryBuilder, CO2BudgetBuilder, GhgForcingBuilder, N2OChemistryBuilder, TerrestrialCarbonBuilder)
from rscm_b200.two_layer import TwoLayerBuilder

from .helpers import oracle_bindings, oracle_from_builder, rel_err


def _single(component, inputs, initial=None, T=3, start=2020):
    axis = syn.time_axis(start, start + T - 1)
    b = ModelBuilder().with_time_axis(axis).with_rust_component(component)
    if initial:
        b.with_initial_values(initial)
    sc = {k: np.full(T, v) if np.isscalar(v) else np.asarray(v, dtype=float) for k, v in inputs.items()}
    return oracle_from_builder(b, sc).run()


# ---- oracle known answers ---------------------------------------------------------------------------------
@pytest.mark.parametrize("params,expected", [
    (dict(ospp_preindustrial=278.0, sensitivity_ospp_to_temperature=0.043, delta_ospp_offsets=[1.5568, 7.4706, 1.2748, 2.4491, 1.5468],
          delta_ospp_coefficients=[-0.013993, -0.20207, -0.12015, -0.12639, -0.15326], sea_surface_temperature_preindustrial=17.9), 339.089),
    (dict(ospp_preindustrial=315.0, sensitivity_ospp_to_temperature=0.0423, delta_ospp_offsets=[1.5, 7.5, 1.3, 2.5, 1.6],
          delta_ospp_coefficients=[-0.02, -0.2, -0.1, -0.14, -0.2], sea_surface_temperature_preindustrial=17.9), 381.003),
])
def test_ocean_surface_partial_pressure_rstest_cases(params, expected):
    # ocean_surface_partial_pressure.rs:220-238 (SST anomaly 4.0, DIC anomaly 5.0, max_relative 10e-5)
    comp = OceanSurfacePartialPressureBuilder.from_parameters(params).build()
    r = _single(comp, {"Sea Surface Temperature": 4.0, "Dissolved Inorganic Carbon": 5.0})
    assert r["Ocean Surface Partial Pressure|CO2"][1] == pytest.approx(expected, rel=10e-5)


def test_four_box_ocean_heat_uptake():
    comp = FourBoxOceanHeatUptakeBuilder.from_parameters({}).build()
    r = _single(comp, {"Effective Radiative Forcing|Aggregated": 2.0})
    np.testing.assert_allclose(r["Heat Uptake|Ocean"][1], [2.4, 1.2, 3.2, 1.2])
    assert np.mean(r["Heat Uptake|Ocean"][1]) == pytest.approx(2.0)
    with pytest.raises(ValueError, match="average to 1.0"):
        FourBoxOceanHeatUptakeBuilder.from_parameters({k: 2.0 for k, _ in FourBoxOceanHeatUptakeBuilder.FIELDS}).build()


def test_co2_budget_mass_conservation():
    # carbon/budget.rs unit tests: 10 GtC/yr emitted, 4 taken up -> net 6, dCO2 = 6 / 2.123 ppm
    comp = CO2BudgetBuilder.from_parameters({}).build()
    r = _single(comp, {"Emissions|CO2|Fossil": 10.0, "Emissions|CO2|Land Use": 0.0, "Carbon Flux|Terrestrial": 2.0, "Carbon Flux|Ocean": 2.0},
                {"Atmospheric Concentration|CO2": 400.0})
    assert r["Emissions|CO2|Net"][1] == pytest.approx(6.0, abs=1e-10)
    assert r["Atmospheric Concentration|CO2"][1] - 400.0 == pytest.approx(6.0 / 2.123, abs=1e-10)
    assert r["Airborne Fraction|CO2"][1] == pytest.approx(0.6)
    r0 = _single(comp, {"Emissions|CO2|Fossil": 0.0, "Emissions|CO2|Land Use": 0.0, "Carbon Flux|Terrestrial": 1.0, "Carbon Flux|Ocean": 1.0},
                 {"Atmospheric Concentration|CO2": 400.0})
    assert r0["Airborne Fraction|CO2"][1] == 0.0


def test_terrestrial_steady_state_and_fertilization():
    comp = TerrestrialCarbonBuilder.from_parameters({}).build()
    pools = {"Carbon Pool|Plant": 884.86, "Carbon Pool|Detritus": 92.77, "Carbon Pool|Soil": 1681.53, "Carbon Pool|Humus": 836.0}
    r = _single(comp, {"Atmospheric Concentration|CO2": 278.0, "Surface Temperature": 0.0, "Emissions|CO2|Land Use": 0.0}, pools)
    for k, v in pools.items():
        assert abs(r[k][1] - v) / v < 0.05                      # terrestrial.rs test_steady_state_at_preindustrial
    assert abs(r["Carbon Flux|Terrestrial"][1]) < 1.0
    hi = _single(comp, {"Atmospheric Concentration|CO2": 556.0, "Surface Temperature": 0.0, "Emissions|CO2|Land Use": 0.0}, pools)
    assert hi["Carbon Flux|Terrestrial"][1] > r["Carbon Flux|Terrestrial"][1] + 5.0   # fertilisation: NPP x (1 + beta ln 2)


def test_ch4_and_n2o_steady_states():
    ch4 = CH4ChemistryBuilder.from_parameters({}).build()
    base = {"Emissions|CH4": 0.0, "Surface Temperature": 0.0, "Emissions|NOx": 0.0, "Emissions|CO": 0.0, "Emissions|NMVOC": 0.0}
    r = _single(ch4, base, {"Atmospheric Concentration|CH4": 722.0}, T=12)
    tau_other = 1.0 / (1.0 / 150.0 + 1.0 / 120.0 + 1.0 / 200.0)
    assert r["Lifetime|CH4"][1] == pytest.approx(1.0 / (1.0 / 9.3 + 1.0 / tau_other), rel=0.02)
    r_hi = _single(ch4, {**base, "Emissions|CH4": 300.0}, {"Atmospheric Concentration|CH4": 722.0}, T=12)
    assert r_hi["Atmospheric Concentration|CH4"][-1] > r["Atmospheric Concentration|CH4"][-1] + 100.0
    n2o = N2OChemistryBuilder.from_parameters({}).build()
    rn = _single(n2o, {"Emissions|N2O": 0.0}, {"Atmospheric Concentration|N2O": 270.0}, T=12)
    assert abs(rn["Lifetime|N2O"][1] - 139.275) / 139.275 < 0.01      # n2o.rs test_steady_state_at_preindustrial
    assert abs(rn["Atmospheric Concentration|N2O"][1] - 270.0) / 270.0 < 0.05
    rn5 = _single(n2o, {"Emissions|N2O": 5.0}, {"Atmospheric Concentration|N2O": 270.0}, T=12)
    # (the reference steps from previous() = index N-1, so the series advances in a two-step sawtooth)
    assert rn5["Atmospheric Concentration|N2O"][-1] > rn["Atmospheric Concentration|N2O"][-1] + 5.0


# ---- GPU parity on a coupled gas-cycle graph ------------------------------------------------------------------
def gas_cycle_builder(delay=2):
    schema = VariableSchema()
    for n in ("CH4", "N2O", "NOx", "CO", "NMVOC", "CO2|Fossil", "CO2|Land Use"):
        schema.add_variable(f"Emissions|{n}", "")
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "")
        schema.add_variable(f"Effective Radiative Forcing|{n}", "W/m^2")
    for n in ("Carbon Flux|Terrestrial", "Carbon Flux|Ocean", "Carbon Pool|Plant", "Carbon Pool|Detritus", "Carbon Pool|Soil", "Carbon Pool|Humus",
              "Emissions|CO2|Net", "Airborne Fraction|CO2", "Lifetime|CH4", "Lifetime|N2O", "Surface Temperature", "Deep Ocean Temperature"):
        schema.add_variable(n, "")
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", [f"Effective Radiative Forcing|{n}" for n in ("CO2", "CH4", "N2O")])
    return (
        ModelBuilder().with_time_axis(syn.time_axis(1850, 2100)).with_schema(schema)
        .with_rust_component(CH4ChemistryBuilder.from_parameters({}).build())
        .with_rust_component(N2OChemistryBuilder.from_parameters({"strat_delay": delay}).build())
        .with_rust_component(TerrestrialCarbonBuilder.from_parameters({}).build())
        .with_rust_component(CO2BudgetBuilder.from_parameters({}).build())
        .with_rust_component(GhgForcingBuilder.from_parameters({}).build())
        .with_rust_component(TwoLayerBuilder.from_parameters(syn.TWO_LAYER_DEFAULTS).build())
        .with_initial_values({"Atmospheric Concentration|CH4": 722.0, "Atmospheric Concentration|N2O": 270.0,
                              "Atmospheric Concentration|CO2": 278.0, "Carbon Pool|Plant": 884.86, "Carbon Pool|Detritus": 92.77,
                              "Carbon Pool|Soil": 1681.53, "Carbon Pool|Humus": 836.0, "Surface Temperature": 0.0,
                              "Deep Ocean Temperature": 0.0})
    )


def gas_cycle_scenarios(n=2):
    years = syn.time_axis(1850, 2100).values()
    ramp = (years - 1850.0) / 250.0
    return [{
        "Emissions|CH4": 50.0 + 350.0 * ramp * (1.0 + 0.2 * s), "Emissions|N2O": 1.0 + 9.0 * ramp, "Emissions|NOx": 5.0 + 35.0 * ramp,
        "Emissions|CO": 100.0 + 700.0 * ramp, "Emissions|NMVOC": 20.0 + 150.0 * ramp, "Emissions|CO2|Fossil": 12.0 * ramp ** 2 * (1.0 + 0.3 * s),
        "Emissions|CO2|Land Use": 0.5 + ramp, "Carbon Flux|Ocean": 2.5 * ramp,
    } for s in range(n)]


GAS_BINDS = {**syn.TWO_LAYER_BINDINGS, "beta": "TerrestrialCarbon.beta", "tau_oh": "CH4Chemistry.tau_oh", "tau_n2o": "N2OChemistry.tau_n2o",
             "ch4_0": "initial:Atmospheric Concentration|CH4"}


def test_gas_cycle_graph_structure():
    b = gas_cycle_builder()
    m = oracle_from_builder(b, gas_cycle_scenarios(1)[0])
    ens = b.build_ensemble(device=-2)
    assert ens.execution_order() == m.execution_order() and ens.variable_names == m.names
    # temperature feedbacks are lagged (TwoLayer is inserted last), concentrations feed GhgForcing in the same step
    assert ens.variable_source(0, "Surface Temperature") == 0 and ens.variable_source(4, "Atmospheric Concentration|CH4") == 2
    r = m.run()
    assert r["Atmospheric Concentration|CH4"][-1] > 1200.0 and r["Atmospheric Concentration|CO2"][-1] > 400.0
    assert 0.5 < r["Surface Temperature"][-1] < 8.0


@pytest.mark.gpu
@pytest.mark.parametrize("delay", [1, 3])
def test_gas_cycle_gpu_parity(delay, tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b = gas_cycle_builder(delay)
    ens = b.build_ensemble().bind_parameters(GAS_BINDS)
    scen = gas_cycle_scenarios(2)
    sc = ens.pack_scenarios(scen)
    rng = np.random.default_rng(23)
    tl = syn.uniform_params(syn.TWO_LAYER_RANGES, 200, 8)
    tl[:, 1] *= 0.2
    p = np.column_stack([tl, rng.uniform(0.4, 0.9, 200), rng.uniform(8.0, 11.0, 200), rng.uniform(110.0, 160.0, 200), rng.uniform(700.0, 760.0, 200)])
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    ref = m.split(m.run_batch(oracle_bindings(b, GAS_BINDS), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    assert np.isfinite(got["Surface Temperature"]).all()


@pytest.mark.gpu
def test_simple_box_components_gpu(tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    ospp = OceanSurfacePartialPressureBuilder.from_parameters(dict(
        ospp_preindustrial=278.0, sensitivity_ospp_to_temperature=0.043, delta_ospp_offsets=[1.5568, 7.4706, 1.2748, 2.4491, 1.5468],
        delta_ospp_coefficients=[-0.013993, -0.20207, -0.12015, -0.12639, -0.15326], sea_surface_temperature_preindustrial=17.9)).build()
    b = (ModelBuilder().with_time_axis(syn.time_axis(2000, 2050)).with_rust_component(ospp)
         .with_rust_component(FourBoxOceanHeatUptakeBuilder.from_parameters({}).build()))
    ens = b.build_ensemble().bind_parameters({"pi": "OceanSurfacePartialPressure.ospp_preindustrial", "r": "FourBoxOceanHeatUptake.northern_ocean_ratio"})
    years = syn.time_axis(2000, 2050).values()
    scen = [{"Sea Surface Temperature": 0.05 * (years - 2000), "Dissolved Inorganic Carbon": 0.1 * (years - 2000),
             "Effective Radiative Forcing|Aggregated": 0.04 * (years - 2000)}]
    sc = ens.pack_scenarios(scen)
    p = np.column_stack([np.linspace(270, 290, 64), np.linspace(1.0, 1.4, 64)])
    got = ens.split_outputs(ens.run(p, sc))
    m = oracle_from_builder(b)
    names = ens.variable_names
    binds = {"pi": "OceanSurfacePartialPressure.ospp_preindustrial", "r": "FourBoxOceanHeatUptake.northern_ocean_ratio"}
    ref = m.split(m.run_batch(oracle_bindings(b, binds), p, ens.exogenous_names, sc, names), names)
    for n in names:
        assert rel_err(got[n], ref[n]) <= 1e-9, n
    assert got["Heat Uptake|Ocean"].shape == (51, 4, 64)
