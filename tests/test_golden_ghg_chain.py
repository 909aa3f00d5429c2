"""MAGICC7 golden vectors 03 / 04 / 05 of the reference's GHG-forcing regression suite
(tests/regression/test_ghg_forcing.py:623-830; data committed as tests/golden/ghg_forcing_03_04_05.npz by
tests/golden/make_golden.py), through the CPU oracle and through the CUDA path:

  04_ecs_sweep_{1.5,2,3,4,4.5}, 05_co2_only_forcing   ERF -> ClimateUDEB global-mean temperature, phased 5 / 3 / 3 %
  (ours, same bar)  CO2 concentration -> GhgForcing -> Sum aggregate -> ClimateUDEB: BASELINE config 4's coupling,
                    driven by the golden concentrations and compared with the golden temperature
  03_emissions_driven   SSP245 emissions -> chemistry + carbon cycle -> forcing -> temperature; `xfail` upstream at 5 %
                        ("documented simplifications"): kept xfail here, the measured errors are printed
"""

import json
import os

import numpy as np
import pytest

from rscm_b200 import synthetic as syn
from rscm_b200.core import GridType, InterpolationStrategy, ModelBuilder, TimeAxis, Timeseries, VariableSchema
from rscm_b200.magicc import (AerosolDirectBuilder, AerosolIndirectBuilder, CH4ChemistryBuilder, ClimateUDEBBuilder, CO2BudgetBuilder,
                              GhgForcingBuilder, N2OChemistryBuilder, OceanCarbonBuilder, OzoneForcingBuilder, TerrestrialCarbonBuilder)

from .helpers import oracle_from_builder
from .test_climate_udeb import AREA_W, phased, recorded_rel_err, udeb_builder

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ghg_forcing_03_04_05.npz"))
ERF_CASES = [f"04_ecs_sweep_{e}" for e in ("1.5", "2.0", "3.0", "4.0", "4.5")] + ["05_co2_only_forcing"]
PHASES = dict(shock_rtol=5e-2, converge_rtol=3e-2, final_rtol=3e-2)   # test_ghg_forcing.py:771-781, 819-829


def case(name):
    cfg = json.loads(str(G[name + "/config"]))
    erf_key = "Effective Radiative Forcing" if name.startswith("05") else "Effective Radiative Forcing|CO2"
    return G[name + "/years"], G[f"{name}/{erf_key}"], G[name + "/Surface Temperature"], cfg


def erf_model(name):
    # build_erf_to_temperature_model — test_ghg_forcing.py:78-137 (its "forcing_2xco2" key is not a field of
    # ClimateUDEBParameters and is dropped by the reference's deserialiser: rf_2xco2 keeps its default 3.71)
    years, erf, expected, cfg = case(name)
    return udeb_builder({"ecs": cfg.get("core_climatesensitivity", 3.0)}, years, erf), expected


def check_phases(actual, expected, label):
    for phase, (err, tol) in phased(actual, expected, **PHASES).items():
        assert err <= tol, f"{label} {phase}: {err:.4f} > {tol}"


@pytest.mark.parametrize("name", ERF_CASES)
def test_oracle_erf_to_temperature_matches_magicc7(name):
    b, expected = erf_model(name)
    check_phases(oracle_from_builder(b).run()["Surface Temperature"] @ AREA_W, expected, name)


def concentration_chain(name):
    """CO2 concentration -> GhgForcing (IPCCTAR; CH4 and N2O held at their pre-industrial values so that only CO2 forces,
    MAGICC's rf_total_runmodus = CO2) -> Sum aggregate -> ClimateUDEB: the coupling of BASELINE config 4."""
    cfg = json.loads(str(G[name + "/config"]))
    years, conc = G[name + "/years"], G[name + "/Atmospheric Concentrations|CO2"]
    axis = TimeAxis.from_bounds(np.concatenate([years, [years[-1] + 1.0]]))
    schema = VariableSchema()
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "ppm")
        schema.add_variable(f"Effective Radiative Forcing|{n}", "W/m^2")
    schema.add_variable("Surface Temperature", "K", GridType.FourBox)
    for n in ("Heat Uptake", "Ocean Heat Content", "Sea Surface Temperature"):
        schema.add_variable(n, "")
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", [f"Effective Radiative Forcing|{n}" for n in ("CO2", "CH4", "N2O")])
    ghg = GhgForcingBuilder.from_parameters({"method": "Ipcctar", "delq2xco2": cfg.get("core_delq2xco2", 3.71), "co2_pi": float(conc[0]),
                                             "ch4_pi": 722.0, "n2o_pi": 270.0,
                                             # IPCCTAR runs carry no rapid adjustment (build_ghg_forcing_model, :166-170)
                                             "adjust_co2": cfg.get("core_rfrapidadjust_co2", 1.0), "adjust_ch4": 1.0, "adjust_n2o": 1.0}).build()
    b = (ModelBuilder().with_time_axis(axis).with_schema(schema)
         .with_rust_component(ghg)
         .with_rust_component(ClimateUDEBBuilder.from_parameters({"ecs": cfg.get("core_climatesensitivity", 3.0)}).build())
         .with_initial_values({"Surface Temperature": 0.0, "Effective Radiative Forcing": 0.0}))
    for n, v in (("CO2", conc), ("CH4", np.full_like(conc, 722.0)), ("N2O", np.full_like(conc, 270.0))):
        b.with_exogenous_variable(f"Atmospheric Concentration|{n}", Timeseries(v, axis, "ppm", InterpolationStrategy.Linear))
    return b, G[name + "/Surface Temperature"], G[name + "/Effective Radiative Forcing|CO2"]


@pytest.mark.parametrize("name", ERF_CASES)
def test_oracle_concentration_to_temperature_chain_matches_magicc7(name):
    b, expected_t, expected_erf = concentration_chain(name)
    r = oracle_from_builder(b).run()
    # ERF|CO2 lands at index N+1 from the concentration at N: actual[1:] vs expected[:-1] (test_ghg_forcing.py:184-193)
    np.testing.assert_allclose(r["Effective Radiative Forcing|CO2"][1:], expected_erf[:-1], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["Effective Radiative Forcing"][1:], expected_erf[:-1], rtol=1e-5, atol=1e-6)
    # the temperature inherits the one-step lag of the forcing it is driven by: T[N+1] answers MAGICC7's T[N]
    check_phases((r["Surface Temperature"] @ AREA_W)[1:], expected_t[:-1], name + " (concentration-driven chain)")


# ---- 03: emissions-driven ------------------------------------------------------------------------------------------------
def emissions_driven_builder(define_first_step=False):
    """build_emissions_driven_model + _extract_emissions + the initial conditions of test_03 — test_ghg_forcing.py:395-660.
    As written upstream the ERF aggregate, the sea surface temperature and the two carbon fluxes have no value at index 0,
    so the first step of ClimateUDEB / TerrestrialCarbon / CO2Budget reads NaN and the run starts
    from clamped values; `define_first_step` gives them 0.0 (not part of the upstream test)."""
    name = "03_emissions_driven"
    cfg = json.loads(str(G[name + "/config"]))
    years = G[name + "/years"]
    get = lambda k: G[f"{name}/{k}"] if f"{name}/{k}" in G.files else np.zeros_like(years)   # noqa: E731
    sectors = lambda base: get(base + "|MAGICC Fossil and Industrial") + get(base + "|MAGICC AFOLU")   # noqa: E731
    em = {"Emissions|CO2|Fossil": get("Emissions|CO2"), "Emissions|CO2|Land Use": np.zeros_like(years),
          "Emissions|CH4": get("Emissions|CH4"), "Emissions|N2O": get("Emissions|N2O"), "EESC": np.zeros_like(years)}
    for sp in ("NOx", "CO", "NMVOC", "SOx", "BC", "OC"):
        em[f"Emissions|{sp}"] = sectors(f"Emissions|{sp}")
    co2_0, ch4_0, n2o_0 = (float(get(f"Atmospheric Concentrations|{g}")[0]) for g in ("CO2", "CH4", "N2O"))
    init = {"Atmospheric Concentration|CO2": co2_0, "Atmospheric Concentration|CH4": ch4_0, "Atmospheric Concentration|N2O": n2o_0,
            "Surface Temperature": 0.0, "Ocean Surface pCO2": co2_0, "Cumulative Ocean Uptake": 0.0, "Carbon Pool|Plant": 884.86,
            "Carbon Pool|Detritus": 92.77, "Carbon Pool|Soil": 1681.53, "Carbon Pool|Humus": 836.0}
    axis = TimeAxis.from_bounds(np.concatenate([years, [years[-1] + 1.0]]))
    schema = VariableSchema()
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "")
    for n in ("CO2|Fossil", "CO2|Land Use", "CH4", "N2O", "NOx", "CO", "NMVOC", "SOx", "BC", "OC"):
        schema.add_variable(f"Emissions|{n}", "")
    schema.add_variable("EESC", "ppt")
    for n in syn.CONFIG4_ERF_PARTS:
        schema.add_variable(n, "W/m^2")
    schema.add_variable("Surface Temperature", "K", GridType.FourBox)
    for n in ("Heat Uptake", "Ocean Heat Content", "Sea Surface Temperature", "Carbon Flux|Terrestrial", "Carbon Flux|Ocean", "Carbon Pool|Plant",
              "Carbon Pool|Detritus", "Carbon Pool|Soil", "Carbon Pool|Humus", "Ocean Surface pCO2", "Cumulative Ocean Uptake",
              "Emissions|CO2|Net", "Airborne Fraction|CO2", "Lifetime|CH4", "Lifetime|N2O"):
        schema.add_variable(n, "")
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", syn.CONFIG4_ERF_PARTS)
    b = (ModelBuilder().with_time_axis(axis).with_schema(schema)
         .with_rust_component(CH4ChemistryBuilder.from_parameters({"ch4_pi": ch4_0}).build())
         .with_rust_component(N2OChemistryBuilder.from_parameters({"n2o_pi": n2o_0}).build())
         .with_rust_component(GhgForcingBuilder.from_parameters({"method": "Ipcctar", "delq2xco2": cfg.get("core_delq2xco2", 3.71),
                                                                 "co2_pi": co2_0, "ch4_pi": ch4_0, "n2o_pi": n2o_0}).build())
         .with_rust_component(OzoneForcingBuilder.from_parameters({}).build())
         .with_rust_component(AerosolDirectBuilder.from_parameters({}).build())
         .with_rust_component(AerosolIndirectBuilder.from_parameters({}).build())
         .with_rust_component(ClimateUDEBBuilder.from_parameters({"ecs": cfg.get("core_climatesensitivity", 3.0)}).build())
         .with_rust_component(TerrestrialCarbonBuilder.from_parameters({}).build())
         .with_rust_component(OceanCarbonBuilder.from_parameters({}).build())
         .with_rust_component(CO2BudgetBuilder.from_parameters({}).build())
         .with_initial_values(init))
    if define_first_step:
        b.with_initial_values({"Effective Radiative Forcing": 0.0, "Sea Surface Temperature": 0.0, "Carbon Flux|Terrestrial": 0.0,
                               "Carbon Flux|Ocean": 0.0})
    for k, v in em.items():
        b.with_exogenous_variable(k, Timeseries(np.asarray(v, dtype=float), axis, "", InterpolationStrategy.Linear))
    expected = {"Atmospheric Concentration|CO2": get("Atmospheric Concentrations|CO2"), "Atmospheric Concentration|CH4": get("Atmospheric Concentrations|CH4"),
                "Atmospheric Concentration|N2O": get("Atmospheric Concentrations|N2O"), "Sea Surface Temperature": get("Surface Temperature")}
    return b, expected


def emissions_errors(series):
    _, expected = emissions_driven_builder()
    return {k: recorded_rel_err(np.asarray(series[k])[1:], e[:-1]) for k, e in expected.items()}


@pytest.mark.xfail(reason="upstream xfail: diverges from MAGICC7 due to documented simplifications (reference issues #108-#110)", strict=False)
def test_oracle_emissions_driven_03():
    b, _ = emissions_driven_builder()
    errs = emissions_errors(oracle_from_builder(b).run())
    print("03_emissions_driven, oracle, max relative error vs MAGICC7:", {k: f"{v:.3f}" for k, v in errs.items()})
    assert max(errs.values()) <= 5e-2, errs


def test_emissions_driven_03_with_a_defined_first_step_on_the_oracle():
    """Not an upstream test: with the index-0 values the upstream test leaves undefined set to 0, the chain runs 1750-2100 on
    the SSP245 emissions, stays finite and tracks the golden concentrations within 30 % (CO2), 20 % (CH4), 10 % (N2O) —
    measured 0.25 / 0.14 / 0.07: the size of the "documented simplifications" the upstream xfail refers to.  As written
    upstream the first step reads NaN forcing and the run starts from clamped values (CO2 at 895 ppm in 1751)."""
    b, _ = emissions_driven_builder(define_first_step=True)
    r = oracle_from_builder(b).run()
    errs = emissions_errors(r)
    print("03_emissions_driven with defined first step, oracle:", {k: f"{v:.3f}" for k, v in errs.items()})
    assert all(np.isfinite(np.asarray(r[k])[1:]).all() for k in errs)
    assert errs["Atmospheric Concentration|CO2"] <= 0.30 and errs["Atmospheric Concentration|CH4"] <= 0.20 and errs["Atmospheric Concentration|N2O"] <= 0.10, errs


# ---- the same vectors through the CUDA path ------------------------------------------------------------------------------
def gpu_series(b):
    model = b.build()
    model.run()
    ts = model.timeseries()
    return model, ts


@pytest.mark.gpu
@pytest.mark.parametrize("name", ERF_CASES)
def test_gpu_erf_to_temperature_matches_magicc7(name, tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b, expected = erf_model(name)
    _, ts = gpu_series(b)
    t4 = np.asarray(ts.get_fourbox_timeseries_by_name("Surface Temperature").values())
    check_phases(t4 @ AREA_W, expected, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["04_ecs_sweep_1.5", "04_ecs_sweep_4.5", "05_co2_only_forcing"])
def test_gpu_concentration_to_temperature_chain_matches_magicc7(name, tmp_path, monkeypatch):
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b, expected_t, expected_erf = concentration_chain(name)
    _, ts = gpu_series(b)
    erf = np.asarray(ts.get_timeseries_by_name("Effective Radiative Forcing").values())
    np.testing.assert_allclose(erf[1:], expected_erf[:-1], rtol=1e-5, atol=1e-6)
    t4 = np.asarray(ts.get_fourbox_timeseries_by_name("Surface Temperature").values())
    check_phases((t4 @ AREA_W)[1:], expected_t[:-1], name + " (concentration-driven chain)")


@pytest.mark.gpu
@pytest.mark.parametrize("define_first_step", [False, True])
def test_gpu_emissions_driven_03_equals_oracle_and_records_the_golden_error(define_first_step, tmp_path, monkeypatch):
    """03 is xfail upstream; what is required here is that the CUDA path and the oracle agree on it to 1e-9 and that the
    distance to MAGICC7 is the same on both sides."""
    monkeypatch.setenv("RSCM_B200_CACHE", str(tmp_path))
    b, expected = emissions_driven_builder(define_first_step)
    _, ts = gpu_series(b)
    ref = oracle_from_builder(b).run()
    got = {k: np.asarray(ts.get_timeseries_by_name(k).values()) for k in expected}
    for k in expected:
        e = np.asarray(ref[k])
        assert np.array_equal(np.isnan(got[k]), np.isnan(e))
        ok = ~np.isnan(e)
        assert np.max(np.abs(got[k][ok] - e[ok])) / np.max(np.abs(e[ok])) <= 1e-9, k
    errs = emissions_errors(got)
    print("03_emissions_driven, CUDA path, max relative error vs MAGICC7:", {k: f"{v:.3f}" for k, v in errs.items()})
