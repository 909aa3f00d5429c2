"""Seeded synthetic scenarios and parameter ensembles of the shapes BASELINE.json names
(SURVEY.md §8(d)); shared by bench.py, __graft_entry__.smoke() and the tests so that
every leg (GPU, CPU oracle, reference arm) sees the same inputs.

``numpy.random.Generator(PCG64(20260101 + config index))``, config index = position in
BASELINE.json's ``configs`` list.
"""

from __future__ import annotations

import numpy as np

from .components import CarbonCycleBuilder, CO2ERFBuilder
from .core import GridType, ModelBuilder, TimeAxis, VariableSchema
from .two_layer import TwoLayerBuilder

from .synthetic_data import (  # noqa: F401  (pure-numpy inputs, shared with the CPU arm of bench.py)
    COUPLED_CARBON_DEFAULTS, COUPLED_ERF_2XCO2, COUPLED_OUTPUTS, COUPLED_RANGES, SEED0, TWO_LAYER_DEFAULTS, TWO_LAYER_RANGES,
    config3_params, config4_scenario, config5_observations, emission_scenarios, ssp_like_forcing, uniform_params, years,
)


def time_axis(start: int = 1750, end: int = 2100) -> TimeAxis:
    # python/rscm/config/builder.py:91-97: np.arange(start, end + 1) -> TimeAxis.from_values
    return TimeAxis.from_values(years(start, end))


def two_layer_builder(params: dict | None = None, axis: TimeAxis | None = None) -> ModelBuilder:
    """BASELINE config 1/2/5 graph (python/rscm/config/builder.py:19-108)."""
    p = dict(TWO_LAYER_DEFAULTS)
    p.update(params or {})
    return (
        ModelBuilder()
        .with_time_axis(axis or time_axis())
        .with_rust_component(TwoLayerBuilder.from_parameters(p).build())
        .with_initial_values({"Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0})
    )


TWO_LAYER_BINDINGS = {k: f"TwoLayer.{k}" for k in TWO_LAYER_RANGES}


def coupled_builder(axis: TimeAxis | None = None, conc_pi: float = 278.0) -> ModelBuilder:
    """BASELINE config 3 graph: CarbonCycle -> CO2ERF -> Sum aggregate -> TwoLayer, built exactly as
    docs/notebooks/coupled_model.py:435-483 (schema, insertion order, initial values)."""
    schema = VariableSchema()
    for name, unit in [
        ("Emissions|CO2|Anthropogenic", "GtC / yr"), ("Surface Temperature", "K"), ("Atmospheric Concentration|CO2", "ppm"),
        ("Cumulative Land Uptake", "Gt C"), ("Cumulative Emissions|CO2", "Gt C"), ("Effective Radiative Forcing|CO2", "W/m^2"),
        ("Deep Ocean Temperature", "K"),
    ]:
        schema.add_variable(name, unit)
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", ["Effective Radiative Forcing|CO2"])
    return (
        ModelBuilder()
        .with_time_axis(axis or time_axis())
        .with_schema(schema)
        .with_rust_component(CarbonCycleBuilder.from_parameters({"conc_pi": conc_pi, **COUPLED_CARBON_DEFAULTS}).build())
        .with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": COUPLED_ERF_2XCO2, "conc_pi": conc_pi}).build())
        .with_rust_component(TwoLayerBuilder.from_parameters(TWO_LAYER_DEFAULTS).build())
        .with_initial_values({
            "Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": conc_pi,
            "Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0,
        })
    )


COUPLED_BINDINGS = {
    "tau": "CarbonCycle.tau",
    "alpha_temperature": "CarbonCycle.alpha_temperature",
    "erf_2xco2": "CO2ERF.erf_2xco2",
    **TWO_LAYER_BINDINGS,
}


def config2(M: int = 1 << 20):
    """two-layer ensemble: M parameter sets x 1 forcing scenario."""
    axis = time_axis()
    params = uniform_params(TWO_LAYER_RANGES, M, SEED0 + 1)
    forcing = ssp_like_forcing(axis.values())
    return two_layer_builder(axis=axis), TWO_LAYER_BINDINGS, params, [{"Effective Radiative Forcing": forcing}]


def config3(M: int = 1 << 18, S: int = 8):
    """coupled carbon cycle + CO2 ERF + two-layer: M members x S emission scenarios."""
    axis = time_axis()
    params = config3_params(M)
    em = emission_scenarios(axis.values(), S)
    return coupled_builder(axis=axis), COUPLED_BINDINGS, params, [{"Emissions|CO2|Anthropogenic": em[s]} for s in range(S)]


# ---- config 4: MAGICC box components on the four-box grid ------------------------------------------------------
CONFIG4_RANGES = {"ecs": (1.5, 4.5), "kappa": (0.5, 1.5), "rlo": (1.1, 1.5)}
CONFIG4_BINDINGS = {"ecs": "ClimateUDEB.ecs", "kappa": "ClimateUDEB.kappa", "rlo": "ClimateUDEB.rlo"}
CONFIG4_ERF_PARTS = [
    "Effective Radiative Forcing|CO2", "Effective Radiative Forcing|CH4", "Effective Radiative Forcing|N2O",
    "Effective Radiative Forcing|O3|Stratospheric", "Effective Radiative Forcing|O3|Tropospheric",
    "Effective Radiative Forcing|O3|Temperature Feedback", "Effective Radiative Forcing|Aerosol|Direct",
    "Effective Radiative Forcing|Aerosol|Indirect",
]
CONFIG4_OUTPUTS = ["Surface Temperature", "Heat Uptake", "Ocean Heat Content", "Sea Surface Temperature", "Effective Radiative Forcing"]


def config4_builder(axis: TimeAxis | None = None) -> ModelBuilder:
    """GhgForcing + OzoneForcing + AerosolDirect (FourBox, stored Scalar) + AerosolIndirect -> Sum aggregate ->
    ClimateUDEB (FourBox Surface Temperature), the schema of the reference's full-forcing regression model
    (tests/regression/test_ghg_forcing.py:395-464).  The aggregate gets an initial value so that ClimateUDEB's
    erf_start is defined at the first step."""
    from .core import GridType
    from .magicc import AerosolDirectBuilder, AerosolIndirectBuilder, ClimateUDEBBuilder, GhgForcingBuilder, OzoneForcingBuilder

    schema = VariableSchema()
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "ppm")
    for n in ("NOx", "CO", "NMVOC", "SOx", "BC", "OC"):
        schema.add_variable(f"Emissions|{n}", "Mt/yr")
    schema.add_variable("EESC", "ppt")
    for n in CONFIG4_ERF_PARTS:
        schema.add_variable(n, "W/m^2")
    schema.add_variable("Surface Temperature", "K", GridType.FourBox)
    for n in ("Heat Uptake", "Ocean Heat Content", "Sea Surface Temperature"):
        schema.add_variable(n, "")
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", CONFIG4_ERF_PARTS)
    return (
        ModelBuilder().with_time_axis(axis or time_axis()).with_schema(schema)
        .with_rust_component(GhgForcingBuilder.from_parameters({}).build())
        .with_rust_component(OzoneForcingBuilder.from_parameters({}).build())
        .with_rust_component(AerosolDirectBuilder.from_parameters({}).build())
        .with_rust_component(AerosolIndirectBuilder.from_parameters({}).build())
        .with_rust_component(ClimateUDEBBuilder.from_parameters({}).build())
        .with_initial_values({"Surface Temperature": 0.0, "Effective Radiative Forcing": 0.0})
    )


def config4(M: int = 100_000):
    axis = time_axis()
    params = uniform_params(CONFIG4_RANGES, M, SEED0 + 3)
    return config4_builder(axis), CONFIG4_BINDINGS, params, [config4_scenario(axis.values())]


# ---- the emissions-driven MAGICC chain: all eleven box components of rscm-magicc in one graph ------------------------------
# (the model of the reference's regression suite, tests/regression/test_ghg_forcing.py:470-620, plus HalocarbonChemistry;
# BASELINE configs[3] at its widest: 124 variables, run-time compiled)
def full_chain_builder(start=1850, end=1950, halocarbons=False):
    """Emissions-driven MAGICC: CH4/N2O chemistry, terrestrial + ocean carbon, CO2 budget, GHG / ozone / aerosol forcing,
    Sum aggregate (with an initial value), ClimateUDEB on the four-box grid.  ``halocarbons=True`` wires HalocarbonChemistry in:
    its EESC replaces the exogenous one and Forcing|Halocarbons joins the ERF aggregate."""
    from .magicc import (AerosolDirectBuilder, AerosolIndirectBuilder, CH4ChemistryBuilder, ClimateUDEBBuilder, CO2BudgetBuilder,
                         GhgForcingBuilder, HalocarbonChemistryBuilder, N2OChemistryBuilder, OceanCarbonBuilder, OzoneForcingBuilder,
                         TerrestrialCarbonBuilder)
    species = HalocarbonChemistryBuilder.species_names() if halocarbons else []
    erf_parts = list(CONFIG4_ERF_PARTS) + (["Forcing|Halocarbons"] if halocarbons else [])
    schema = VariableSchema()
    for s in species:
        schema.add_variable(f"Emissions|{s}", "kt/yr")
        schema.add_variable(f"Atmospheric Concentration|{s}", "ppt")
    for n in (("Forcing|Halocarbons", "Forcing|F-gases", "Forcing|Montreal Gases") if halocarbons else ()):
        schema.add_variable(n, "W/m^2")
    for n in ("CH4", "N2O", "NOx", "CO", "NMVOC", "SOx", "BC", "OC", "CO2|Fossil", "CO2|Land Use"):
        schema.add_variable(f"Emissions|{n}", "")
    schema.add_variable("EESC", "ppt")
    for n in ("CO2", "CH4", "N2O"):
        schema.add_variable(f"Atmospheric Concentration|{n}", "")
    for n in CONFIG4_ERF_PARTS:
        schema.add_variable(n, "W/m^2")
    schema.add_variable("Surface Temperature", "K", GridType.FourBox)
    for n in ("Heat Uptake", "Ocean Heat Content", "Sea Surface Temperature", "Carbon Flux|Terrestrial", "Carbon Flux|Ocean", "Carbon Pool|Plant",
              "Carbon Pool|Detritus", "Carbon Pool|Soil", "Carbon Pool|Humus", "Ocean Surface pCO2", "Cumulative Ocean Uptake",
              "Emissions|CO2|Net", "Airborne Fraction|CO2", "Lifetime|CH4", "Lifetime|N2O"):
        schema.add_variable(n, "")
    schema.add_aggregate("Effective Radiative Forcing", "W/m^2", "Sum", erf_parts)
    b = ModelBuilder().with_time_axis(time_axis(start, end)).with_schema(schema)
    if halocarbons:
        b = (b.with_rust_component(HalocarbonChemistryBuilder.from_parameters({}).build())
             .with_initial_values({f"Atmospheric Concentration|{s}": (500.0 if s == "CH3Cl" else 5.0 if s == "CH3Br" else 0.0) for s in species}))
    return (
        b
        .with_rust_component(CH4ChemistryBuilder.from_parameters({}).build())
        .with_rust_component(N2OChemistryBuilder.from_parameters({}).build())
        .with_rust_component(TerrestrialCarbonBuilder.from_parameters({}).build())
        .with_rust_component(OceanCarbonBuilder.from_parameters({}).build())
        .with_rust_component(CO2BudgetBuilder.from_parameters({}).build())
        .with_rust_component(GhgForcingBuilder.from_parameters({"method": "Ipcctar"}).build())
        .with_rust_component(OzoneForcingBuilder.from_parameters({}).build())
        .with_rust_component(AerosolDirectBuilder.from_parameters({}).build())
        .with_rust_component(AerosolIndirectBuilder.from_parameters({}).build())
        .with_rust_component(ClimateUDEBBuilder.from_parameters({}).build())
        .with_initial_values({"Atmospheric Concentration|CH4": 722.0, "Atmospheric Concentration|N2O": 270.0,
                              "Atmospheric Concentration|CO2": 278.0, "Carbon Pool|Plant": 884.86, "Carbon Pool|Detritus": 92.77,
                              "Carbon Pool|Soil": 1681.53, "Carbon Pool|Humus": 836.0, "Ocean Surface pCO2": 278.0,
                              "Cumulative Ocean Uptake": 0.0, "Surface Temperature": 0.0, "Effective Radiative Forcing": 0.0,
                              "Sea Surface Temperature": 0.0, "Carbon Flux|Terrestrial": 0.0, "Carbon Flux|Ocean": 0.0})
    )


def full_chain_scenario(start=1850, end=1950, f=1.0):
    years = time_axis(start, end).values()
    ramp = (years - start) / 100.0
    return {"Emissions|CH4": 50.0 + 250.0 * ramp * f, "Emissions|N2O": 1.0 + 6.0 * ramp, "Emissions|NOx": 5.0 + 30.0 * ramp,
            "Emissions|CO": 100.0 + 500.0 * ramp, "Emissions|NMVOC": 20.0 + 100.0 * ramp, "Emissions|SOx": 1.0 + 50.0 * ramp,
            "Emissions|BC": 2.5 + 4.0 * ramp, "Emissions|OC": 10.0 + 15.0 * ramp, "Emissions|CO2|Fossil": 8.0 * ramp ** 2 * f,
            "Emissions|CO2|Land Use": 0.5 + 0.8 * ramp, "EESC": 1000.0 + 900.0 * ramp}


FULL_CHAIN_BINDINGS = {"ecs": "ClimateUDEB.ecs", "beta": "TerrestrialCarbon.beta", "tau": "OceanCarbon.gas_exchange_tau", "tau_oh": "CH4Chemistry.tau_oh"}
FULL_CHAIN_RANGES = {"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}
FULL_CHAIN_OUTPUTS = ["Surface Temperature", "Atmospheric Concentration|CO2", "Atmospheric Concentration|CH4", "Effective Radiative Forcing", "EESC"]


def halocarbon_emissions(n: int, f: float = 1.0) -> dict:
    """Ramping emissions (kt/yr) of the 41 halocarbon species over n time points."""
    from .magicc import HalocarbonChemistryBuilder
    t = np.arange(n) / 100.0
    return {f"Emissions|{s}": f * (1.0 + i % 7) * (0.5 + 4.0 * t) for i, s in enumerate(HalocarbonChemistryBuilder.species_names())}


def full_chain(M: int = 100_000, start: int = 1850, end: int = 2100):
    """Builder, bindings, parameter sets and the scenario of the full chain (HalocarbonChemistry's EESC replaces the exogenous one)."""
    scen = full_chain_scenario(start, end)
    scen.pop("EESC")
    scen.update(halocarbon_emissions(end - start + 1))
    return full_chain_builder(start, end, halocarbons=True), FULL_CHAIN_BINDINGS, uniform_params(FULL_CHAIN_RANGES, M, 43), [scen]
