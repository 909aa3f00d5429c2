"""Seeded synthetic scenarios and parameter ensembles of the shapes BASELINE.json names
(SURVEY.md §8(d)); shared by bench.py, __graft_entry__.smoke() and the tests so that
every leg (GPU, CPU oracle, reference arm) sees the same inputs.

``numpy.random.Generator(PCG64(20260101 + config index))``, config index = position in
BASELINE.json's ``configs`` list.
"""

from __future__ import annotations

import numpy as np

from .components import CarbonCycleBuilder, CO2ERFBuilder
from .core import ModelBuilder, TimeAxis, VariableSchema
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
