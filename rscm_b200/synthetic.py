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

SEED0 = 20260101

# typical ranges: python/rscm/config/models/two_layer.py:53-103 (reference)
TWO_LAYER_RANGES = {
    "lambda0": (0.8, 1.5),
    "a": (0.0, 0.1),
    "efficacy": (1.0, 1.8),
    "eta": (0.5, 1.0),
    "heat_capacity_surface": (5.0, 15.0),
    "heat_capacity_deep": (50.0, 200.0),
}
TWO_LAYER_DEFAULTS = {  # configs/two-layer/defaults.toml
    "lambda0": 1.0, "a": 0.0, "efficacy": 1.0, "eta": 0.7, "heat_capacity_surface": 8.0, "heat_capacity_deep": 100.0,
}
COUPLED_RANGES = {"tau": (15.0, 40.0), "alpha_temperature": (0.0, 0.15), "erf_2xco2": (3.4, 4.0), **TWO_LAYER_RANGES}


def time_axis(start: int = 1750, end: int = 2100) -> TimeAxis:
    # python/rscm/config/builder.py:91-97: np.arange(start, end + 1) -> TimeAxis.from_values
    return TimeAxis.from_values(np.arange(start, end + 1, dtype=float))


def ssp_like_forcing(years: np.ndarray) -> np.ndarray:
    """cfg 1/2: F(t) = 3.71 log2(C/278), C = 278 exp(0.0045 max(0, t-1850)) (about 4 W/m2 in 2100)."""
    conc = 278.0 * np.exp(0.0045 * np.maximum(0.0, years - 1850.0))
    return 3.71 * np.log2(conc / 278.0)


def emission_scenarios(years: np.ndarray, n: int = 8) -> np.ndarray:
    """cfg 3: piecewise-linear GtC/yr through the knots of docs/notebooks/coupled_model.py:397-398,
    the n scenarios scaling the post-2020 knots by {0, 0.25, ...}.  Returns [n, T]."""
    knots = np.array([1750.0, 1850.0, 1950.0, 2000.0, 2020.0, 2050.0, 2100.0])
    base = np.array([0.0, 0.5, 3.0, 7.0, 10.0, 5.0, 1.0])
    out = np.empty((n, years.size))
    for s in range(n):
        v = base.copy()
        v[5:] *= 0.25 * s
        out[s] = np.interp(years, knots, v)
    return out


def uniform_params(ranges: dict, M: int, seed: int) -> np.ndarray:
    """[M, len(ranges)] row-per-member matrix (the reference's &[Vec<f64>])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo = np.array([r[0] for r in ranges.values()])
    hi = np.array([r[1] for r in ranges.values()])
    return lo + (hi - lo) * rng.random((M, lo.size))


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
        .with_rust_component(CarbonCycleBuilder.from_parameters({"tau": 20.3, "conc_pi": conc_pi, "alpha_temperature": 0.05}).build())
        .with_rust_component(CO2ERFBuilder.from_parameters({"erf_2xco2": 3.7, "conc_pi": conc_pi}).build())
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
COUPLED_OUTPUTS = [
    "Atmospheric Concentration|CO2", "Cumulative Emissions|CO2", "Cumulative Land Uptake",
    "Effective Radiative Forcing|CO2", "Effective Radiative Forcing", "Surface Temperature", "Deep Ocean Temperature",
]


def config2(M: int = 1 << 20):
    """two-layer ensemble: M parameter sets x 1 forcing scenario."""
    axis = time_axis()
    params = uniform_params(TWO_LAYER_RANGES, M, SEED0 + 1)
    forcing = ssp_like_forcing(axis.values())
    return two_layer_builder(axis=axis), TWO_LAYER_BINDINGS, params, [{"Effective Radiative Forcing": forcing}]


def config3(M: int = 1 << 18, S: int = 8):
    """coupled carbon cycle + CO2 ERF + two-layer: M members x S emission scenarios."""
    axis = time_axis()
    params = uniform_params(COUPLED_RANGES, M, SEED0 + 2)
    em = emission_scenarios(axis.values(), S)
    return coupled_builder(axis=axis), COUPLED_BINDINGS, params, [{"Emissions|CO2|Anthropogenic": em[s]} for s in range(S)]


def config5_observations(truth_temperature: np.ndarray, years: np.ndarray, sigma: float = 0.1, seed: int = SEED0 + 4):
    """annual observations 1850-2020 (K = 171) = truth + N(0, sigma^2); `truth_temperature` [T]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    obs = []
    for i, y in enumerate(years):
        if 1850.0 <= y <= 2020.0:
            obs.append(("Surface Temperature", float(y), float(truth_temperature[i] + sigma * rng.standard_normal()), sigma))
    return obs


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


def config4_scenario(years: np.ndarray) -> dict:
    """Smooth synthetic SSP-like concentration / emission curves."""
    ramp = np.maximum(0.0, years - 1850.0) / 250.0
    return {
        "Atmospheric Concentration|CO2": 278.0 * np.exp(0.0035 * np.maximum(0.0, years - 1850.0)),
        "Atmospheric Concentration|CH4": 722.0 + 1100.0 * ramp,
        "Atmospheric Concentration|N2O": 270.0 + 60.0 * ramp,
        "EESC": 1000.0 + 1500.0 * np.exp(-((years - 2000.0) / 40.0) ** 2),
        "Emissions|NOx": 10.0 + 30.0 * ramp, "Emissions|CO": 300.0 * ramp, "Emissions|NMVOC": 100.0 * ramp,
        "Emissions|SOx": 1.0 + 60.0 * ramp * np.exp(-np.maximum(0.0, years - 1990.0) / 60.0),
        "Emissions|BC": 2.5 + 5.0 * ramp, "Emissions|OC": 10.0 + 20.0 * ramp,
    }


def config4(M: int = 100_000):
    axis = time_axis()
    params = uniform_params(CONFIG4_RANGES, M, SEED0 + 3)
    return config4_builder(axis), CONFIG4_BINDINGS, params, [config4_scenario(axis.values())]
