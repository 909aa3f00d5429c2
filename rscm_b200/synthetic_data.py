"""Seeded synthetic scenario curves, parameter ranges and parameter matrices of the shapes BASELINE.json names
(SURVEY.md §8(d)).  Pure numpy and no package imports: ``bench.py --impl reference`` loads this file by path, so the
CPU arm sees the same inputs as the GPU arm without mapping the CUDA library into its process.

``numpy.random.Generator(PCG64(20260101 + config index))``, config index = position in BASELINE.json's ``configs`` list.
"""

from __future__ import annotations

import numpy as np

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
COUPLED_CARBON_DEFAULTS = {"tau": 20.3, "alpha_temperature": 0.05}
COUPLED_ERF_2XCO2 = 3.7
COUPLED_OUTPUTS = [
    "Atmospheric Concentration|CO2", "Cumulative Emissions|CO2", "Cumulative Land Uptake",
    "Effective Radiative Forcing|CO2", "Effective Radiative Forcing", "Surface Temperature", "Deep Ocean Temperature",
]
COUPLED_EXOGENOUS = "Emissions|CO2|Anthropogenic"


def years(start: int = 1750, end: int = 2100) -> np.ndarray:
    # python/rscm/config/builder.py:91-97: np.arange(start, end + 1)
    return np.arange(start, end + 1, dtype=float)


def time_bounds(values: np.ndarray) -> np.ndarray:
    """TimeAxis::from_values (crates/rscm-core/src/timeseries.rs:66-79): one extra bound = last + (last - previous)."""
    v = np.asarray(values, dtype=float)
    return np.concatenate([v, [v[-1] + (v[-1] - v[-2])]])


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


def config3_params(M: int, rank: int = 0, world: int = 1) -> np.ndarray:
    """Parameter block of BASELINE config 3 for one rank.  Weak scaling: every rank owns an M-member block of its own
    (rank 0's block is the single-GPU ensemble)."""
    return uniform_params(COUPLED_RANGES, M, SEED0 + 2 + 1000 * rank)


def config5_observations(truth_temperature: np.ndarray, years: np.ndarray, sigma: float = 0.1, seed: int = SEED0 + 4):
    """annual observations 1850-2020 (K = 171) = truth + N(0, sigma^2); `truth_temperature` [T]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    obs = []
    for i, y in enumerate(years):
        if 1850.0 <= y <= 2020.0:
            obs.append(("Surface Temperature", float(y), float(truth_temperature[i] + sigma * rng.standard_normal()), sigma))
    return obs


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
