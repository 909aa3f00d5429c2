"""BASELINE workloads described directly to the CPU ORACLE (test infrastructure, NOT product code).

``bench.py --impl reference`` must not map the CUDA library into its process, so the CPU arm cannot go through
``rscm_b200.core.ModelBuilder``.  This module states the headline graph (BASELINE configs[2]) a second time, against the
oracle's own builder, from the constants in ``rscm_b200/synthetic_data.py`` (loaded by file path: no package import).
``tests/test_bench_contract.py`` checks that this description and ``rscm_b200.synthetic.coupled_builder`` lowered through
``tests/helpers.oracle_from_builder`` give bit-identical oracle runs.
"""

from __future__ import annotations

import importlib.util
import os

import numpy as np

from . import oracle as orc

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synthetic_data():
    """rscm_b200/synthetic_data.py as a stand-alone module (pure numpy; does not run rscm_b200/__init__.py)."""
    spec = importlib.util.spec_from_file_location("_rscm_synthetic_data", os.path.join(_ROOT, "rscm_b200", "synthetic_data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def coupled_model(sd=None, conc_pi: float = 278.0):
    """CarbonCycle -> CO2ERF -> Sum aggregate -> TwoLayer exactly as docs/notebooks/coupled_model.py:435-483 builds it
    (schema, insertion order, initial values).  Returns (OracleModel, bindings) with one binding list per column of
    ``synthetic_data.COUPLED_RANGES``."""
    sd = sd or synthetic_data()
    m = orc.OracleModel()
    cc = sd.COUPLED_CARBON_DEFAULTS
    m.add_component(orc.CARBON_CYCLE, [cc["tau"], conc_pi, cc["alpha_temperature"], 0.1])  # step_size: SolverOptions default
    m.add_component(orc.CO2_ERF, [sd.COUPLED_ERF_2XCO2, conc_pi])
    tl = sd.TWO_LAYER_DEFAULTS
    tl_fields = ["lambda0", "a", "efficacy", "eta", "heat_capacity_surface", "heat_capacity_deep"]
    m.add_component(orc.TWO_LAYER, [tl[k] for k in tl_fields])
    m.set_time_bounds(sd.time_bounds(sd.years()))
    for name in ("Emissions|CO2|Anthropogenic", "Surface Temperature", "Atmospheric Concentration|CO2", "Cumulative Land Uptake",
                 "Cumulative Emissions|CO2", "Effective Radiative Forcing|CO2", "Deep Ocean Temperature"):
        m.add_schema_variable(name, orc.SCALAR)
    m.add_aggregate("Effective Radiative Forcing", orc.AGG_SUM, ["Effective Radiative Forcing|CO2"], None, orc.SCALAR)
    for k, v in {"Cumulative Land Uptake": 0.0, "Cumulative Emissions|CO2": 0.0, "Atmospheric Concentration|CO2": conc_pi,
                 "Surface Temperature": 0.0, "Deep Ocean Temperature": 0.0}.items():
        m.set_initial_value(k, v)
    m.build()
    where = {"tau": (0, 0), "alpha_temperature": (0, 2), "erf_2xco2": (1, 0), **{k: (2, i) for i, k in enumerate(tl_fields)}}
    bindings = [[where[k]] for k in sd.COUPLED_RANGES]
    return m, bindings


def coupled_scenarios(sd, S: int) -> np.ndarray:
    """[S][T] emission series, the oracle's scenario layout for one scalar exogenous variable."""
    return np.ascontiguousarray(sd.emission_scenarios(sd.years(), S))
