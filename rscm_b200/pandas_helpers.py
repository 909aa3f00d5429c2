"""pandas views of calibration and ensemble results.

Mirrors ``rscm.calibrate.pandas_helpers`` of the reference (python/rscm/calibrate/pandas_helpers.py: ``chain_to_dataframe``
:12-102, ``target_from_dataframe`` :105-225) and adds the two frames an ensemble produces: across-member quantile bands
(``Ensemble.run_quantiles``) and member timeseries (``Ensemble.run`` / ``split_outputs``).

One deliberate difference: the reference reshapes ``Chain.flat_samples`` — which is stored iteration-major
(crates/rscm-calibrate/src/sampler/chain.rs: one [walkers, params] block per stored iteration) — as if it were walker-major,
so its (walker, iteration) labels do not match the samples.  Here the labels follow the storage order.
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from .calibrate import Chain, Target

__all__ = ["chain_to_dataframe", "target_from_dataframe", "quantiles_to_dataframe", "members_to_dataframe"]


def chain_to_dataframe(chain: Chain, discard: int = 0) -> pd.DataFrame:
    """Long-form frame indexed by (walker, iteration) with one column per parameter plus ``log_prob``."""
    names = chain.param_names
    n_stored = len(chain) - discard
    flat = chain.flat_samples(discard)
    if n_stored <= 0 or flat.shape[0] == 0:
        return pd.DataFrame(columns=[*names, "log_prob"])
    n_walkers = flat.shape[0] // n_stored
    samples = flat.reshape(n_stored, n_walkers, len(names))               # iteration-major, as stored
    log_probs = chain.flat_log_probs(discard).reshape(n_stored, n_walkers)
    iterations = np.arange(discard, discard + n_stored) * chain.thin
    index = pd.MultiIndex.from_product([np.arange(n_walkers), iterations], names=["walker", "iteration"])
    data = {n: samples[:, :, j].T.reshape(-1) for j, n in enumerate(names)}  # walker-major rows, like the reference's index
    data["log_prob"] = log_probs.T.reshape(-1)
    return pd.DataFrame(data, index=index)


def target_from_dataframe(df: pd.DataFrame, time_col: str = "time", value_col: str = "value", uncertainty_col: str | None = None,
                          relative_error: float | None = None) -> Target:
    """``Target`` from a frame with a ``variable`` column; uncertainties from ``uncertainty_col`` (default column
    ``uncertainty``) or ``relative_error * |value|``."""
    if "variable" not in df.columns:
        raise ValueError("DataFrame must have 'variable' column for automatic variable detection. "
                         "For single-variable data, create Target manually:\n  target = Target()\n"
                         "  target.add_variable(variable_name, observations)")
    target = Target()
    for var_name, var_df in df.groupby("variable"):
        times, values = var_df[time_col].to_numpy(), var_df[value_col].to_numpy()
        if relative_error is not None:
            for t, v in zip(times, values):
                target.add_observation_relative(str(var_name), float(t), float(v), float(relative_error))
            continue
        col = uncertainty_col if uncertainty_col is not None else ("uncertainty" if "uncertainty" in var_df.columns else None)
        if col is None:
            raise ValueError(f"No uncertainty information provided for variable '{var_name}'. "
                             "Specify uncertainty_col or relative_error parameter.")
        for t, v, u in zip(times, values, var_df[col].to_numpy()):
            target.add_observation(str(var_name), float(t), float(v), float(u))
    return target


def quantiles_to_dataframe(quantiles: dict[str, np.ndarray], q, times, scenario_names=None) -> pd.DataFrame:
    """``Ensemble.run_quantiles`` result ``{variable: [n_q, T, (R,) S]}`` -> frame indexed by (variable, region, scenario, time)
    with one column per quantile."""
    frames = []
    for var, arr in quantiles.items():
        a = arr if arr.ndim == 4 else arr[:, :, None, :]
        nq, nt, nr, ns = a.shape
        scen = list(scenario_names) if scenario_names is not None else list(range(ns))
        idx = pd.MultiIndex.from_product([[var], range(nr), scen, np.asarray(times)[:nt]], names=["variable", "region", "scenario", "time"])
        frames.append(pd.DataFrame({float(qk): a[k].transpose(1, 2, 0).reshape(-1) for k, qk in enumerate(q)}, index=idx))
    return pd.concat(frames) if frames else pd.DataFrame()


def members_to_dataframe(outputs: dict[str, np.ndarray], times, n_members: int, scenario_names=None) -> pd.DataFrame:
    """``Ensemble.split_outputs`` result ``{variable: [T, (R,) S*M]}`` -> wide frame: rows (variable, region, scenario, member),
    one column per time — the shape scmdata / pandas users of the reference work with.  Meant for subsamples, not 2M runs."""
    frames = []
    for var, arr in outputs.items():
        a = arr if arr.ndim == 3 else arr[:, None, :]
        nt, nr, runs = a.shape
        ns = runs // n_members
        scen = list(scenario_names) if scenario_names is not None else list(range(ns))
        idx = pd.MultiIndex.from_product([[var], range(nr), scen, range(n_members)], names=["variable", "region", "scenario", "member"])
        frames.append(pd.DataFrame(a.transpose(1, 2, 0).reshape(nr * runs, nt), index=idx, columns=np.asarray(times)[:nt]))
    return pd.concat(frames) if frames else pd.DataFrame()
