"""pandas helpers (reference: python/rscm/calibrate/pandas_helpers.py and its tests/test_calibrate_pandas.py)."""

import numpy as np
import pandas as pd
import pytest

from rscm_b200.calibrate import Chain
from rscm_b200.pandas_helpers import chain_to_dataframe, members_to_dataframe, quantiles_to_dataframe, target_from_dataframe


def make_chain(n_iter=6, n_walkers=4, thin=2):
    c = Chain(["a", "b"], thin)
    for it in range(n_iter):
        pos = np.array([[100 * it + w, -(100 * it + w)] for w in range(n_walkers)], dtype=float)
        c.push(pos, np.full(n_walkers, float(it)))
    return c


def test_chain_to_dataframe_labels_follow_the_samples():
    c = make_chain()
    df = chain_to_dataframe(c)
    assert list(df.columns) == ["a", "b", "log_prob"] and df.index.names == ["walker", "iteration"]
    assert len(df) == 3 * 4 and sorted(df.index.get_level_values("iteration").unique()) == [0, 2, 4]
    assert df.loc[(3, 4), "a"] == 403.0 and df.loc[(3, 4), "b"] == -403.0 and df.loc[(1, 2), "log_prob"] == 2.0
    assert len(chain_to_dataframe(c, discard=1)) == 2 * 4 and chain_to_dataframe(c, discard=1).index.get_level_values("iteration").min() == 2
    empty = chain_to_dataframe(c, discard=3)
    assert empty.empty and list(empty.columns) == ["a", "b", "log_prob"]


def test_target_from_dataframe():
    df = pd.DataFrame({"variable": ["T", "T", "OHC", "OHC"], "time": [1950, 2000, 1950, 2000], "value": [0.5, 1.0, 100.0, 200.0],
                       "uncertainty": [0.1, 0.15, 20.0, 30.0]})
    t = target_from_dataframe(df)
    assert sorted(t.variable_names()) == ["OHC", "T"] and t.total_observations() == 4
    flat = {(v, y): (x, s) for v, y, x, s in t._flat()}
    assert flat[("T", 2000.0)] == (1.0, 0.15) and flat[("OHC", 1950.0)] == (100.0, 20.0)
    rel = target_from_dataframe(df.drop(columns="uncertainty"), relative_error=0.1)
    assert {(v, y): s for v, y, _, s in rel._flat()}[("OHC", 2000.0)] == pytest.approx(20.0)
    with pytest.raises(ValueError, match="'variable' column"):
        target_from_dataframe(df.drop(columns="variable"))
    with pytest.raises(ValueError, match="No uncertainty information"):
        target_from_dataframe(df.drop(columns="uncertainty"))


def test_ensemble_frames():
    times = np.arange(1750.0, 1753.0)
    q = {"Surface Temperature": np.arange(2 * 3 * 2, dtype=float).reshape(2, 3, 2), "Box": np.zeros((2, 3, 4, 2))}
    df = quantiles_to_dataframe(q, [0.05, 0.95], times, ["low", "high"])
    assert list(df.columns) == [0.05, 0.95] and df.index.names == ["variable", "region", "scenario", "time"]
    assert df.loc[("Surface Temperature", 0, "high", 1751.0), 0.95] == q["Surface Temperature"][1, 1, 1] and len(df) == 3 * 2 + 3 * 4 * 2
    out = {"Surface Temperature": np.arange(3 * 6, dtype=float).reshape(3, 6)}   # T=3, S=2, M=3
    wide = members_to_dataframe(out, times, n_members=3)
    assert wide.shape == (6, 3) and wide.loc[("Surface Temperature", 0, 1, 2), 1752.0] == out["Surface Temperature"][2, 1 * 3 + 2]
