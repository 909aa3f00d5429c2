"""Across-member quantiles on the device against numpy.nanquantile — bit for bit: exact order statistics
(radix selection) and numpy's own interpolation formula; NaNs skipped, ties, constant and all-NaN segments, infinities,
ragged sizes; and the Ensemble.run_quantiles surface against the quantiles of the full member output."""

import numpy as np
import pytest

from rscm_b200 import _ffi, synthetic as syn

pytestmark = pytest.mark.gpu
Q5 = [0.05, 0.17, 0.5, 0.83, 0.95]


def device_quantiles(data, S, M, q):
    """data [rows][S*M] numpy -> [len(q)][rows][S] through the C ABI."""
    import ctypes as C

    import torch
    d = torch.from_numpy(np.ascontiguousarray(data)).cuda()
    rows = data.shape[0]
    res = torch.empty((len(q), rows, S), dtype=torch.float64, device="cuda")
    qs = (C.c_double * len(q))(*q)
    _ffi.check(_ffi.lib.rscm_b200_member_quantiles(d.data_ptr(), rows, S, M, qs, len(q), res.data_ptr(), None))
    torch.cuda.synchronize()
    return res.cpu().numpy()


def numpy_quantiles(data, S, M, q):
    blk = data.reshape(data.shape[0], S, M)
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return np.nanquantile(blk, q, axis=2, method="linear")


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("M", [1, 2, 3, 31, 1000, 5003])
def test_random_blocks_with_nans_ties_and_constants(M):
    rng = np.random.default_rng(M)
    rows, S = 23, 3
    data = rng.standard_normal((rows, S * M)) * np.exp(rng.uniform(-20, 20, size=(rows, 1)))
    data[1] = np.round(data[1] / np.abs(data[1]).max() * 3.0)            # heavy ties: a handful of distinct values
    data[2] = 4.25                                                        # constant
    data[3] = np.nan                                                      # all NaN -> NaN
    data[4, rng.random(S * M) < 0.3] = np.nan                             # NaNs are skipped
    data[5, : S * M // 2] = 0.0                                           # signed zeros and a sign change
    data[5, S * M // 2:] = -0.0
    data[6] = np.abs(data[6])                                             # single sign / narrow exponent range
    data[7, ::7] = np.inf
    data[8, ::5] = -np.inf
    data[9] = rng.integers(-3, 4, size=S * M) * 1e-310                    # subnormals
    for q in (Q5, [0.0, 1.0], [0.5], [0.999, 0.001, 0.25]):
        got, want = device_quantiles(data, S, M, q), numpy_quantiles(data, S, M, q)
        assert got.shape == want.shape == (len(q), rows, S)
        if not same(got, want):
            bad = np.argwhere(~((got == want) | (np.isnan(got) & np.isnan(want))))
            raise AssertionError(f"M={M} q={q}: first mismatch at {bad[0]}: {got[tuple(bad[0])]!r} vs {want[tuple(bad[0])]!r}")


def test_headline_member_count_smooth_and_degenerate():
    """262 144 members per segment: smooth data takes the histogram-histogram-collect route, a two-valued segment forces the
    selection through all 64 key bits."""
    rng = np.random.default_rng(7)
    M, S = 262_144, 2
    data = np.empty((4, S * M))
    data[0] = 1.5 + 0.8 * rng.standard_normal(S * M)
    data[1] = rng.lognormal(0.0, 2.0, S * M)
    data[2] = np.where(rng.random(S * M) < 0.5, 1.0, np.nextafter(1.0, 2.0))
    data[3] = rng.integers(0, 3, S * M).astype(float)
    assert same(device_quantiles(data, S, M, Q5), numpy_quantiles(data, S, M, Q5))


def test_run_quantiles_matches_quantiles_of_the_member_outputs():
    b, binds, params, scen = syn.config3(M=3000, S=2)
    ens = b.build_ensemble().bind_parameters(binds)
    sc = ens.pack_scenarios(scen)
    ens.select_outputs(["Surface Temperature", "Atmospheric Concentration|CO2", "Effective Radiative Forcing"], t_start=100, t_step=10)
    full = ens.split_outputs(ens.run(params, sc))
    qs = ens.run_quantiles(params, sc, Q5 + [0.0, 1.0])
    for name, series in full.items():
        nt = series.shape[0]
        want = numpy_quantiles(series.reshape(nt, -1), 2, 3000, Q5 + [0.0, 1.0])
        assert qs[name].shape == (7, nt, 2) and same(qs[name], want), name
    t = qs["Surface Temperature"]
    assert np.all(np.diff(t[[5, 0, 1, 2, 3, 4, 6]], axis=0)[:, 1:] >= 0.0)   # ordered in q: min, 5 %, ..., 95 %, max


def test_argument_errors():
    import ctypes as C

    import torch
    d = torch.zeros(8, dtype=torch.float64, device="cuda")
    q6 = (C.c_double * 6)(*([0.5] * 6))
    with pytest.raises(_ffi.EngineError, match="1 to 5 quantiles"):
        _ffi.check(_ffi.lib.rscm_b200_member_quantiles(d.data_ptr(), 1, 1, 8, q6, 6, d.data_ptr(), None))
    bad = (C.c_double * 1)(1.5)
    with pytest.raises(_ffi.EngineError, match=r"range \[0, 1\]"):
        _ffi.check(_ffi.lib.rscm_b200_member_quantiles(d.data_ptr(), 1, 1, 8, bad, 1, d.data_ptr(), None))
