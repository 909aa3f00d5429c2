"""The engine's own fp64 exp / log / pow (rscm_b200/csrc/components.cuh: constant-bank coefficients, table-driven exp) against
the host library over the ranges the components use and over the special cases that take the library fall-back."""

import numpy as np
import pytest

from rscm_b200 import _ffi


def device(op, x):
    import torch
    d_x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    d_y = torch.empty(d_x.shape[0], dtype=torch.float64, device="cuda")     # op 2 (pow): x is [n][2] = (base, exponent)
    _ffi.check(_ffi.lib.rscm_b200_device_math(op, d_x.data_ptr(), d_y.numel(), d_y.data_ptr(), None))
    torch.cuda.synchronize()
    return d_y.cpu().numpy()


def ulp_error(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


@pytest.mark.gpu
def test_exp_within_one_ulp_and_special_cases():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-2.0, 2.0, 400_000), rng.uniform(-40.0, 40.0, 200_000), rng.uniform(-699.0, 699.0, 200_000),
                        np.linspace(-1e-8, 1e-8, 1001), np.arange(-64, 65) * (np.log(2.0) / 64.0), [0.0, -0.0]])
    got, want = device(0, x), np.exp(x)
    assert np.max(ulp_error(got, want)) <= 2.0      # <= 1 ulp from the exact value each
    assert device(0, np.array([0.0]))[0] == 1.0
    sp = np.array([700.0, 709.7, 710.0, -700.0, -745.0, -800.0, np.inf, -np.inf, np.nan, 1e300, -1e300])
    with np.errstate(over="ignore", under="ignore"):
        ws = np.exp(sp)
    gs = device(0, sp)
    assert np.array_equal(np.isnan(gs), np.isnan(ws))
    ok = ~np.isnan(ws)
    assert np.all((gs[ok] == ws[ok]) | (ulp_error(gs[ok], ws[ok]) <= 1.0))


@pytest.mark.gpu
def test_log_within_one_ulp_and_special_cases():
    rng = np.random.default_rng(2)
    y = np.concatenate([rng.uniform(0.5, 8.0, 400_000), 1.0 + rng.uniform(-1e-3, 1e-3, 100_000), np.exp(rng.uniform(-700.0, 700.0, 200_000)),
                        2.0 ** np.arange(-1000, 1001, 7), np.sqrt(2.0) * (1.0 + np.linspace(-1e-12, 1e-12, 101)), [1.0, 2.0, 0.5, 278.0, 556.0]])
    got, want = device(1, y), np.log(y)
    nz = want != 0.0
    assert np.max(ulp_error(got[nz], want[nz])) <= 2.0
    assert np.all(got[~nz] == 0.0)             # log(1) is exactly 0: CO2ERF is exactly 0 at the pre-industrial concentration
    sp = np.array([0.0, -0.0, -1.0, np.inf, np.nan, 5e-324, 2.2e-308, 1e-310])
    with np.errstate(divide="ignore", invalid="ignore"):
        ws = np.log(sp)
    gs = device(1, sp)
    assert np.array_equal(np.isnan(gs), np.isnan(ws))
    ok = ~np.isnan(ws)
    assert np.all((gs[ok] == ws[ok]) | (ulp_error(gs[ok], ws[ok]) <= 1.0))


@pytest.mark.gpu
def test_pow_within_two_ulp_and_special_cases():
    rng = np.random.default_rng(3)
    # the components' ranges: burden ratios >= 1 with lifetime exponents, the CH4 x N2O products of the overlap terms
    # (1e5 .. 1e7 ppb^2, exponents 0.75 and 1.52), EESC ratios with the chlorine exponent; then a wide sweep
    base = np.concatenate([rng.uniform(1.0, 12.0, 300_000), rng.uniform(1e4, 1e8, 200_000), rng.uniform(1e-3, 1.0, 100_000),
                           np.exp(rng.uniform(-300.0, 300.0, 200_000)), 1.0 + rng.uniform(-1e-6, 1e-6, 50_000)])
    expo = np.concatenate([rng.uniform(-1.0, 2.0, 300_000), rng.choice([0.75, 1.52], 200_000), rng.uniform(0.5, 2.5, 100_000),
                           rng.uniform(-2.0, 2.0, 200_000), rng.uniform(-50.0, 50.0, 50_000)])
    got, want = device(2, np.stack([base, expo], axis=1)), np.power(base, expo)
    assert np.max(ulp_error(got, want)) <= 2.0
    # exact cases and the library fall-back
    sp = np.array([[1.0, 3.7], [5.0, 0.0], [np.nan, 0.0], [2.0, 10.0], [4.0, 0.5], [0.0, 2.0], [0.0, -1.0], [-8.0, 3.0], [-8.0, 0.5],
                   [np.inf, 2.0], [np.inf, -2.0], [3.0, np.inf], [0.5, np.inf], [3.0, np.nan], [1e300, 5.0], [1e-300, 5.0],
                   [5e-324, 0.5], [10.0, 308.0], [10.0, -320.0], [2.0, 1023.0]])
    with np.errstate(all="ignore"):
        ws = np.power(sp[:, 0], sp[:, 1])
    gs = device(2, sp)
    assert np.array_equal(np.isnan(gs), np.isnan(ws))
    ok = ~np.isnan(ws)
    assert np.all((gs[ok] == ws[ok]) | (ulp_error(gs[ok], ws[ok]) <= 2.0))
    assert gs[0] == 1.0 and gs[1] == 1.0 and gs[2] == 1.0      # 1^y, x^0 and NaN^0 are exactly 1
