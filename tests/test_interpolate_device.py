"""Scenario ingestion on the device: rscm_b200_interpolate_device against the host restatement of
Timeseries::interpolate_into (rscm_b200/core.py, itself checked against the reference's strategy vectors in
tests/test_host_logic.py) — bit for bit, for the three strategies, scalar and grid series, boundary snapping and
extrapolation beyond both ends; then a coarse scenario file resampled on the device and fed straight to the ensemble kernel."""

import numpy as np
import pytest

from rscm_b200 import synthetic as syn
from rscm_b200.core import InterpolationStrategy, TimeAxis, Timeseries, _interp, interpolate_device

STRATEGIES = [InterpolationStrategy.Linear, InterpolationStrategy.Next, InterpolationStrategy.Previous]


def host(strategy, t_old, values, t_new):
    """values [..., K] -> [..., T] with the host routine, one target at a time (as interpolate_into does)."""
    v = np.asarray(values, dtype=float)
    flat = v.reshape(-1, v.shape[-1])
    out = np.empty((flat.shape[0], t_new.size))
    for s in range(flat.shape[0]):
        for i, t in enumerate(t_new):
            out[s, i] = _interp(strategy, t_old, flat[s], float(t))
    return out.reshape(v.shape[:-1] + (t_new.size,))


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", STRATEGIES)
def test_device_interpolation_equals_host_bitwise(strategy):
    rng = np.random.default_rng(3)
    t_old = np.sort(np.concatenate([[1750.0, 2100.0], rng.uniform(1750.0, 2100.0, 60)]))
    vals = rng.normal(size=(5, 3, t_old.size)).cumsum(axis=-1)
    t_new = np.concatenate([
        np.arange(1700.0, 2151.0, 1.0),                       # beyond both ends: extrapolation
        t_old,                                                # exact boundary hits
        t_old * (1.0 + 3e-9), t_old * (1.0 - 3e-9),           # inside the is_close! tolerance (1e-8 relative)
        t_old * (1.0 + 3e-8), t_old * (1.0 - 3e-8),           # just outside it
        rng.uniform(1740.0, 2110.0, 500),
    ])
    got = interpolate_device(t_old, vals, t_new, strategy).cpu().numpy()
    want = host(strategy, t_old, vals, t_new)
    assert got.shape == want.shape == (5, 3, t_new.size)
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", STRATEGIES)
def test_device_interpolation_of_grid_series_and_nan_values(strategy):
    rng = np.random.default_rng(4)
    t_old = np.arange(1850.0, 2101.0, 10.0)
    vals = rng.normal(size=(7, t_old.size, 4))
    vals[2, 5, 1] = np.nan                                     # a missing value propagates into its segments only
    t_new = np.arange(1840.0, 2111.0, 1.0)
    got = interpolate_device(t_old, vals, t_new, strategy).cpu().numpy()      # [..., K, R] -> [..., T, R]
    want = np.moveaxis(host(strategy, t_old, np.moveaxis(vals, -1, -2), t_new), -2, -1)
    assert got.shape == (7, t_new.size, 4) and np.array_equal(got, want, equal_nan=True)


@pytest.mark.gpu
def test_coarse_scenarios_resampled_on_the_device_drive_the_same_run():
    """Emission scenarios given every ten years: resampled onto the annual model axis by the device kernel and handed to
    rscm_b200_run_device without a host pass; the run equals the one fed by the host's Timeseries.interpolate_into."""
    import torch

    axis = syn.time_axis(1750, 2100)
    b = syn.coupled_builder(axis=axis)
    ens = b.build_ensemble().bind_parameters(syn.COUPLED_BINDINGS)
    ens.select_outputs(["Atmospheric Concentration|CO2", "Surface Temperature"])
    coarse_t = np.arange(1750.0, 2101.0, 10.0)
    coarse = syn.emission_scenarios(coarse_t, 4)                          # [S][K]
    caxis = TimeAxis.from_values(coarse_t)
    host_scen = [{"Emissions|CO2|Anthropogenic": Timeseries(coarse[s], caxis, "GtC / yr", InterpolationStrategy.Linear)
                  .interpolate_into(axis)._values[:, 0]} for s in range(4)]
    params = syn.uniform_params(syn.COUPLED_RANGES, 200, 8)
    want = ens.run(params, ens.pack_scenarios(host_scen))
    d_scen = interpolate_device(coarse_t, torch.from_numpy(coarse).cuda(), axis.values(), InterpolationStrategy.Linear)   # [S][T] = [S][n_exo][T][1]
    d_p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    d_o = torch.empty((ens.output_rows, 4 * 200), dtype=torch.float64, device="cuda")
    ens.run_device(d_p, d_scen, d_o, layout=0, S=4)
    torch.cuda.synchronize()
    assert np.array_equal(d_o.cpu().numpy(), want, equal_nan=True)
