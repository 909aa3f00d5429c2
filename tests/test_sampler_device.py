"""Device-side stretch move: the Philox stream against the Random123 known answers, the propose / accept
kernels against a numpy restatement of sampler/moves.rs driven by the same stream (exact decisions), and the whole
DeviceEnsembleSampler against the host EnsembleSampler on the two-layer calibration problem of
tests/test_calibration_integration.py (statistical parity — the reference's thread_rng is not reproducible)."""

import numpy as np
import pytest

from rscm_b200 import _ffi, synthetic as syn
from rscm_b200.calibrate import (DeviceEnsembleSampler, EnsembleSampler, GaussianLikelihood, ModelRunner, ParameterSet, Target, Uniform,
                                 WalkerInit)

M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c, k):
    """Philox4x32-10 (Salmon et al., SC'11) on arrays of counters: c = 4 uint64 arrays holding 32-bit words, k = 2 words."""
    c = [np.asarray(x, dtype=np.uint64) for x in c]
    k = [np.uint64(x) for x in k]
    m0, m1, w0, w1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & M32, p1 & M32, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & M32, p0 & M32]
        k = [(k[0] + w0) & M32, (k[1] + w1) & M32]
    return c


def u01(hi, lo):
    return ((((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def draws(walkers, step, purpose, seed):
    w = np.asarray(walkers, dtype=np.uint64)
    r = philox4x32_10([w & M32, w >> np.uint64(32), np.full_like(w, step), np.full_like(w, purpose)], [seed & 0xFFFFFFFF, seed >> 32])
    return u01(r[0], r[1]), u01(r[2], r[3])


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
           ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
           ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1])]
    for c, k, want in kat:
        assert [int(x[0]) for x in philox4x32_10([[v] for v in c], k)] == want


def test_stretch_factor_distribution_of_the_stream():
    """moves.rs tests :150-200: z in [1/a, a], E[z] of g(z) ~ 1/sqrt(z); and the draws are uniform."""
    a = 2.0
    u, v = draws(np.arange(200_000), 7, 0, 0x1234_5678_9ABC_DEF0)
    z = ((a - 1.0) * u + 1.0) ** 2 / a
    assert z.min() >= 1.0 / a and z.max() <= a
    assert abs(z.mean() - (a * a + a + 1.0) / (3.0 * a)) < 5e-3          # E[z] = (a^2 + a + 1)/(3a)
    assert abs(u.mean() - 0.5) < 3e-3 and abs(v.mean() - 0.5) < 3e-3 and abs(np.corrcoef(u, v)[0, 1]) < 1e-2


def two_layer_problem(seed=11):
    """Synthetic observations of the two-layer model at known parameters (tests/test_calibration_integration.py:37-70)."""
    b, binds, _, scen = syn.config2(M=4)
    names = ["lambda0", "efficacy"]
    binds = {k: binds[k] for k in names}
    truth = np.array([[1.1, 1.3]])
    runner = ModelRunner(b, binds, ["Surface Temperature"], scenarios=None)
    runner._scenarios = runner.ensemble.pack_scenarios(scen)
    t_true = runner.run_batch_arrays(truth)["Surface Temperature"][:, 0]
    years = syn.time_axis().values()
    rng = np.random.default_rng(seed)
    target = Target()
    for i in range(100, 271, 10):
        target.add_observation("Surface Temperature", float(years[i]), float(t_true[i] + 0.05 * rng.standard_normal()), 0.05)
    params = ParameterSet().add("lambda0", Uniform(0.6, 1.8)).add("efficacy", Uniform(0.8, 2.0))
    return params, runner, target


def restated_half_update(pos, logp, lp_of, a0, c0, half, a, seed, step):
    """update_group (ensemble.rs:489-546) with the device's random stream."""
    w = np.arange(a0, a0 + half)
    u, v = draws(w, step, 0, seed)
    z = ((a - 1.0) * u + 1.0) ** 2 / a
    j = np.minimum((v * half).astype(np.int64), half - 1) + c0
    prop = pos[j] + z[:, None] * (pos[w] - pos[j])
    lp_new = lp_of(prop)
    with np.errstate(over="ignore", invalid="ignore"):
        p = np.where(np.isfinite(lp_new), np.fmin(np.exp((pos.shape[1] - 1.0) * np.log(z) + (lp_new - logp[w])), 1.0), 0.0)
    acc = draws(w, step, 1, seed)[0] < p
    pos[w[acc]] = prop[acc]
    logp[w[acc]] = lp_new[acc]
    return int(acc.sum())


@pytest.mark.gpu
def test_device_moves_follow_the_restated_update_exactly():
    params, runner, target = two_layer_problem()
    seed, W, iters = 0x5EED_0123_4567, 64, 6
    dev = DeviceEnsembleSampler(params, runner, GaussianLikelihood(), target, stretch=2.0, seed=3)
    init = WalkerInit.explicit(params.sample_random(W, np.random.default_rng(5)))
    chain = dev.run(iters, init, n_walkers=W, seed=seed)
    pos = init.positions.copy()
    logp = dev.log_posterior_batch(pos)
    n_acc = 0
    for it in range(iters):
        for hidx, (a0, c0) in enumerate(((0, W // 2), (W // 2, 0))):
            n_acc += restated_half_update(pos, logp, dev.log_posterior_batch, a0, c0, W // 2, 2.0, seed, 2 * it + hidx)
        np.testing.assert_allclose(chain._samples[it], pos, rtol=1e-12, atol=0)
        np.testing.assert_allclose(chain._log_probs[it], logp, rtol=1e-9)
    assert dev.acceptance_rate == n_acc / (iters * W) and 0.1 < dev.acceptance_rate < 0.95
    # reproducible from the seed; a different seed moves differently
    again = dev.run(iters, init, n_walkers=W, seed=seed)
    assert np.array_equal(again._samples[-1], chain._samples[-1])
    other = dev.run(iters, init, n_walkers=W, seed=seed + 1)
    assert not np.array_equal(other._samples[-1], chain._samples[-1])


@pytest.mark.gpu
def test_device_sampler_matches_host_sampler_statistically():
    params, runner, target = two_layer_problem()
    W, iters, burn = 128, 400, 100
    init = WalkerInit.ball([1.0, 1.4], 0.2)
    host = EnsembleSampler(params, runner, GaussianLikelihood(), target, seed=1).run(iters, init, n_walkers=W)
    dsam = DeviceEnsembleSampler(params, runner, GaussianLikelihood(), target, seed=2)
    devc = dsam.run(iters, init, n_walkers=W, thin=1)
    h, d = host.flat_samples(burn), devc.flat_samples(burn)
    assert h.shape == d.shape == ((iters - burn) * W, 2)
    for j in range(2):
        sd = h[:, j].std()
        assert abs(h[:, j].mean() - d[:, j].mean()) < 0.15 * sd     # ~ 38k correlated samples each: well within
        assert 0.8 < d[:, j].std() / sd < 1.25
    assert abs(np.corrcoef(h.T)[0, 1] - np.corrcoef(d.T)[0, 1]) < 0.1
    assert 0.2 < dsam.acceptance_rate < 0.9
    assert abs(d[:, 0].mean() - 1.1) < 0.15 and abs(d[:, 1].mean() - 1.3) < 0.25   # recovers the truth


@pytest.mark.gpu
def test_thinning_progress_and_argument_errors():
    params, runner, target = two_layer_problem()
    s = DeviceEnsembleSampler(params, runner, GaussianLikelihood(), target, seed=4)
    seen = []
    chain = s.run(7, WalkerInit.from_prior(), thin=3, n_walkers=32, progress=lambda info: seen.append((info.iteration, info.acceptance_rate)))
    assert len(chain) == 3 and chain.total_iterations == 7 and [i for i, _ in seen] == list(range(7))
    assert chain.flat_samples().shape == (3 * 32, 2) and np.isfinite(chain.flat_log_probs()).any()
    with pytest.raises(ValueError, match="even"):
        s.run(1, WalkerInit.from_prior(), n_walkers=31)
    import torch
    t = torch.zeros(4, dtype=torch.float64, device="cuda")
    with pytest.raises(_ffi.EngineError, match="stretch parameter"):
        _ffi.check(_ffi.lib.rscm_b200_stretch_propose(t.data_ptr(), 2, 1, 0, 1, 1, 1, 1.0, 1, 0, t.data_ptr(), 1, t.data_ptr(), None))


@pytest.mark.gpu
def test_graph_replay_equals_eager_iterations_and_single_rank_comm_entry_points():
    """rscm_b200_sampler_iterate: the iterations replayed from the captured CUDA graph (device-side iteration counter) give
    the chain the eagerly enqueued iterations give, bit for bit; with one rank the sharded entry points reduce to the plain
    ones (no NCCL needed)."""
    import torch

    from rscm_b200.dist import Comm

    params, runner, target = two_layer_problem()
    s = DeviceEnsembleSampler(params, runner, GaussianLikelihood(), target, seed=9)
    init = WalkerInit.explicit(params.sample_random(96, np.random.default_rng(8)))
    g = s.run(9, init, n_walkers=96, seed=1234, thin=2, use_graph=True)
    e = s.run(9, init, n_walkers=96, seed=1234, thin=2, use_graph=False)
    assert len(g) == len(e) == 5
    for a, b in zip(g._samples, e._samples):
        assert np.array_equal(a, b)
    assert np.array_equal(np.asarray(g._log_probs), np.asarray(e._log_probs))
    comm = Comm.single()
    assert comm.world == 1 and not comm.peer_access and comm.shard(10) == (0, 10)
    x = torch.arange(5, dtype=torch.float64, device="cuda")
    y = torch.empty(5, dtype=torch.float64, device="cuda")
    comm.allgather(x, y)
    ens = runner.ensemble
    p = torch.from_numpy(np.ascontiguousarray(init.positions.T)).cuda()
    sc = torch.from_numpy(runner._scenarios).cuda()
    lp = comm.symmetric_empty(96)
    comm.log_posterior_sharded(ens, p, sc, lp, M=96, S=1, layout=0)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), np.arange(5.0))
    assert np.array_equal(lp.cpu().numpy(), s.log_posterior_batch(init.positions))
