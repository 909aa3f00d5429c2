"""Multi-GPU parity (needs >= 2 CUDA devices; `gpurun --gpus 2`): per-member results are bit-identical
whatever the number of GPUs the ensemble is sharded over, and every rank ends up with all
log-posteriors after the NCCL all-gather."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, M, ret):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from rscm_b200 import _ffi
        from rscm_b200 import synthetic as syn
        from rscm_b200.dist import ShardedLogPosterior, member_shard

        b, binds, params, scen = syn.config3(M=M, S=2)
        ens = b.build_ensemble(device=rank).bind_parameters(binds)
        sc = ens.pack_scenarios(scen)
        obs = [("Surface Temperature", float(y), 0.4, 0.2) for y in range(1900, 2021, 5)]
        priors = [(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.COUPLED_RANGES.values()]
        ens.set_target(obs).set_priors(priors)
        full = ShardedLogPosterior(ens, sc)(params).cpu().numpy()
        # timeseries outputs of this rank's member block
        lo, hi = member_shard(M, rank, world)
        ens.select_outputs(["Surface Temperature"])
        part = ens.run(params[lo:hi], sc)
        ret[f"lp{rank}"] = full
        ret[f"ts{rank}"] = (lo, hi, part)
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharding_is_bit_identical_to_one_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from rscm_b200 import _ffi
    from rscm_b200 import synthetic as syn

    M = 1001
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, M, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    b, binds, params, scen = syn.config3(M=M, S=2)
    ens = b.build_ensemble(device=0).bind_parameters(binds)
    sc = ens.pack_scenarios(scen)
    obs = [("Surface Temperature", float(y), 0.4, 0.2) for y in range(1900, 2021, 5)]
    ens.set_target(obs).set_priors([(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.COUPLED_RANGES.values()])
    want = ens.log_posterior(params, sc)
    assert np.array_equal(ret["lp0"], want) and np.array_equal(ret["lp1"], want)
    ens.select_outputs(["Surface Temperature"])
    full = ens.run(params, sc).reshape(351, 2, M)
    for r in range(2):
        lo, hi, part = ret[f"ts{r}"]
        assert np.array_equal(part.reshape(351, 2, hi - lo), full[:, :, lo:hi])


def _sampler_worker(rank, world, port, ret, W):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from rscm_b200.calibrate import DeviceEnsembleSampler, GaussianLikelihood, WalkerInit

        from tests.test_sampler_device import two_layer_problem

        params, runner, target = two_layer_problem()
        s = DeviceEnsembleSampler(params, runner, GaussianLikelihood(), target, seed=100 + rank)  # rank 0's seed and walkers win
        chain = s.run(12, WalkerInit.from_prior(), n_walkers=W, seed=(77 if rank == 0 else 78), distributed=True)
        ret[f"pos{rank}"] = chain._samples[-1]
        ret[f"first{rank}"] = chain._samples[0]
        ret[f"acc{rank}"] = s.acceptance_rate
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W", [66, 64])  # ragged shards (padded all-gather) and equal shards (gathered straight into place)
def test_two_gpu_sampler_replicas_stay_identical_and_match_one_gpu(W):
    """Replicated walker state + sharded log-posterior + all-gather: both ranks hold the same chain, and it
    is the chain one GPU produces from the same seed and initial ensemble (per-member results do not depend on sharding)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, 2, port, ret, W)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert np.array_equal(ret["pos0"], ret["pos1"]) and ret["acc0"] == ret["acc1"] and 0.05 < ret["acc0"] < 0.95
    from rscm_b200.calibrate import DeviceEnsembleSampler, GaussianLikelihood, WalkerInit

    from tests.test_sampler_device import two_layer_problem

    params, runner, target = two_layer_problem()
    s = DeviceEnsembleSampler(params, runner, GaussianLikelihood(), target, seed=100)
    one = s.run(12, WalkerInit.from_prior(), n_walkers=W, seed=77)
    assert np.array_equal(one._samples[-1], ret["pos0"]) and np.array_equal(one._samples[0], ret["first0"])


# ---- the collective behind the C ABI, driven through ctypes with a file rendezvous (no torch.distributed): what a Rust host does ----
def _cabi_worker(rank, world, idfile, ret, no_p2p):
    if no_p2p:
        os.environ["RSCM_B200_NO_P2P"] = "1"
    import torch

    torch.cuda.set_device(rank)
    from rscm_b200 import _ffi
    from rscm_b200 import synthetic as syn
    from rscm_b200.dist import Comm

    comm = Comm.from_file(idfile, rank, world, device=rank)
    ret[f"p2p{rank}"] = comm.peer_access
    # 1. rscm_b200_allgather_f64
    n = 1000
    local = torch.arange(n, dtype=torch.float64, device="cuda") + 10000.0 * rank
    out = torch.empty(world * n, dtype=torch.float64, device="cuda")
    comm.allgather(local, out)
    torch.cuda.synchronize()
    ret[f"ag{rank}"] = out.cpu().numpy()
    # 2. rscm_b200_logpost_sharded_device: S = 2 scenarios, ragged M, global parameter matrix on every rank
    M = 1001
    b, binds, params, scen = syn.config3(M=M, S=2)
    ens = b.build_ensemble(device=rank).bind_parameters(binds)
    sc = torch.from_numpy(ens.pack_scenarios(scen)).cuda()
    obs = [("Surface Temperature", float(y), 0.4, 0.2) for y in range(1900, 2021, 5)]
    ens.set_target(obs).set_priors([(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.COUPLED_RANGES.values()])
    d_p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    bufs = [comm.symmetric_empty(2 * M), comm.symmetric_empty(2 * M)]
    plain = torch.empty(2 * M, dtype=torch.float64, device="cuda")   # not symmetric memory: NCCL carries this one
    for k in range(4):                                                # alternate destinations; repeated epochs
        comm.log_posterior_sharded(ens, d_p, sc, bufs[k & 1], M=M, S=2, layout=0)
    comm.log_posterior_sharded(ens, d_p, sc, plain, M=M, S=2, layout=0)
    torch.cuda.synchronize()
    comm.check()
    ret[f"lp{rank}"] = [bufs[0].cpu().numpy(), bufs[1].cpu().numpy(), plain.cpu().numpy()]
    # layout 1 ([M][cols], the reference's &[Vec<f64>]) through the same entry point
    d_p1 = torch.from_numpy(np.ascontiguousarray(params)).cuda()
    comm.log_posterior_sharded(ens, d_p1, sc, bufs[0], M=M, S=2, layout=1)
    torch.cuda.synchronize()
    ret[f"lp1_{rank}"] = bufs[0].cpu().numpy()
    comm.close()


@pytest.mark.parametrize("no_p2p", [False, True])
def test_c_abi_collectives_two_ranks(no_p2p, tmp_path):
    """rscm_b200_comm_init / allgather_f64 / logpost_sharded_device from two processes that share nothing but a file: every
    rank ends with the log-posteriors one GPU computes, bit for bit, through the fused peer-store path and through NCCL."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from rscm_b200 import _ffi
    from rscm_b200 import synthetic as syn

    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_cabi_worker, args=(r, 2, str(tmp_path / "nccl_id"), ret, no_p2p)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    want_ag = np.concatenate([np.arange(1000.0), np.arange(1000.0) + 10000.0])
    assert np.array_equal(ret["ag0"], want_ag) and np.array_equal(ret["ag1"], want_ag)
    M = 1001
    b, binds, params, scen = syn.config3(M=M, S=2)
    ens = b.build_ensemble(device=0).bind_parameters(binds)
    obs = [("Surface Temperature", float(y), 0.4, 0.2) for y in range(1900, 2021, 5)]
    ens.set_target(obs).set_priors([(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.COUPLED_RANGES.values()])
    want = ens.log_posterior(params, ens.pack_scenarios(scen))
    for r in range(2):
        for got in ret[f"lp{r}"]:
            assert np.array_equal(got, want)
        assert np.array_equal(ret[f"lp1_{r}"], want)
    if no_p2p:
        assert not ret["p2p0"] and not ret["p2p1"]
    print("peer access:", ret["p2p0"], ret["p2p1"])
