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
