"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: member sharding and the
all-gather of per-member log-posteriors.  The local evaluator is the CPU oracle here (tests may
use it; the product path evaluates on the GPU)."""

import os
import socket

import numpy as np
import pytest

from rscm_b200 import _ffi
from rscm_b200 import synthetic as syn
from rscm_b200.dist import member_shard

from .helpers import oracle_bindings, oracle_from_builder


def test_member_shards_cover_ragged_sizes():
    for M in (1, 7, 8, 1000, 262144, 1048577):
        for G in (1, 2, 4, 8):
            blocks = [member_shard(M, r, G) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == M
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, M, ret):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rscm_b200.dist import ShardedLogPosterior

        b, binds, params, scen = syn.config2(M=M)
        m = oracle_from_builder(b)
        years = b._time_axis.values()
        obs = [("Surface Temperature", float(y), 0.5, 0.2) for y in years[100:271:10]]
        priors = [(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.TWO_LAYER_RANGES.values()]
        sc = np.stack([scen[0]["Effective Radiative Forcing"], 1.5 * scen[0]["Effective Radiative Forcing"]])

        def evaluator(p_local):
            return m.log_posterior_batch(oracle_bindings(b, binds), p_local, ["Effective Radiative Forcing"], sc, priors, obs, n_threads=1)

        sharded = ShardedLogPosterior(None, sc, evaluator=evaluator)
        full = sharded(params).numpy()
        if rank == 0:
            want = m.log_posterior_batch(oracle_bindings(b, binds), params, ["Effective Radiative Forcing"], sc, priors, obs, n_threads=1)
            ret["equal"] = bool(np.array_equal(full, want))
            ret["n"] = int(full.size)
        # every rank holds the same gathered vector
        t = torch.from_numpy(full.copy())
        dist.broadcast(t, src=0)
        ret[f"same{rank}"] = bool(np.array_equal(t.numpy(), full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("M", [64, 101])
def test_sharded_log_posterior_allgather_two_ranks(M):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, M, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret["equal"] and ret["n"] == 2 * M and ret["same0"] and ret["same1"]
