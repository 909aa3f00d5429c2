"""Multi-GPU plumbing for the ensemble path: one process per GPU (``torch.distributed``).

The path shards by member with no data-path exchange: rank r runs members
``[r*M/G, (r+1)*M/G)`` of every scenario (``ModelRunner::run_batch`` maps ``run`` over independent
rows, crates/rscm-calibrate/src/model_runner.rs:261-266) and keeps its timeseries outputs local.
The one collective is the all-gather of per-member log-posteriors (8 B per run) that lets every
rank's replica of the sampler advance identically — issued on the device buffer
right after the fused kernel, NCCL over NVLink on GPUs, gloo in the CPU tests.
"""

from __future__ import annotations

from typing import Callable

import numpy as np


def member_shard(M: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous member block of `rank`: [r*M/G, (r+1)*M/G) (integer arithmetic, covers ragged M)."""
    return (rank * M) // world, ((rank + 1) * M) // world


def allgather_members(local, M: int, group=None):
    """All-gather per-member values laid out [S, M_local] (or [M_local]) into [S, M] on every rank.

    Shards may be ragged (M not divisible by the world size): every rank pads to the largest
    shard so that a single fixed-size ``all_gather_into_tensor`` is issued."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    squeeze = local.dim() == 1
    loc = local.reshape(1, -1) if squeeze else local
    S = loc.shape[0]
    sizes = [member_shard(M, r, world)[1] - member_shard(M, r, world)[0] for r in range(world)]
    assert loc.shape[1] == sizes[rank], (loc.shape, sizes[rank])
    width = max(sizes)
    send = torch.full((S, width), float("nan"), dtype=loc.dtype, device=loc.device)
    send[:, : sizes[rank]] = loc
    recv = torch.empty((world * S, width), dtype=loc.dtype, device=loc.device)  # concatenated along dim 0
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, S, width)
    out = torch.cat([recv[r, :, : sizes[r]] for r in range(world)], dim=1)
    return out.reshape(-1) if squeeze else out


class ShardedLogPosterior:
    """``EnsembleSampler::log_posterior_batch`` over G GPUs: every rank is handed the same global
    parameter matrix [M, n_cols], evaluates its member block with the fused kernel, and all ranks
    end up with all S*M log-posteriors (run index = s*M + m as on one GPU)."""

    def __init__(self, ensemble, scenarios: np.ndarray | None, group=None, evaluator: Callable | None = None):
        import torch

        self.ens = ensemble
        self.group = group
        self._evaluator = evaluator
        self._scen_host = scenarios
        self._scen_dev = None
        if evaluator is None and scenarios is not None:
            self._scen_dev = torch.from_numpy(np.ascontiguousarray(scenarios)).cuda()

    def __call__(self, params: np.ndarray):
        import torch
        import torch.distributed as dist

        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        M = params.shape[0]
        lo, hi = member_shard(M, rank, world)
        S = 1 if self._scen_host is None else self._scen_host.shape[0]
        if self._evaluator is not None:  # CPU tests inject the local evaluator
            local = torch.from_numpy(np.asarray(self._evaluator(params[lo:hi])).reshape(S, hi - lo))
        else:
            p = torch.from_numpy(np.ascontiguousarray(params[lo:hi].T)).cuda()  # [cols][M_local]
            local = torch.empty((S, hi - lo), dtype=torch.float64, device="cuda")
            self.ens.log_posterior_device(p, self._scen_dev, local, layout=0, M=hi - lo, S=S if self._scen_dev is not None else 0)
        full = allgather_members(local, M, self.group)
        return full.reshape(-1)
