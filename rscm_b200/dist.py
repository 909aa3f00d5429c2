"""Multi-GPU plumbing for the ensemble path: one process per GPU (``torch.distributed``).

The path shards by member with no data-path exchange: rank r runs members
``[r*M/G, (r+1)*M/G)`` of every scenario (``ModelRunner::run_batch`` maps ``run`` over independent
rows, crates/rscm-calibrate/src/model_runner.rs:261-266) and keeps its timeseries outputs local.
The one collective is the all-gather of per-member log-posteriors (8 B per run) that lets every
rank's replica of the sampler advance identically — issued on the device buffer
right after the fused kernel, NCCL over NVLink on GPUs, gloo in the CPU tests.
"""

from __future__ import annotations

import ctypes as C
import os
import time
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
    end up with all S*M log-posteriors (run index = s*M + m as on one GPU).

    On GPUs this is ``rscm_b200_logpost_sharded_device`` (the member block is evaluated in place and the all-gather is
    fused into the kernel over peer memory, or NCCL); ``evaluator`` + a ``torch.distributed`` group is the CPU (gloo) test
    harness for the sharding logic."""

    def __init__(self, ensemble, scenarios: np.ndarray | None, group=None, evaluator: Callable | None = None, comm: "Comm | None" = None):
        self.ens = ensemble
        self.group = group
        self._evaluator = evaluator
        self._scen_host = scenarios
        self._scen_dev = None
        self._comm = comm
        self._bufs = {}
        self._flip = 0
        if evaluator is None:
            import torch
            if scenarios is not None:
                self._scen_dev = torch.from_numpy(np.ascontiguousarray(scenarios)).cuda()
            if self._comm is None:
                self._comm = Comm.from_torch(group, torch.cuda.current_device())

    def __call__(self, params: np.ndarray):
        import torch

        M = params.shape[0]
        S = 1 if self._scen_host is None else self._scen_host.shape[0]
        if self._evaluator is not None:  # CPU tests inject the local evaluator
            import torch.distributed as dist
            world = dist.get_world_size(self.group)
            rank = dist.get_rank(self.group)
            lo, hi = member_shard(M, rank, world)
            local = torch.from_numpy(np.asarray(self._evaluator(params[lo:hi])).reshape(S, hi - lo))
            return allgather_members(local, M, self.group).reshape(-1)
        p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()  # the GLOBAL matrix [cols][M]; the kernel takes the block offset
        n = S * M
        if n not in self._bufs:   # two symmetric destinations, used alternately (include/rscm_b200.h)
            self._bufs[n] = (self._comm.symmetric_empty(n), self._comm.symmetric_empty(n))
        out = self._bufs[n][self._flip]
        self._flip ^= 1
        self._comm.log_posterior_sharded(self.ens, p, self._scen_dev, out, M=M, S=S if self._scen_dev is not None else 0, layout=0)
        torch.cuda.synchronize()
        return out.clone()


class _DeviceBlock:
    """A raw device allocation exposed through ``__cuda_array_interface__`` (so ``torch.as_tensor`` can view it)."""

    def __init__(self, ptr: int, n: int, owner):
        self.ptr, self.n, self._owner = ptr, n, owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


class Comm:
    """The C-ABI communicator (include/rscm_b200.h: rscm_b200_comm_*): NCCL for the generic collectives plus peer-mapped
    symmetric buffers for the fused log-posterior + all-gather kernel.  One per process; every method that talks to other
    ranks is collective."""

    def __init__(self, unique_id: bytes | None, rank: int, world: int, device: int = -1):
        from . import _ffi
        self._ffi = _ffi
        self._h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, _ffi.UNIQUE_ID_BYTES) if unique_id is not None else None
        _ffi.check_comm(_ffi.lib.rscm_b200_comm_init(buf, rank, world, device, C.byref(self._h)))
        self.rank, self.world = rank, world

    @staticmethod
    def unique_id() -> bytes:
        from . import _ffi
        buf = C.create_string_buffer(_ffi.UNIQUE_ID_BYTES)
        _ffi.check_comm(_ffi.lib.rscm_b200_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch(cls, group=None, device: int = -1) -> "Comm":
        """Rank / world size from ``torch.distributed``; the unique id travels through ``broadcast_object_list``."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(box[0], rank, world, device)

    @classmethod
    def from_file(cls, path: str, rank: int, world: int, device: int = -1, timeout: float = 120.0) -> "Comm":
        """Rendezvous through a file every rank can see (what a Rust host would do with its own launcher): rank 0 writes the
        unique id atomically, the others wait for it."""
        if world == 1:
            return cls(None, 0, 1, device)
        if rank == 0:
            uid = cls.unique_id()
            with open(path + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(path + ".tmp", path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > timeout:
                    raise TimeoutError("no unique id from rank 0")
                time.sleep(0.01)
            uid = open(path, "rb").read()
        return cls(uid, rank, world, device)

    @classmethod
    def single(cls, device: int = -1) -> "Comm":
        return cls(None, 0, 1, device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._ffi.lib.rscm_b200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def peer_access(self) -> bool:
        return bool(self._ffi.lib.rscm_b200_comm_peer_access(self._h))

    def shard(self, M: int) -> tuple[int, int]:
        lo, hi = C.c_int64(), C.c_int64()
        self._ffi.check_comm(self._ffi.lib.rscm_b200_comm_shard(self._h, M, C.byref(lo), C.byref(hi)), self._h)
        return lo.value, hi.value

    def symmetric_empty(self, n: int):
        """``n`` zeroed float64 elements of symmetric device memory as a torch tensor (lives as long as the communicator)."""
        import torch
        p = C.c_void_p()
        self._ffi.check_comm(self._ffi.lib.rscm_b200_comm_symmetric_alloc(self._h, n * 8, C.byref(p)), self._h)
        return torch.as_tensor(_DeviceBlock(p.value, n, self), device="cuda")

    def allgather(self, local, out, stream: int = 0) -> None:
        """``out[r*n + i] = rank r's local[i]`` (CUDA tensors, float64)."""
        self._ffi.check_comm(self._ffi.lib.rscm_b200_allgather_f64(self._h, local.data_ptr(), local.numel(), out.data_ptr(), stream), self._h)

    def log_posterior_sharded(self, ens, params, scenarios, logpost, *, M: int, S: int, layout: int = 0, stream: int = 0) -> None:
        """Every rank passes the same GLOBAL parameter matrix; all ranks end with all S*M log-posteriors in ``logpost``."""
        sp = 0 if scenarios is None else scenarios.data_ptr()
        self._ffi.check_comm(self._ffi.lib.rscm_b200_logpost_sharded_device(ens._h, self._h, params.data_ptr(), M, layout, sp, S,
                                                                           logpost.data_ptr(), stream), self._h)

    def check(self) -> None:
        self._ffi.check_comm(self._ffi.lib.rscm_b200_comm_check(self._h), self._h)
