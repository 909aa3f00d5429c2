"""Mirror of ``rscm._lib.calibrate`` (python/rscm/_lib/calibrate.pyi) for the ensemble path.

``ModelRunner`` here is the GPU replacement for ``DefaultModelRunner``
(crates/rscm-calibrate/src/model_runner.rs:116-266): instead of a factory closure that
rebuilds a model per member it takes a ``ModelBuilder`` plus a column -> parameter-slot
binding, compiles the graph once, and ``run_batch`` evaluates every member in one fused
kernel launch.  ``EnsembleSampler`` keeps the reference's Goodman–Weare stretch-move loop
(sampler/ensemble.rs:412-546, sampler/moves.rs:55-125) on the host and calls the fused
log-posterior kernel for each half-ensemble.

Priors, target and likelihood values are only *described* here; their arithmetic runs on the
device (kernel.cuh: prior_ln_pdf, obs_accumulate).  The ``ln_pdf`` methods on the distribution
classes are host conveniences for building walkers and are not used by the hot path.
"""

from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from . import _ffi
from .core import Ensemble, ModelBuilder

__all__ = [
    "Uniform", "Normal", "LogNormal", "Bound", "ParameterSet", "Observation", "VariableTarget", "Target",
    "GaussianLikelihood", "ModelRunner", "WalkerInit", "Chain", "EnsembleSampler", "DeviceEnsembleSampler", "ProgressInfo",
]

_LN_2PI = math.log(2.0 * math.pi)


# ---- distributions: crates/rscm-calibrate/src/distribution.rs ----------------
class Uniform:
    def __init__(self, low: float, high: float) -> None:
        if not low < high:
            raise ValueError("Uniform requires low < high")
        self._low, self._high = float(low), float(high)

    low = property(lambda self: self._low)
    high = property(lambda self: self._high)

    def sample(self, rng=None) -> float:
        return float((rng or np.random.default_rng()).uniform(self._low, self._high))

    def sample_n(self, n: int, rng=None) -> np.ndarray:
        return (rng or np.random.default_rng()).uniform(self._low, self._high, size=n)

    def ln_pdf(self, x: float) -> float:  # :157-163
        return -math.inf if (x < self._low or x > self._high) else -math.log(self._high - self._low)

    def bounds(self):
        return (self._low, self._high)

    def _abi(self):
        return (_ffi.PRIOR_UNIFORM, self._low, self._high, 0.0, 0.0)


class Normal:
    def __init__(self, mean: float, std_dev: float) -> None:
        if not std_dev > 0:
            raise ValueError("Normal requires std_dev > 0")
        self._mean, self._std = float(mean), float(std_dev)

    mean = property(lambda self: self._mean)
    std_dev = property(lambda self: self._std)

    def sample(self, rng=None) -> float:
        return float((rng or np.random.default_rng()).normal(self._mean, self._std))

    def sample_n(self, n: int, rng=None) -> np.ndarray:
        return (rng or np.random.default_rng()).normal(self._mean, self._std, size=n)

    def ln_pdf(self, x: float) -> float:  # :256-259
        z = (x - self._mean) / self._std
        return -0.5 * z * z - math.log(self._std) - 0.5 * _LN_2PI

    def bounds(self):
        return None

    def _abi(self):
        return (_ffi.PRIOR_NORMAL, self._mean, self._std, 0.0, 0.0)


class LogNormal:
    def __init__(self, mu: float, sigma: float) -> None:
        if not sigma > 0:
            raise ValueError("LogNormal requires sigma > 0")
        self._mu, self._sigma = float(mu), float(sigma)

    mu = property(lambda self: self._mu)
    sigma = property(lambda self: self._sigma)

    def sample(self, rng=None) -> float:
        return float((rng or np.random.default_rng()).lognormal(self._mu, self._sigma))

    def sample_n(self, n: int, rng=None) -> np.ndarray:
        return (rng or np.random.default_rng()).lognormal(self._mu, self._sigma, size=n)

    def ln_pdf(self, x: float) -> float:  # :353-360
        if x <= 0.0:
            return -math.inf
        ln_x = math.log(x)
        z = (ln_x - self._mu) / self._sigma
        return -0.5 * z * z - ln_x - math.log(self._sigma) - 0.5 * _LN_2PI

    def bounds(self):
        return (0.0, math.inf)

    def _abi(self):
        return (_ffi.PRIOR_LOGNORMAL, self._mu, self._sigma, 0.0, 0.0)


class Bound:
    """Truncation wrapper; unnormalised inner pdf inside [low, high] (:490-497)."""

    def __init__(self, distribution, low: float, high: float) -> None:
        if not low < high:
            raise ValueError("Bound requires low < high")
        self._dist, self._low, self._high = distribution, float(low), float(high)

    def sample(self, rng=None) -> float:
        for _ in range(10000):
            x = self._dist.sample(rng)
            if self._low <= x <= self._high:
                return x
        raise RuntimeError("Bound.sample: rejection sampling failed")

    def sample_n(self, n: int, rng=None) -> np.ndarray:
        """Vectorised rejection sampling (same acceptance rule as `sample`)."""
        out = np.empty(n)
        todo = np.arange(n)
        for _ in range(10000):
            x = self._dist.sample_n(todo.size, rng)
            ok = (x >= self._low) & (x <= self._high)
            out[todo[ok]] = x[ok]
            todo = todo[~ok]
            if todo.size == 0:
                return out
        raise RuntimeError("Bound.sample: rejection sampling failed")

    def ln_pdf(self, x: float) -> float:
        return -math.inf if (x < self._low or x > self._high) else self._dist.ln_pdf(x)

    def bounds(self):
        return (self._low, self._high)

    def _abi(self):
        k, a, b, _, _ = self._dist._abi()
        kind = {_ffi.PRIOR_NORMAL: _ffi.PRIOR_BOUND_NORMAL, _ffi.PRIOR_LOGNORMAL: _ffi.PRIOR_BOUND_LOGNORMAL,
                _ffi.PRIOR_UNIFORM: _ffi.PRIOR_BOUND_UNIFORM}[k]
        return (kind, a, b, self._low, self._high)


class ParameterSet:
    """crates/rscm-calibrate/src/parameter_set.rs (insertion-ordered)."""

    def __init__(self) -> None:
        self._params: dict[str, object] = {}

    def add(self, name: str, distribution) -> "ParameterSet":
        self._params[name] = distribution
        return self

    def __len__(self) -> int:
        return len(self._params)

    @property
    def param_names(self) -> list[str]:
        return list(self._params)

    def sample_random(self, n: int, rng=None) -> np.ndarray:
        """[n, n_params]: n independent draws of every parameter (parameter_set.rs sample_random), drawn column by column."""
        rng = rng or np.random.default_rng()
        return np.column_stack([d.sample_n(n, rng) for d in self._params.values()]) if len(self) else np.empty((n, 0))

    def sample_lhs(self, n: int, rng=None) -> np.ndarray:
        rng = rng or np.random.default_rng()
        out = np.empty((n, len(self)))
        for j, d in enumerate(self._params.values()):
            b = d.bounds()
            if b is None or not all(map(math.isfinite, b)):
                out[:, j] = [d.sample(rng) for _ in range(n)]
                continue
            u = (rng.permutation(n) + rng.uniform(size=n)) / n
            out[:, j] = b[0] + u * (b[1] - b[0])
        return out

    def log_prior(self, params: Sequence[float]) -> float:  # :255-270
        if len(params) != len(self):
            raise ValueError("parameter vector length does not match parameter set size")
        return float(sum(d.ln_pdf(float(x)) for x, d in zip(params, self._params.values())))

    def bounds(self):
        lo, hi = [], []
        for d in self._params.values():
            b = d.bounds()
            lo.append(-math.inf if b is None else b[0])
            hi.append(math.inf if b is None else b[1])
        return lo, hi

    def _abi(self):
        return [d._abi() for d in self._params.values()]


# ---- targets: crates/rscm-calibrate/src/target.rs -------------------------------
class Observation:
    def __init__(self, time: float, value: float, uncertainty: float) -> None:
        if not uncertainty > 0:
            raise ValueError("uncertainty must be positive")
        self.time, self.value, self.uncertainty = float(time), float(value), float(uncertainty)


class VariableTarget:
    def __init__(self, name: str) -> None:
        self.name = name
        self.observations: list[Observation] = []
        self.reference_period = None

    def add(self, time, value, uncertainty) -> "VariableTarget":
        self.observations.append(Observation(time, value, uncertainty))
        return self

    def add_relative(self, time, value, relative_uncertainty) -> "VariableTarget":
        return self.add(time, value, abs(value) * relative_uncertainty)

    def with_reference_period(self, start, end) -> "VariableTarget":
        self.reference_period = (float(start), float(end))
        return self

    def time_range(self):
        if not self.observations:
            return None
        t = [o.time for o in self.observations]
        return (min(t), max(t))


class Target:
    def __init__(self) -> None:
        self._vars: dict[str, VariableTarget] = {}

    def add_variable(self, name: str) -> VariableTarget:
        return self._vars.setdefault(name, VariableTarget(name))

    def add_observation(self, variable, time, value, uncertainty) -> "Target":
        self.add_variable(variable).add(time, value, uncertainty)
        return self

    def add_observation_relative(self, variable, time, value, relative_uncertainty) -> "Target":
        self.add_variable(variable).add_relative(time, value, relative_uncertainty)
        return self

    def set_reference_period(self, variable, start, end) -> "Target":
        self.add_variable(variable).with_reference_period(start, end)
        return self

    def get_variable(self, name):
        return self._vars.get(name)

    def variable_names(self) -> list[str]:
        return list(self._vars)

    def total_observations(self) -> int:
        return sum(len(v.observations) for v in self._vars.values())

    def time_range(self):
        r = [v.time_range() for v in self._vars.values() if v.observations]
        return None if not r else (min(a for a, _ in r), max(b for _, b in r))

    def _flat(self):
        return [(v.name, o.time, o.value, o.uncertainty) for v in self._vars.values() for o in v.observations]


class GaussianLikelihood:
    """likelihood.rs:99-253 — evaluated on the device; this object only carries `normalize`."""

    def __init__(self, normalize: bool = False) -> None:
        self.normalize = bool(normalize)


# ---- runner ------------------------------------------------------------------------
class ModelRunner:
    """GPU ``ModelRunner``: ``builder`` describes the model, ``bindings`` maps each name in
    ``param_names`` to the slot(s) it feeds (see :meth:`Ensemble.bind_parameters`)."""

    def __init__(self, builder: ModelBuilder, bindings: dict, output_variables: Sequence[str], *, scenarios=None,
                 dtype: str = "f64", device: int = -1) -> None:
        self._ens: Ensemble = builder.build_ensemble(dtype=dtype, device=device)
        self._ens.bind_parameters(bindings)
        for v in output_variables:
            if v not in self._ens.variable_names:
                raise ValueError(f"Model output missing variable: {v}")
            if self._ens.regions(v) != 1:
                raise ValueError(f"Grid variables not yet supported: {v}")  # model_runner.rs:175-190
        self._ens.select_outputs(output_variables)
        self._outputs = list(output_variables)
        self._scenarios = scenarios

    @property
    def ensemble(self) -> Ensemble:
        return self._ens

    @property
    def param_names(self) -> list[str]:
        return list(self._ens.param_names)

    @property
    def output_variables(self) -> list[str]:
        return list(self._outputs)

    def run_batch_arrays(self, param_sets) -> dict[str, np.ndarray]:
        """{variable: [T, M]} for all members (NaN where the reference stores NaN)."""
        p = np.atleast_2d(np.asarray(param_sets, dtype=np.float64))
        if p.shape[1] != len(self.param_names):
            raise ValueError(f"Expected {len(self.param_names)} parameters, got {p.shape[1]}")
        out = self._ens.run(p, self._scenarios, layout=1)
        return self._ens.split_outputs(out)

    def run_batch(self, param_sets) -> list[dict[str, dict[float, float]]]:
        """``ModelRunner::run_batch`` — one ``{variable: {time: value}}`` per member, NaN
        entries skipped like ``extract_outputs`` (model_runner.rs:161-216)."""
        arrays = self.run_batch_arrays(param_sets)
        times = self._ens.selected_times()
        M = next(iter(arrays.values())).shape[1]
        res = []
        for m in range(M):
            res.append({v: {float(t): float(x) for t, x in zip(times, a[:, m]) if not math.isnan(x)} for v, a in arrays.items()})
        return res

    def run(self, params: Sequence[float]) -> dict[str, dict[float, float]]:
        if len(params) != len(self.param_names):
            raise ValueError(f"Expected {len(self.param_names)} parameters, got {len(params)}")
        return self.run_batch([list(params)])[0]


# ---- sampler -------------------------------------------------------------------------
class ProgressInfo:
    def __init__(self, iteration, total, acceptance_rate, mean_log_prob):
        self.iteration, self.total = iteration, total
        self.acceptance_rate, self.mean_log_prob = acceptance_rate, mean_log_prob


class WalkerInit:
    def __init__(self, kind, center=None, radius=None, positions=None):
        self.kind, self.center, self.radius, self.positions = kind, center, radius, positions

    @staticmethod
    def from_prior() -> "WalkerInit":
        return WalkerInit("prior")

    @staticmethod
    def ball(center, radius) -> "WalkerInit":
        return WalkerInit("ball", center=[float(c) for c in center], radius=float(radius))

    @staticmethod
    def explicit(positions) -> "WalkerInit":
        return WalkerInit("explicit", positions=np.asarray(positions, dtype=np.float64))

    def initialize(self, n_walkers: int, params: ParameterSet, rng) -> np.ndarray:  # sampler/init.rs:40
        if self.kind == "prior":
            return params.sample_random(n_walkers, rng)
        if self.kind == "ball":
            if len(self.center) != len(params):
                raise ValueError("Ball center length does not match parameter count")
            off = rng.uniform(size=(n_walkers, len(params))) - 0.5
            return np.asarray(self.center)[None, :] + off * self.radius
        if self.positions.shape != (n_walkers, len(params)):
            raise ValueError("Explicit positions have the wrong shape")
        return self.positions.copy()


class Chain:
    """sampler/chain.rs — samples [n_stored, n_walkers, n_params] with thinning."""

    def __init__(self, param_names, thin=1):
        self._names, self._thin = list(param_names), max(1, int(thin))
        self._samples: list[np.ndarray] = []
        self._log_probs: list[np.ndarray] = []
        self._total = 0

    def push(self, positions, log_probs) -> None:
        if self._total % self._thin == 0:
            self._samples.append(positions.copy())
            self._log_probs.append(log_probs.copy())
        self._total += 1

    param_names = property(lambda self: list(self._names))
    thin = property(lambda self: self._thin)
    total_iterations = property(lambda self: self._total)

    def __len__(self) -> int:
        return len(self._samples)

    def flat_samples(self, discard: int = 0) -> np.ndarray:
        s = np.asarray(self._samples[discard:])
        return s.reshape(-1, len(self._names))

    def flat_log_probs(self, discard: int = 0) -> np.ndarray:
        return np.asarray(self._log_probs[discard:]).reshape(-1)

    def to_param_dict(self, discard: int = 0) -> dict:
        f = self.flat_samples(discard)
        return {n: f[:, j] for j, n in enumerate(self._names)}


class EnsembleSampler:
    """Goodman & Weare stretch move; two fused log-posterior launches per iteration."""

    def __init__(self, params: ParameterSet, runner: ModelRunner, likelihood: GaussianLikelihood, target: Target, *,
                 stretch: float = 2.0, seed: int | None = None) -> None:
        if list(params.param_names) != list(runner.param_names):
            raise ValueError("ParameterSet and ModelRunner parameter names/order differ")
        if not stretch > 1.0:
            raise ValueError("stretch parameter must be > 1")
        self.params, self.runner, self.likelihood, self.target = params, runner, likelihood, target
        self.a = float(stretch)
        self._rng = np.random.default_rng(seed)
        ens = runner.ensemble
        ens.set_target(target._flat(), normalize=likelihood.normalize)
        ens.set_priors(params._abi())
        self._default_n_walkers = max(2 * len(params), 32)  # sampler/ensemble.rs:115-117
        self.acceptance_rate = float("nan")

    def default_n_walkers(self) -> int:
        return self._default_n_walkers

    def log_posterior_batch(self, param_sets: np.ndarray) -> np.ndarray:  # sampler/ensemble.rs:143-178
        return self.runner.ensemble.log_posterior(np.asarray(param_sets, dtype=np.float64), self.runner._scenarios, layout=1)

    def run(self, n_iterations: int, init: WalkerInit, thin: int = 1, n_walkers: int | None = None, progress=None) -> Chain:
        n_walkers = n_walkers or self._default_n_walkers
        if n_walkers < 2:
            raise ValueError("Must have at least 2 walkers")
        if n_walkers % 2:
            raise ValueError("Number of walkers must be even")
        rng = self._rng
        pos = init.initialize(n_walkers, self.params, rng)
        logp = self.log_posterior_batch(pos)
        chain = Chain(self.params.param_names, thin)
        n_params = len(self.params)
        half = n_walkers // 2
        n_acc = n_prop = 0
        for it in range(n_iterations):
            for active, comp in ((slice(0, half), slice(half, n_walkers)), (slice(half, n_walkers), slice(0, half))):
                cpos = pos[comp].copy()
                na = pos[active].shape[0]
                z = ((self.a - 1.0) * rng.uniform(size=na) + 1.0) ** 2 / self.a  # moves.rs:55-59
                ci = rng.integers(0, cpos.shape[0], size=na)
                prop = cpos[ci] + z[:, None] * (pos[active] - cpos[ci])  # moves.rs:110-125
                lp_new = self.log_posterior_batch(prop)
                with np.errstate(over="ignore", invalid="ignore"):
                    log_ratio = (n_params - 1.0) * np.log(z) + (lp_new - logp[active])
                    acc_p = np.where(np.isfinite(lp_new), np.minimum(np.exp(log_ratio), 1.0), 0.0)  # moves.rs:76-92
                acc = rng.uniform(size=na) < acc_p
                idx = np.arange(n_walkers)[active][acc]
                pos[idx] = prop[acc]
                logp[idx] = lp_new[acc]
                n_acc += int(acc.sum())
                n_prop += na
            chain.push(pos, logp)
            if progress is not None:
                progress(ProgressInfo(it, n_iterations, n_acc / max(1, n_prop), float(np.mean(logp))))
        self.acceptance_rate = n_acc / max(1, n_prop)
        return chain


class DeviceEnsembleSampler(EnsembleSampler):
    """The same Goodman & Weare loop with nothing but the chain leaving the GPU.

    Walker positions ``[n_params][n_walkers]``, proposals, stretch factors and log-posteriors stay in HBM and the loop itself
    runs behind the C ABI (``rscm_b200_sampler_iterate``, include/rscm_b200.h; sampler/ensemble.rs:489-546,
    sampler/moves.rs:55-125): per half-update ``stretch_propose`` -> fused log-posterior kernel -> ``stretch_accept`` on one
    stream, each iteration replayed from one CUDA graph.  Random draws are Philox4x32-10 functions of (seed, walker, step,
    purpose), so a run is reproducible from ``seed`` and, with several ranks, every rank keeps an identical replica of the
    walker state: each rank evaluates its member block of the active half in place and the 8-byte log-posteriors are
    all-gathered — stored straight into every peer's buffer by the log-posterior kernel when the GPUs map each other's
    memory (``collective == "fused peer stores"``), by NCCL otherwise — no other exchange.
    """

    MAX_DEVICE_CHAIN_BYTES = 4 << 30

    def _stream(self):
        import torch
        return int(torch.cuda.current_stream().cuda_stream)

    def _comm(self, distributed):
        """``distributed``: False / None (one GPU), True (torch.distributed's default group), a process group, or a
        :class:`rscm_b200.dist.Comm`."""
        from .dist import Comm
        if isinstance(distributed, Comm):
            return distributed
        key = "single" if not distributed else ("world" if distributed is True else id(distributed))
        cache = self.__dict__.setdefault("_comms", {})
        if key not in cache:
            import torch
            dev = torch.cuda.current_device()
            cache[key] = Comm.single(dev) if not distributed else Comm.from_torch(None if distributed is True else distributed, dev)
        return cache[key]

    def run(self, n_iterations: int, init: WalkerInit, thin: int = 1, n_walkers: int | None = None, progress=None, *,
            seed: int | None = None, distributed: bool | object = False, use_graph: bool = True) -> Chain:
        import ctypes as C

        import torch

        W = n_walkers or self._default_n_walkers
        if W < 2:
            raise ValueError("Must have at least 2 walkers")
        if W % 2:
            raise ValueError("Number of walkers must be even")
        scen = self.runner._scenarios
        if scen is not None and scen.shape[0] != 1:
            raise ValueError("the sampler evaluates one scenario per walker; the runner holds %d" % scen.shape[0])
        comm = self._comm(distributed)
        dev = torch.device("cuda", torch.cuda.current_device())
        d_scen = None if scen is None else torch.from_numpy(np.ascontiguousarray(scen)).to(dev)
        S = 0 if d_scen is None else 1
        P, half, thin = len(self.params), W // 2, max(1, int(thin))
        if seed is None:
            seed = int(self._rng.integers(0, 2 ** 63))
        pos0 = torch.from_numpy(np.ascontiguousarray(init.initialize(W, self.params, self._rng).T)).to(dev)  # [P][W]
        if comm.world > 1:   # one seed, one initial ensemble on every rank
            box = torch.empty(P * W + 1, dtype=torch.float64, device=dev)
            box[0] = float(seed % (1 << 52))
            box[1:] = pos0.reshape(-1)
            gathered = torch.empty(comm.world * box.numel(), dtype=torch.float64, device=dev)
            comm.allgather(box, gathered, self._stream())
            torch.cuda.synchronize()
            seed = int(gathered[0].item())
            pos0 = gathered[1:box.numel()].reshape(P, W).clone()
        ens, lib = self.runner.ensemble, _ffi.lib
        d_pos = pos0.contiguous()
        # initial log-posterior of all W walkers (sharded like every later evaluation)
        key = (id(comm), W)
        if getattr(self, "_sym_key", None) != key:   # symmetric memory lives as long as the communicator: allocate once per shape
            self._sym = (comm.symmetric_empty(W), comm.symmetric_empty(half), comm.symmetric_empty(half))
            self._sym_key = key
        d_logp, d_lpn0, d_lpn1 = self._sym
        comm.log_posterior_sharded(ens, d_pos, d_scen, d_logp, M=W, S=S, layout=0, stream=self._stream())
        d_prop = torch.empty((P, half), dtype=torch.float64, device=dev)
        d_z = torch.empty(half, dtype=torch.float64, device=dev)
        d_nacc = torch.zeros(1, dtype=torch.int64, device=dev)
        n_keep = (n_iterations + thin - 1) // thin
        on_device = n_keep * W * (P + 1) * 8 <= self.MAX_DEVICE_CHAIN_BYTES
        if on_device:
            d_samples = torch.empty((n_keep, P, W), dtype=torch.float64, device=dev)
            d_logps = torch.empty((n_keep, W), dtype=torch.float64, device=dev)
        chain = Chain(self.params.param_names, thin)
        st = _ffi.SamplerState()
        st.positions, st.ld, st.n_cols, st.n_walkers = d_pos.data_ptr(), W, P, W
        st.logpost, st.proposals, st.z = d_logp.data_ptr(), d_prop.data_ptr(), d_z.data_ptr()
        st.logpost_new[0], st.logpost_new[1] = d_lpn0.data_ptr(), d_lpn1.data_ptr()
        st.n_accepted, st.a, st.seed = d_nacc.data_ptr(), self.a, seed
        sp = 0 if d_scen is None else d_scen.data_ptr()

        if on_device:
            st.thin, st.chain_positions, st.chain_logpost, st.chain_capacity = thin, d_samples.data_ptr(), d_logps.data_ptr(), n_keep

        def advance(first: int, n: int) -> None:
            st.first_iteration = first
            _ffi.check_comm(lib.rscm_b200_sampler_iterate(ens._h, comm._h, C.byref(st), sp, S, n, 1 if use_graph else 0, self._stream()), comm._h)

        if on_device and progress is None:
            advance(0, n_iterations)      # the whole run is enqueued by one call; kept samples are recorded on the device
        else:
            for it in range(n_iterations):
                advance(it, 1)
                if not on_device and it % thin == 0:
                    chain._samples.append(d_pos.T.cpu().numpy())
                    chain._log_probs.append(d_logp.cpu().numpy())
                if progress is not None:  # a host read-back per iteration: only when asked for
                    progress(ProgressInfo(it, n_iterations, int(d_nacc.item()) / ((it + 1) * W), float(d_logp.mean().item())))
        torch.cuda.synchronize()
        comm.check()
        if on_device and n_keep:
            chain._samples = list(d_samples.permute(0, 2, 1).contiguous().cpu().numpy())
            chain._log_probs = list(d_logps.cpu().numpy())
        chain._total = n_iterations
        self.acceptance_rate = int(d_nacc.item()) / max(1, n_iterations * W)
        self.final_positions = d_pos.T.cpu().numpy()
        self.final_log_probs = d_logp.cpu().numpy()
        self.collective = None if comm.world == 1 else ("fused peer stores (NVLink)" if comm.peer_access else "ncclAllGather")
        return chain
