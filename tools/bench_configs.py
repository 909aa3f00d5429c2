#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs other than the headline one (bench.py covers configs[2]).

  config 2  two-layer ensemble, 1 048 576 members x 1 forcing scenario, fp64 (+ fp32)
  config 4  MAGICC box components on the four-box grid, 100 000 members, fp64 vs fp32
  config 5  log-posterior of 1 048 576 two-layer members against 171 synthetic observations (fused K4 kernel)

Each line: device-resident throughput (CUDA events on the launch stream, 3 warm-ups, best of 5), parity of a strided
subsample against the CPU oracle (fp64 1e-9 / fp32 reported), and the CPU oracle's own throughput on a bounded sample.
Run on a GPU box:  python tools/bench_configs.py > gpurun_out/configs.jsonl
"""

from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402  (checker + CPU baseline leg only)
from rscm_b200 import _ffi, synthetic as syn  # noqa: E402
from tests.helpers import oracle_bindings, oracle_from_builder, rel_err  # noqa: E402


def time_launches(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run_config(name, builder, binds, params, scen, outputs, dtype, sub_stride, years=350):
    ens = builder.build_ensemble(dtype=dtype).bind_parameters(binds)
    ens.select_outputs(outputs)
    sc_host = ens.pack_scenarios(scen)
    M, S = params.shape[0], len(scen)
    d_p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    d_s = torch.from_numpy(sc_host).cuda()
    d_o = torch.empty((ens.output_rows, S * M), dtype=torch.float64, device="cuda")
    ms = time_launches(lambda: ens.run_device(d_p, d_s, d_o, layout=0))
    torch.cuda.synchronize()
    idx = np.arange(0, M, sub_stride)
    sub = d_o[:, torch.from_numpy(np.concatenate([idx + s * M for s in range(S)])).cuda()].cpu().numpy()
    m = oracle_from_builder(builder)
    t0 = time.perf_counter()
    ref = m.run_batch(oracle_bindings(builder, binds), params[idx], ens.exogenous_names, sc_host, outputs)
    cpu_s = time.perf_counter() - t0
    got, want = ens.split_outputs(sub), m.split(ref, outputs)
    errs = {n: rel_err(got[n], want[n]) for n in outputs}
    # per-run view of the same comparison: max |gpu - cpu| over the run's series / max |cpu| over it
    per_run = np.zeros(sub.shape[-1])
    for n in outputs:
        a, e = got[n].reshape(-1, sub.shape[-1]), want[n].reshape(-1, sub.shape[-1])
        with np.errstate(invalid="ignore"):
            per_run = np.maximum(per_run, np.nanmax(np.abs(a - e), axis=0) / np.maximum(np.nanmax(np.abs(e), axis=0), 1e-300))
    line = {"config": name, "dtype": dtype, "members": M, "scenarios": S, "ms": ms, "member_years_per_s": M * S * years / (ms * 1e-3),
            "jit": ens.program_is_jit(), "parity_subsample": int(idx.size * S), "max_rel_err": max(errs.values()), "rel_err": errs,
            "per_run_rel_err": {"median": float(np.median(per_run)), "p99": float(np.percentile(per_run, 99)), "max": float(per_run.max()),
                                "frac_within_1e-4": float(np.mean(per_run <= 1e-4)), "frac_within_1e-9": float(np.mean(per_run <= 1e-9))},
            "cpu_oracle_member_years_per_s": idx.size * S * years / cpu_s, "cpu_threads": orc.max_threads()}
    print(json.dumps(line), flush=True)
    return ens


def main():
    only = set(sys.argv[1:]) or {"2", "4", "5"}  # e.g. `bench_configs.py 4`
    if "2" in only:
        b, binds, params, scen = syn.config2(M=1 << 20)
        for dt in ("f64", "f32"):
            run_config("2: two-layer 1M x 1", b, binds, params, scen, ["Surface Temperature", "Deep Ocean Temperature"], dt, 257)
    if "4" in only:
        b, binds, params, scen = syn.config4(M=100_000)
        for dt in ("f64", "f32"):
            run_config("4: MAGICC boxes + ClimateUDEB (four-box) 100k", b, binds, params, scen, syn.CONFIG4_OUTPUTS, dt, 499)
    if "5" not in only:
        return
    # config 5: fused log-posterior
    b, binds, params, scen = syn.config2(M=1 << 20)
    ens = b.build_ensemble().bind_parameters(binds)
    sc_host = ens.pack_scenarios(scen)
    truth = dict(syn.TWO_LAYER_DEFAULTS, lambda0=1.1, efficacy=1.3, a=0.05)
    ens.select_outputs(["Surface Temperature"])
    t_true = ens.run(np.array([[truth[k] for k in syn.TWO_LAYER_RANGES]]), sc_host)[:, 0]
    obs = syn.config5_observations(t_true, syn.time_axis().values())
    priors = [(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.TWO_LAYER_RANGES.values()]
    ens.set_target(obs).set_priors(priors)
    M = params.shape[0]
    d_p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    d_s = torch.from_numpy(sc_host).cuda()
    d_lp = torch.empty(M, dtype=torch.float64, device="cuda")
    d_sum = torch.zeros(5, dtype=torch.float64, device="cuda")
    ms = time_launches(lambda: ens.log_posterior_device(d_p, d_s, d_lp, d_sum, layout=0))
    idx = np.arange(0, M, 257)
    m = oracle_from_builder(b)
    t0 = time.perf_counter()
    ref = m.log_posterior_batch(oracle_bindings(b, binds), params[idx], ens.exogenous_names, sc_host, priors, obs)
    cpu_s = time.perf_counter() - t0
    got = d_lp.cpu().numpy()[idx]
    fin = np.isfinite(ref)
    err = float(np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin])))
    print(json.dumps({"config": "5: log-posterior, 1M two-layer members, 171 observations", "dtype": "f64", "members": M, "ms": ms,
                      "member_years_per_s": M * 350 / (ms * 1e-3), "bytes_out_per_member": 8, "max_rel_err": err,
                      "inf_match": bool(np.array_equal(np.isinf(got), np.isinf(ref))),
                      "cpu_oracle_member_years_per_s": idx.size * 350 / cpu_s, "cpu_threads": orc.max_threads()}), flush=True)


def sampler_loop(W=1 << 20, iters=20):
    """config 5 as a loop: EnsembleSampler iterations over W walkers (two half-ensemble log-posterior evaluations each),
    host loop (numpy proposals, host<->device copies every half-update) vs DeviceEnsembleSampler (state resident in HBM)."""
    from rscm_b200.calibrate import (DeviceEnsembleSampler, EnsembleSampler, GaussianLikelihood, ModelRunner, ParameterSet, Target, Uniform,
                                     WalkerInit)
    b, binds, _, scen = syn.config2(M=4)
    runner = ModelRunner(b, binds, ["Surface Temperature"], scenarios=None)
    runner._scenarios = runner.ensemble.pack_scenarios(scen)
    truth = dict(syn.TWO_LAYER_DEFAULTS, lambda0=1.1, efficacy=1.3, a=0.05)
    t_true = runner.run_batch_arrays(np.array([[truth[k] for k in syn.TWO_LAYER_RANGES]]))["Surface Temperature"][:, 0]
    target = Target()
    for name, year, value, sigma in syn.config5_observations(t_true, syn.time_axis().values()):
        target.add_observation(name, year, value, sigma)
    ps = ParameterSet()
    for k, (lo, hi) in syn.TWO_LAYER_RANGES.items():
        ps.add(k, Uniform(lo, hi))
    res = {}
    for name, cls, n in (("device", DeviceEnsembleSampler, iters), ("host", EnsembleSampler, max(2, iters // 5))):
        s = cls(ps, runner, GaussianLikelihood(), target, seed=1)
        s.run(2, WalkerInit.from_prior(), n_walkers=W, thin=1000)   # warm-up (allocations, first launches)
        wall = []
        for k in (n, 3 * n):   # two run lengths: the difference removes walker initialisation and the final read-back
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.run(k, WalkerInit.from_prior(), n_walkers=W, thin=100000)
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
        dt = (wall[1] - wall[0]) / (2 * n)
        res[name] = {"s_per_iteration": dt, "member_years_per_s": W * 350 / dt, "acceptance_rate": s.acceptance_rate,
                     "setup_and_readback_s": wall[0] - n * dt}
    print(json.dumps({"config": "5 (loop): ensemble sampler iterations, %d walkers, 6 parameters, 171 observations" % W, **res,
                      "note": "per iteration: 2 half-updates = W log-posterior evaluations; steady state from two run lengths"}),
          flush=True)


def sampler_loop_distributed(W=1 << 20, iters=30):
    """config 5 over N GPUs (launch with torchrun): walker state replicated, log-posterior evaluation sharded by member,
    NCCL all-gather of the per-walker log-posteriors after every half-update (rscm_b200/dist.py)."""
    import torch.distributed as dist
    from rscm_b200.calibrate import DeviceEnsembleSampler, GaussianLikelihood, ModelRunner, ParameterSet, Target, Uniform, WalkerInit
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b, binds, _, scen = syn.config2(M=4)
    runner = ModelRunner(b, binds, ["Surface Temperature"], scenarios=None, device=local)
    runner._scenarios = runner.ensemble.pack_scenarios(scen)
    truth = dict(syn.TWO_LAYER_DEFAULTS, lambda0=1.1, efficacy=1.3, a=0.05)
    t_true = runner.run_batch_arrays(np.array([[truth[k] for k in syn.TWO_LAYER_RANGES]]))["Surface Temperature"][:, 0]
    target = Target()
    for name, year, value, sigma in syn.config5_observations(t_true, syn.time_axis().values()):
        target.add_observation(name, year, value, sigma)
    ps = ParameterSet()
    for k, (lo, hi) in syn.TWO_LAYER_RANGES.items():
        ps.add(k, Uniform(lo, hi))
    s = DeviceEnsembleSampler(ps, runner, GaussianLikelihood(), target, seed=1)
    s.run(3, WalkerInit.from_prior(), n_walkers=W, thin=1000, seed=5, distributed=True)   # warm-up (NCCL, allocations)
    wall = []
    for k in (iters, 3 * iters):   # two run lengths: the difference removes walker initialisation and the final read-back
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        s.run(k, WalkerInit.from_prior(), n_walkers=W, thin=100000, seed=6, distributed=True)
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
    dt = torch.tensor([(wall[1] - wall[0]) / (2 * iters)], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "5 (loop, %d GPUs): ensemble sampler iterations, %d walkers, 6 parameters, 171 observations" % (world, W),
                          "n_gpus": world, "s_per_iteration": float(dt.item()), "member_years_per_s": W * 350 / float(dt.item()),
                          "acceptance_rate": s.acceptance_rate,
                          "note": "replicated walker state, sharded log-posterior, all-gather of log-posteriors per half-update; "
                                  "steady state from two run lengths; max over ranks"}), flush=True)
    dist.destroy_process_group()


def full_magicc(M=37_888):
    """Not a BASELINE config: the emissions-driven MAGICC chain of the reference's regression suite (11 components incl.
    HalocarbonChemistry, 124 variables, run-time compiled), 1850-2100, as a throughput data point."""
    b, binds, params, scen = syn.full_chain(M=M)
    run_config("full MAGICC chain (11 components, 124 variables) %d members" % M, b, binds, params, scen, syn.FULL_CHAIN_OUTPUTS, "f64", 997, years=250)


def summaries(M=262_144):
    """F3: across-member quantiles of the headline workload's output block, where it lives."""
    b, binds, params, scen = syn.config3(M=M, S=8)
    ens = b.build_ensemble().bind_parameters(binds)
    sc = torch.from_numpy(ens.pack_scenarios(scen)).cuda()
    d_p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
    q = [0.05, 0.17, 0.5, 0.83, 0.95]
    for label, outs in (("Surface Temperature", ["Surface Temperature"]), ("all 7 series", syn.COUPLED_OUTPUTS)):
        ens.select_outputs(outs)
        d_o = torch.empty((ens.output_rows, 8 * M), dtype=torch.float64, device="cuda")
        ens.run_device(d_p, sc, d_o, layout=0)
        res = torch.empty((5, ens.output_rows, 8), dtype=torch.float64, device="cuda")
        ms = time_launches(lambda: ens.member_quantiles_device(d_o, q, res, M=M, S=8), reps=3, warm=1)
        nbytes = ens.output_rows * 8 * M * 8
        # parity of a strided row subsample against numpy
        rows = np.arange(0, ens.output_rows, max(1, ens.output_rows // 12))
        host = d_o[torch.from_numpy(rows).cuda()].cpu().numpy().reshape(rows.size, 8, M)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.nanquantile(host, q, axis=2)
        ok = bool(np.array_equal(res.cpu().numpy()[:, rows, :], want, equal_nan=True))
        print(json.dumps({"config": "3 (summary): 5 quantiles across %d members x 8 scenarios, %s" % (M, label), "rows": ens.output_rows,
                          "block_GB": nbytes / 1e9, "ms": ms, "block_GB_per_s": nbytes / 1e9 / (ms * 1e-3),
                          "d2h_bytes_instead_of_block": int(res.numel() * 8), "bit_identical_to_numpy_on_subsample": ok}), flush=True)
        del d_o, res


if __name__ == "__main__":
    if "summary" in sys.argv[1:]:
        summaries()
        sys.exit(0)
    if "magicc" in sys.argv[1:]:
        full_magicc()
        sys.exit(0)
    if "5dist" in sys.argv[1:]:
        sampler_loop_distributed()
        sys.exit(0)
    main()
    if not sys.argv[1:] or "5" in sys.argv[1:]:
        sampler_loop()
