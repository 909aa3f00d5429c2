#!/usr/bin/env python
"""Small instances of every kernel family for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py > gpurun_out/memcheck.log 2>&1
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/racecheck.log 2>&1

coupled graph (AOT, TMA staging), log-posterior with the ticket reduction, lane-quad ClimateUDEB (shuffles, shared-memory
off-diagonal columns), the emissions-driven chain (OceanCarbon block prefix sums in shared memory, global history),
device quantiles and the captured sampler iteration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rscm_b200 import _ffi, synthetic as syn  # noqa: E402


def main():
    axis = syn.time_axis(1850, 1890)
    # 1. coupled AOT graph: write + log-posterior + summary
    b = syn.coupled_builder(axis=axis)
    ens = b.build_ensemble().bind_parameters(syn.COUPLED_BINDINGS)
    params = syn.uniform_params(syn.COUPLED_RANGES, 300, 3)
    em = syn.emission_scenarios(axis.values(), 2)
    sc = ens.pack_scenarios([{"Emissions|CO2|Anthropogenic": em[s]} for s in range(2)])
    out = ens.run(params, sc)
    obs = [("Surface Temperature", float(y), 0.3, 0.2) for y in range(1860, 1891, 10)]
    ens.set_target(obs).set_priors([(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.COUPLED_RANGES.values()])
    lp, summ = ens.log_posterior(params, sc, with_summary=True)
    q = ens.run_quantiles(params, sc, [0.1, 0.5, 0.9])
    print("coupled ok", out.shape, summ["n_finite"], list(q)[:1])
    # 2. config 4: lane-quad ClimateUDEB
    b4 = syn.config4_builder(axis)
    e4 = b4.build_ensemble().bind_parameters(syn.CONFIG4_BINDINGS)
    p4 = syn.uniform_params(syn.CONFIG4_RANGES, 41, 5)
    o4 = e4.run(p4, e4.pack_scenarios([syn.config4_scenario(axis.values())]))
    print("config4 ok", o4.shape, float(np.nanmax(o4)))
    # 3. emissions-driven chain (OceanCarbon in lane quads, halocarbons)
    from tests.test_ocean_carbon import FULL_BINDS, full_magicc_builder, full_magicc_scenario
    bf = full_magicc_builder(end=1875)
    ef = bf.build_ensemble().bind_parameters(FULL_BINDS)
    pf = syn.uniform_params({"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}, 19, 41)
    ef.select_outputs(["Surface Temperature", "Atmospheric Concentration|CO2"])
    of = ef.run(pf, ef.pack_scenarios([full_magicc_scenario(end=1875)]))
    print("full chain ok", of.shape)
    # 4. sampler iteration (eager + captured graph)
    from rscm_b200.calibrate import DeviceEnsembleSampler, GaussianLikelihood, ModelRunner, ParameterSet, Target, Uniform, WalkerInit
    b2, binds2, _, scen2 = syn.config2(M=4)
    runner = ModelRunner(b2, {k: binds2[k] for k in ("lambda0", "efficacy")}, ["Surface Temperature"])
    runner._scenarios = runner.ensemble.pack_scenarios(scen2)
    target = Target()
    for y in range(1900, 2001, 20):
        target.add_observation("Surface Temperature", float(y), 0.5, 0.2)
    ps = ParameterSet().add("lambda0", Uniform(0.6, 1.8)).add("efficacy", Uniform(0.8, 2.0))
    s = DeviceEnsembleSampler(ps, runner, GaussianLikelihood(), target, seed=1)
    chain = s.run(4, WalkerInit.from_prior(), n_walkers=64)
    torch.cuda.synchronize()
    print("sampler ok", len(chain), s.acceptance_rate)


if __name__ == "__main__":
    main()
