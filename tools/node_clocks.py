#!/usr/bin/env python
"""Cycles per model year that one thread (thread 0 of CTA 0) spends in each node of the component graph, from the emitter's
RSCM_B200_NODE_CLOCKS instrumentation (graph.cpp: clock64 around every node; lane nodes include their barrier waits).  One
full wave of members so that the SM is loaded as in the benchmarks.  Usage: python tools/node_clocks.py [magicc|config4]"""
import os
import sys
import tempfile

os.environ["RSCM_B200_NODE_CLOCKS"] = "1"
os.environ.setdefault("RSCM_B200_CACHE", tempfile.mkdtemp(prefix="rscm_node_clocks_"))
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscm_b200 import synthetic as syn

which = sys.argv[1] if len(sys.argv) > 1 else "magicc"
M = 148 * 3 * 32
if which == "magicc":
    from tests.test_ocean_carbon import FULL_BINDS, full_magicc_builder, full_magicc_scenario
    b = full_magicc_builder(end=2100, halocarbons=True)
    ens = b.build_ensemble().bind_parameters(FULL_BINDS)
    from tests.test_halocarbon import ramp_scenario
    scen = full_magicc_scenario(end=2100)
    scen.pop("EESC")
    scen.update(ramp_scenario(251))
    params = syn.uniform_params({"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}, M, 43)
    ens.select_outputs(["Surface Temperature", "Atmospheric Concentration|CO2", "Atmospheric Concentration|CH4", "Effective Radiative Forcing", "EESC"])
else:
    axis = syn.time_axis(1850, 2100)
    b = syn.config4_builder(axis)
    ens = b.build_ensemble().bind_parameters(syn.CONFIG4_BINDINGS)
    params = syn.uniform_params(syn.CONFIG4_RANGES, M, 7)
    scen = syn.config4_scenario(axis.values())
    ens.select_outputs(syn.CONFIG4_OUTPUTS)
print("execution order", ens.execution_order(), flush=True)
sc = torch.from_numpy(ens.pack_scenarios([scen])).cuda()
p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
out = torch.empty((ens.output_rows, M), dtype=torch.float64, device="cuda")
for _ in range(2):
    ens.run_device(p, sc, out, layout=0)
torch.cuda.synchronize()
print("kernel ms", ens.kernel_ms(), flush=True)
