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
    b, binds, params, scens = syn.full_chain(M=M)
    scen = scens[0]
    ens = b.build_ensemble().bind_parameters(binds)
    ens.select_outputs(syn.FULL_CHAIN_OUTPUTS)
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
