#!/usr/bin/env python
"""Short config-4 run (MAGICC boxes + ClimateUDEB) for ncu: one wave of the lane-group kernel (148 SMs x 4 CTAs x 32 members), 60 years."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscm_b200 import synthetic as syn

M = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 4 * 32
axis = syn.time_axis(1850, 1910)
b = syn.config4_builder(axis)
params = syn.uniform_params(syn.CONFIG4_RANGES, M, 7)
ens = b.build_ensemble().bind_parameters(syn.CONFIG4_BINDINGS)
ens.select_outputs(syn.CONFIG4_OUTPUTS)
sc = torch.from_numpy(ens.pack_scenarios([syn.config4_scenario(axis.values())])).cuda()
p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
out = torch.empty((ens.output_rows, M), dtype=torch.float64, device="cuda")
for _ in range(3):
    ens.run_device(p, sc, out, layout=0)
torch.cuda.synchronize()
print("kernel ms", ens.kernel_ms(), "member-years/s", M * 60 / (ens.kernel_ms() * 1e-3))
