#!/usr/bin/env python
"""Short run of the full emissions-driven MAGICC chain (11 components, 124 variables) for ncu: one full wave of the
lane-group kernel (148 SMs x 4 CTAs x 32 members), 1850-1950."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscm_b200 import synthetic as syn

M = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 4 * 32
b, binds, params, scen = syn.full_chain(M=M, end=1950)
ens = b.build_ensemble().bind_parameters(binds)
ens.select_outputs(["Surface Temperature", "Atmospheric Concentration|CO2", "EESC"])
sc = torch.from_numpy(ens.pack_scenarios(scen)).cuda()
p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
out = torch.empty((ens.output_rows, M), dtype=torch.float64, device="cuda")
for _ in range(3):
    ens.run_device(p, sc, out, layout=0)
torch.cuda.synchronize()
print("kernel ms", ens.kernel_ms(), "member-years/s", M * 100 / (ens.kernel_ms() * 1e-3))
