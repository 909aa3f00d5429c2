#!/usr/bin/env python
"""Short run of the full emissions-driven MAGICC chain (11 components, 124 variables) for ncu: 37 888 members, 1850-1950."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscm_b200 import synthetic as syn
from tests.test_halocarbon import ramp_scenario
from tests.test_ocean_carbon import FULL_BINDS, full_magicc_builder, full_magicc_scenario

M = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
b = full_magicc_builder(end=1950, halocarbons=True)
ens = b.build_ensemble().bind_parameters(FULL_BINDS)
s = full_magicc_scenario(end=1950)
s.pop("EESC")
s.update(ramp_scenario(101))
ens.select_outputs(["Surface Temperature", "Atmospheric Concentration|CO2", "EESC"])
sc = torch.from_numpy(ens.pack_scenarios([s])).cuda()
params = syn.uniform_params({"ecs": (2.0, 4.5), "beta": (0.4, 0.9), "tau": (6.5, 9.5), "tau_oh": (8.5, 10.5)}, M, 43)
p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
out = torch.empty((ens.output_rows, M), dtype=torch.float64, device="cuda")
for _ in range(3):
    ens.run_device(p, sc, out, layout=0)
torch.cuda.synchronize()
print("kernel ms", ens.kernel_ms(), "member-years/s", M * 100 / (ens.kernel_ms() * 1e-3))
