#!/usr/bin/env python
"""Phase timing of one DeviceEnsembleSampler half-update (CUDA events): propose / log-posterior / accept."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rscm_b200 import _ffi, synthetic as syn
from rscm_b200.calibrate import ModelRunner

W = 1 << 20
half = W // 2
b, binds, params, scen = syn.config2(M=W)
runner = ModelRunner(b, binds, ["Surface Temperature"], scenarios=None)
ens = runner.ensemble
sc = torch.from_numpy(ens.pack_scenarios(scen)).cuda()
t_true = ens.run(params[:1], sc.cpu().numpy())[:, 0]
ens.set_target(syn.config5_observations(t_true, syn.time_axis().values()))
ens.set_priors([(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.TWO_LAYER_RANGES.values()])
P = 6
d_pos = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
d_logp = torch.empty(W, dtype=torch.float64, device="cuda")
d_prop = torch.empty((P, half), dtype=torch.float64, device="cuda")
d_z = torch.empty(half, dtype=torch.float64, device="cuda")
d_lpn = torch.empty(half, dtype=torch.float64, device="cuda")
d_nacc = torch.zeros(1, dtype=torch.int64, device="cuda")
ens.log_posterior_device(d_pos, sc, d_logp, layout=0, M=W, S=1)
lib = _ffi.lib
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = np.zeros(3)
for it in range(12):
    ev[0].record()
    _ffi.check(lib.rscm_b200_stretch_propose(d_pos.data_ptr(), W, P, 0, half, half, half, 2.0, 7, it, d_prop.data_ptr(), half, d_z.data_ptr(), None))
    ev[1].record()
    ens.log_posterior_device(d_prop, sc, d_lpn, layout=0, M=half, S=1)
    ev[2].record()
    _ffi.check(lib.rscm_b200_stretch_accept(d_pos.data_ptr(), W, P, 0, half, d_prop.data_ptr(), half, d_z.data_ptr(), d_lpn.data_ptr(), d_logp.data_ptr(), 7, it, d_nacc.data_ptr(), None))
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        tot += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
print("ms per half-update: propose %.3f  logpost %.3f  accept %.3f" % tuple(tot / 10))
t0 = time.perf_counter()
for it in range(20):
    _ffi.check(lib.rscm_b200_stretch_propose(d_pos.data_ptr(), W, P, 0, half, half, half, 2.0, 7, it, d_prop.data_ptr(), half, d_z.data_ptr(), None))
    ens.log_posterior_device(d_prop, sc, d_lpn, layout=0, M=half, S=1)
    _ffi.check(lib.rscm_b200_stretch_accept(d_pos.data_ptr(), W, P, 0, half, d_prop.data_ptr(), half, d_z.data_ptr(), d_lpn.data_ptr(), d_logp.data_ptr(), 7, it, d_nacc.data_ptr(), None))
torch.cuda.synchronize()
print("wall ms per half-update (no per-phase sync): %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
