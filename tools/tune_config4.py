#!/usr/bin/env python
"""Config 4 (MAGICC boxes + ClimateUDEB, 100 000 members) throughput for a few register budgets of the lane-quad kernel
(RSCM_B200_LANES_MIN_BLOCKS = CTAs per SM the kernel is compiled for), with parity of a subsample against the oracle."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def one():
    import numpy as np, torch
    from rscm_b200 import synthetic as syn
    from tests.helpers import oracle_bindings, oracle_from_builder, rel_err
    M = int(os.environ.get("M", 100000))
    b, binds, params, scen = syn.config4(M=M)
    for dt in ("f64", "f32"):
        ens = b.build_ensemble(dtype=dt).bind_parameters(binds)
        ens.select_outputs(syn.CONFIG4_OUTPUTS)
        sc_host = ens.pack_scenarios(scen)
        d_p = torch.from_numpy(np.ascontiguousarray(params.T)).cuda()
        d_s = torch.from_numpy(sc_host).cuda()
        d_o = torch.empty((ens.output_rows, M), dtype=torch.float64, device="cuda")
        for _ in range(2):
            ens.run_device(d_p, d_s, d_o, layout=0)
        torch.cuda.synchronize(); ens.kernel_ms(reset=True)
        for _ in range(3):
            ens.run_device(d_p, d_s, d_o, layout=0)
        torch.cuda.synchronize()
        ms = ens.kernel_ms()
        idx = np.arange(0, M, max(1, M // 200))
        sub = d_o[:, torch.from_numpy(idx).cuda()].cpu().numpy()
        m = oracle_from_builder(b)
        ref = m.run_batch(oracle_bindings(b, binds), params[idx], ens.exogenous_names, sc_host, syn.CONFIG4_OUTPUTS)
        got, want = ens.split_outputs(sub), m.split(ref, syn.CONFIG4_OUTPUTS)
        err = max(rel_err(got[n], want[n]) for n in syn.CONFIG4_OUTPUTS)
        print(json.dumps({"min_blocks": os.environ.get("RSCM_B200_LANES_MIN_BLOCKS", "default"), "stage_exo": os.environ.get("RSCM_B200_LANES_STAGE_EXO", "default"),
                          "smem": ens.shared_bytes(), "dtype": dt, "members": M, "kernel_ms": ms,
                          "member_years_per_s": M * 350 / (ms * 1e-3), "max_rel_err": err}), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for mb in (sys.argv[1:] or ["2", "3", "4"]):       # "3" or "4:0" = CTAs per SM [: exogenous rows staged in shared memory 0/1]
            mb, _, st = mb.partition(":")
            env = dict(os.environ, RSCM_B200_LANES_MIN_BLOCKS=mb, RSCM_B200_CACHE=f"/tmp/rscm_cache_mb{mb}_{st}")
            if st:
                env["RSCM_B200_LANES_STAGE_EXO"] = st
            subprocess.run([sys.executable, __file__, "one"], env=env, check=False)
