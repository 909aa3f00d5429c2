"""bench.py's JSON contract, checked on CPU through the `--impl reference` arm (the oracle timed on the host cores): one JSON
line on stdout with the keys the driver reads; non-zero ranks of a multi-rank launch stay silent."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-members", "256")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("ensemble member-years/sec") and d["unit"] == "member-years/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert "workload" in d["config"] and "BASELINE configs[2]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the CPU arm never maps the CUDA library (its inputs come from rscm_b200/synthetic_data.py loaded by path)
    assert d["cuda_library_mapped"] is False


def test_reference_arm_uses_every_core_whatever_omp_num_threads_says():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the 'all host cores' arm must not inherit it."""
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample-members", "256",
                  env={"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["cpu_baseline"]["omp_num_threads_env"] == "1"


def test_both_arms_emit_the_same_config_object():
    sys.path.insert(0, ROOT)
    import bench

    class A:
        members, scenarios, e2e_outputs = 1 << 18, 8, "Surface Temperature"
    cfg = bench.workload_config(A)
    assert set(cfg) >= {"workload", "members_per_gpu", "scenarios", "years", "outputs", "l2", "timing"}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(args)') == 2   # GPU arm and CPU arm


def test_oracle_workload_description_equals_the_product_builder():
    """oracle/workloads.py states the headline graph without the product package; it must be the same model."""
    import numpy as np

    from oracle import workloads as wl
    from rscm_b200 import synthetic as syn
    from tests.helpers import oracle_bindings, oracle_from_builder

    sd = wl.synthetic_data()
    m1, ob1 = wl.coupled_model(sd)
    b, binds, params, scen = syn.config3(M=64, S=3)
    m2 = oracle_from_builder(b)
    ob2 = oracle_bindings(b, binds)
    assert ob1 == ob2 and m1.names == m2.names and m1.execution_order() == m2.execution_order()
    em = wl.coupled_scenarios(sd, 3)
    assert np.array_equal(em, np.stack([s["Emissions|CO2|Anthropogenic"] for s in scen]))
    assert np.array_equal(sd.config3_params(64), params)
    o1 = m1.run_batch(ob1, params, [sd.COUPLED_EXOGENOUS], em, sd.COUPLED_OUTPUTS)
    o2 = m2.run_batch(ob2, params, ["Emissions|CO2|Anthropogenic"], em, syn.COUPLED_OUTPUTS)
    assert np.array_equal(o1, o2, equal_nan=True)


def test_parity_helpers():
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench

    idx, cols = bench.subsample_columns(1000, 3, 10)
    assert idx.tolist() == list(range(0, 1000, 100)) and cols.size == 30 and cols[10] == 1000 and cols[-1] == 2900
    a = {"x": np.array([[1.0, np.nan], [2.0, 4.0]])}
    e = {"x": np.array([[1.0, np.nan], [2.0, 4.0 + 4e-9]])}
    worst, nan_ok, per = bench.series_rel_err(a, e, ["x"])
    assert nan_ok and abs(worst - 1e-9) < 1e-12
    e["x"][0, 1] = 0.0
    assert bench.series_rel_err(a, e, ["x"])[1] is False


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
