#!/usr/bin/env python
"""bench.py — ensemble member-years/second on BASELINE.json's coupled carbon+two-layer config.

    python bench.py --gpus N --steps K --warmup W             (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the oracle port)

A "step" = one pass of the hot path over the whole batch: M members x S scenarios x 350 years
(`ModelRunner::run_batch` of the coupled graph).  One JSON line on stdout (rank 0).

  value        device-resident throughput: params + scenarios already in HBM, all 7 output series
               written to HBM ([rows][runs] fp64), CUDA-event timed per step, L2 flushed between steps
  e2e          the same workload through the host C-ABI call (rscm_b200_run_host): parameters from
               pinned host memory, the `output_variables` the runner was asked for copied back to
               pinned host memory, copies inside the timed region
  roofline     FP64-pipe roofline of the fused kernel: algorithmic flop (1710 per
               member-year) / kernel time vs the DFMA peak measured in this run; HBM-write roofline beside it
  cpu_baseline the CPU oracle (port of the reference arithmetic) on a bounded sample, rank 0, N = 1
  parity       the LAST TIMED launch's output, a strided subsample of >= 4096 runs, all 7 series, against the CPU
               oracle (max |gpu - cpu| / max |cpu| per series <= 1e-9, NaN positions equal); max over ranks
  strong       the same global ensemble (262144 members x 8 scenarios) split by member over the N GPUs (BASELINE
               configs[2] "sharded over 1/2/4/8"), timed like `value`; plus a bit-identity check of every rank's block
               against a single-GPU recomputation of the same members
  e2e_summary  end to end with across-member quantiles returned instead of member series (112 KB back instead of 5.9 GB)
  secondary    BASELINE configs[1], [3], [4]: value, kernel ms and parity per config; config 5 also as a sampler loop with
               the all-gather of log-posteriors over the N GPUs
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_MEMBER_YEAR = 1710.0   # coupled graph, as written in the reference
BYTES_PER_MEMBER_YEAR = 56.0    # 7 fp64 output series
YEARS = 350


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--members", type=int, default=1 << 18, help="members per GPU (BASELINE config: 262144)")
    ap.add_argument("--scenarios", type=int, default=8)
    ap.add_argument("--e2e-outputs", default="Surface Temperature", help="comma list, or 'all'")
    ap.add_argument("--cpu-sample-members", type=int, default=0, help="0 = auto-size to ~10 s")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs 2/4/5 block")
    ap.add_argument("--parity-runs", type=int, default=4096, help="runs of the timed launch checked against the oracle")
    return ap.parse_args()


class ClockSampler:
    """SM clocks + throttle reasons polled DURING the timed region (B200_PROFILING.md recipe).  NVML in a thread every
    ~5 ms (the timed region is ~150 ms: `nvidia-smi -lms 100` would see one sample); nvidia-smi as the fallback."""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc, self.nvml, self._stop = None, None, threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map torch's device through its UUID (CUDA_VISIBLE_DEVICES may reorder)
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            self.handle = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                if uuid in (u.decode() if isinstance(u, bytes) else u):
                    self.handle = h
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.source = "nvidia-smi"
            self.lines = []
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=2)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    self.sm.append(float(f[0])); self.mx.append(float(f[1]))
                except ValueError:
                    continue
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


WORKLOAD = "coupled carbon cycle + CO2 ERF + two-layer, 262144 members x 8 scenarios x 350 yr (BASELINE configs[2])"


def workload_config(args) -> dict:
    """The `config` object — identical in both arms (the driver compares them)."""
    e2e_names = "all 7 series" if args.e2e_outputs == "all" else args.e2e_outputs
    return {
        "workload": WORKLOAD, "members_per_gpu": args.members, "scenarios": args.scenarios, "years": YEARS,
        "outputs": "all 7 series, fp64, [rows][runs]",
        "l2": "256 MB flush between timed steps; outputs per step (%.1f GB) also exceed L2"
              % (7 * (YEARS + 1) * args.members * args.scenarios * 8 / 1e9),
        "timing": "GPU arm: CUDA events around each step, summed, max over ranks; CPU arm: wall clock per step",
        "e2e_output_variables": e2e_names,
    }


def host_threads() -> int:
    """Every core this process may run on — NOT what OMP_NUM_THREADS says (torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which would time the 'all host cores' arm on one core)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_leg(args, M_sample, steps, warmup):
    """The CPU restatement (oracle) timed with all host threads on a bounded sample of the same workload.  Built from
    oracle/workloads.py + rscm_b200/synthetic_data.py only: the CUDA library is not mapped into a process that runs
    nothing but this leg."""
    from oracle import workloads as wl

    sd = wl.synthetic_data()
    m, ob = wl.coupled_model(sd)
    params = sd.config3_params(M_sample)
    em = wl.coupled_scenarios(sd, args.scenarios)
    threads = host_threads()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        m.run_batch(ob, params, [sd.COUPLED_EXOGENOUS], em, sd.COUPLED_OUTPUTS, n_threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    my = M_sample * args.scenarios * YEARS
    return my / (sum(times) / len(times)), threads, sum(times) / len(times)


def auto_cpu_sample(args, target_s=10.0):
    v, cores, dt = cpu_leg(args, 256, 1, 0)
    M = int(max(256, min(args.members, target_s * v / (args.scenarios * YEARS))))
    return (M // 64) * 64


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    M = args.cpu_sample_members or auto_cpu_sample(args, target_s=8.0)
    v, cores, dt = cpu_leg(args, M, args.steps, min(args.warmup, 1))
    sample = f"{M} members x {args.scenarios} scenarios x {YEARS} years per step (bounded sample of {args.members} x {args.scenarios})"
    line = {
        "impl": "reference", "metric": "ensemble member-years/sec, coupled carbon+two-layer", "value": v, "unit": "member-years/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "member-years/s", "cores": cores, "kind": "port", "sample": sample,
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"), "host_cpus": os.cpu_count()},
        "e2e": {"value": v, "unit": "member-years/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cuda_library_mapped": any("librscm_b200" in ln for ln in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
        "note": "reference is Rust (no toolchain here): this arm is the C oracle restating its arithmetic, OpenMP over members on every "
                "core the process may use (thread count set explicitly, OMP_NUM_THREADS ignored); it omits the reference's per-step "
                "string/HashMap overhead and per-member model rebuild, so it is faster than the real reference",
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def sources_sha16() -> str:
    """Hash of the device sources the headline kernel is built from: keys profiles/traffic.json to the code it measured."""
    import hashlib
    h = hashlib.sha256()
    for f in ("kernel.cuh", "components.cuh", "math_tables.cuh"):
        h.update(open(os.path.join(ROOT, "rscm_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def bind_to_gpu_numa(local: int) -> dict:
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host buffer is allocated (first touch decides
    where the pages live): eight ranks copying 5.9 GB each through one node's memory controller is what capped the
    round-1 end-to-end number."""
    import torch
    info = {"bound": False}
    try:
        prop = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        info.update({"pci": bus, "numa_node": node, "local_cpus": len(ids)})
        if ids:
            os.sched_setaffinity(0, ids)
            info["bound"] = True
    except Exception as e:  # containers without /sys topology: stay unbound
        info["error"] = str(e)[:120]
    return info


def series_rel_err(got: dict, want: dict, names) -> tuple[float, bool, dict]:
    """per series max |gpu - cpu| / max |cpu| over the sub-sample; NaN positions must match exactly."""
    worst, nan_ok, per = 0.0, True, {}
    for n in names:
        a, e = np.asarray(got[n]), np.asarray(want[n])
        nan_ok = nan_ok and bool(np.array_equal(np.isnan(a), np.isnan(e)))
        ok = ~np.isnan(e) & ~np.isnan(a)
        err = float(np.max(np.abs(a[ok] - e[ok])) / max(np.max(np.abs(e[ok])), 1e-300)) if ok.any() else 0.0
        per[n] = err
        worst = max(worst, err)
    return worst, nan_ok, per


def subsample_columns(M: int, S: int, n_per_scenario: int):
    """Strided member indices and the matching run columns of an out[rows][S*M] block (run = s*M + m)."""
    n = max(1, min(M, n_per_scenario))
    idx = (np.arange(n) * (M // n)).astype(np.int64)
    cols = np.concatenate([idx + s * M for s in range(S)])
    return idx, cols


def oracle_check(builder, binds, ens, d_out, params, sc_host, outputs, M, S, n_per_scenario, tol):
    """Compare a strided sub-sample of a device output block with the CPU oracle on the same members."""
    import torch
    from tests.helpers import oracle_bindings, oracle_from_builder

    idx, cols = subsample_columns(M, S, n_per_scenario)
    sub = d_out[:, torch.from_numpy(cols).to(d_out.device)].cpu().numpy()
    m = oracle_from_builder(builder)
    ref = m.run_batch(oracle_bindings(builder, binds), params[idx], ens.exogenous_names, sc_host, outputs, n_threads=host_threads())
    worst, nan_ok, per = series_rel_err(ens.split_outputs(sub), m.split(ref, outputs), outputs)
    return {"max_rel_err": worst, "nan_positions_match": nan_ok, "n": int(cols.size), "tol": tol, "ok": bool(nan_ok and worst <= tol),
            "series": len(outputs), "per_series": per}


def time_device(fn, steps, warmup, flush, barrier):
    """CUDA events around each call on the current stream, L2 flushed before each; returns summed ms over `steps`."""
    import torch
    for _ in range(max(warmup, 3)):
        fn()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        fn()
        e1.record()
    barrier()
    return sum(e0.elapsed_time(e1) for e0, e1 in ev)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else {"bound": False, "note": "single rank: not bound"}

    from rscm_b200 import _ffi, synthetic as syn
    from rscm_b200.dist import member_shard

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    M, S = args.members, args.scenarios
    b, binds, params0, scen = syn.config3(M=M, S=S)
    params = params0 if world == 1 else syn.config3_params(M, rank, world)   # weak scaling: every rank its own member block
    ens = b.build_ensemble(device=local).bind_parameters(binds)
    ens.select_outputs(syn.COUPLED_OUTPUTS)
    sc_host = ens.pack_scenarios(scen)
    runs = S * M
    d_params = torch.from_numpy(np.ascontiguousarray(params.T)).to(dev)      # [cols][M] SoA
    d_scen = torch.from_numpy(sc_host).to(dev)
    d_out = torch.empty((ens.output_rows, runs), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)            # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        ens.run_device(d_params, d_scen, d_out, layout=0)

    # ---- headline: weak scaling, device-resident ---------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ens.kernel_ms(reset=True)
    launches0 = ens.launch_count()
    clocks = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        step()
        e1.record()
    barrier()
    total_ms = max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in ev))
    kernel_ms = ens.kernel_ms(reset=True)
    launches = ens.launch_count() - launches0
    clk = clocks.stop() if clocks else None
    my_per_step = world * runs * YEARS
    value = my_per_step * args.steps / (total_ms * 1e-3)

    # ---- parity of the last timed launch against the CPU oracle (every rank its own block) ----------------------
    par = oracle_check(b, binds, ens, d_out, params, sc_host, syn.COUPLED_OUTPUTS, M, S, max(1, args.parity_runs // S), 1e-9)
    par["max_rel_err"] = max_over_ranks(par["max_rel_err"])
    par["ok"] = bool(max_over_ranks(0.0 if par["ok"] else 1.0) == 0.0)
    par["ranks_checked"] = world
    par["what"] = "last timed launch, strided members of every scenario, all 7 series vs the CPU oracle on the same inputs"

    # ---- strong scaling: the single-GPU ensemble split by member ------------------------------------------------
    strong = {"value": value, "ms_per_step": total_ms / args.steps, "members_total": M, "members_per_gpu": M,
              "note": "N = 1: identical to the headline run"}
    if world > 1:
        lo, hi = member_shard(M, rank, world)
        Ml = hi - lo
        d_pl = torch.from_numpy(np.ascontiguousarray(params0[lo:hi].T)).to(dev)
        d_ol = d_out.view(-1)[: ens.output_rows * S * Ml].view(ens.output_rows, S * Ml)
        s_ms = max_over_ranks(time_device(lambda: ens.run_device(d_pl, d_scen, d_ol, layout=0), args.steps, args.warmup, flush, barrier))
        # bit identity: n_chk members of every rank's block, recomputed by every rank in one single-GPU launch
        n_chk = min(64, Ml)
        kidx = (np.arange(n_chk) * (Ml // n_chk)).astype(np.int64)
        cols = torch.from_numpy(np.concatenate([kidx + s * Ml for s in range(S)])).to(dev)
        mine = d_ol[:, cols].contiguous()                                       # [rows][S*n_chk]
        gathered = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, mine)
        shards = [member_shard(M, r, world) for r in range(world)]
        members = None
        if all(b1 - b0 == Ml for b0, b1 in shards):   # equal blocks: every rank picked the same strided offsets
            members = np.concatenate([b0 + kidx for b0, _ in shards])
        identical = None
        if members is not None:
            d_pc = torch.from_numpy(np.ascontiguousarray(params0[members].T)).to(dev)
            d_oc = torch.empty((ens.output_rows, S * members.size), dtype=torch.float64, device=dev)
            ens.run_device(d_pc, d_scen, d_oc, layout=0)
            torch.cuda.synchronize()
            want = d_oc.view(ens.output_rows, S, world, n_chk)
            got = gathered.view(world, ens.output_rows, S, n_chk).permute(1, 2, 0, 3)
            identical = bool(torch.equal(got.contiguous().view(torch.int64), want.contiguous().view(torch.int64)))
            identical = bool(max_over_ranks(0.0 if identical else 1.0) == 0.0)
            del d_oc, d_pc
        spar = oracle_check(b, binds, ens, d_ol, params0[lo:hi], sc_host, syn.COUPLED_OUTPUTS, Ml, S, max(1, args.parity_runs // S // world), 1e-9)
        strong = {"value": M * S * YEARS * args.steps / (s_ms * 1e-3), "ms_per_step": s_ms / args.steps, "members_total": M,
                  "members_per_gpu": Ml, "scaling": "strong",
                  "sharded_bit_identical_to_single_gpu": identical, "bit_identity_runs": int(world * n_chk * S),
                  "parity_max_rel_err": max_over_ranks(spar["max_rel_err"]), "parity_n": spar["n"] * world}
        del d_pl, mine, gathered

    # ---- e2e: host buffers through rscm_b200_run_host --------------------------------------------------------------
    e2e_names = syn.COUPLED_OUTPUTS if args.e2e_outputs == "all" else [s.strip() for s in args.e2e_outputs.split(",")]
    ens.select_outputs(e2e_names)
    h_params = torch.from_numpy(np.ascontiguousarray(params)).pin_memory()   # [M][cols] = the reference's &[Vec<f64>]
    h_scen = torch.from_numpy(sc_host).pin_memory()
    h_out = torch.empty((ens.output_rows, runs), dtype=torch.float64).pin_memory()
    hp, hs, ho = h_params.numpy(), h_scen.numpy(), h_out.numpy()
    for _ in range(2):
        ens.run(hp, hs, layout=1, out=ho)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ens.run(hp, hs, layout=1, out=ho)
    barrier()
    e2e_value = my_per_step * args.steps / max_over_ranks(time.perf_counter() - t0)
    h2d = hp.nbytes + hs.nbytes
    d2h = ho.nbytes
    # the host path returned what the device path computed for the same members (bitwise, sub-sample of columns)
    e2e_same = None
    if world == 1:   # (at N > 1 the strong-scaling run has reused d_out)
        idx, cols = subsample_columns(M, S, 64)
        row0 = ens.output_layout()[e2e_names[0]][0]
        ens.select_outputs(syn.COUPLED_OUTPUTS)
        full_row0 = ens.output_layout()[e2e_names[0]][0]
        dcheck = d_out[full_row0:full_row0 + (YEARS + 1), torch.from_numpy(cols).to(dev)].cpu().numpy()
        e2e_same = bool(np.array_equal(ho[row0:row0 + YEARS + 1][:, cols], dcheck, equal_nan=True))
    del h_out, ho

    # ---- e2e_summary: across-member quantiles back instead of member series -----------------------------------
    ens.select_outputs(e2e_names)
    q = [0.05, 0.17, 0.5, 0.83, 0.95]
    rows_q = ens.output_rows
    d_oq = d_out.view(-1)[: rows_q * runs].view(rows_q, runs)
    d_pq = torch.empty((params.shape[1], M), dtype=torch.float64, device=dev)
    d_res = torch.empty((len(q), rows_q, S), dtype=torch.float64, device=dev)
    h_res = torch.empty((len(q), rows_q, S), dtype=torch.float64).pin_memory()
    h_params_soa = torch.from_numpy(np.ascontiguousarray(params.T)).pin_memory()

    def summary_step():
        d_pq.copy_(h_params_soa, non_blocking=True)
        d_sq = h_scen.to(dev, non_blocking=True)
        ens.run_device(d_pq, d_sq, d_oq, layout=0)
        ens.member_quantiles_device(d_oq, q, d_res, M=M, S=S)
        h_res.copy_(d_res, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        summary_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        summary_step()
    barrier()
    e2e_summary_value = my_per_step * args.steps / max_over_ranks(time.perf_counter() - t0)
    e2e_summary = {"value": e2e_summary_value, "unit": "member-years/s", "h2d_bytes_per_step": int(h_params_soa.numpy().nbytes + hs.nbytes),
                   "d2h_bytes_per_step": int(h_res.numpy().nbytes),
                   "api": "Ensemble.run_quantiles path: pinned host params in, rscm_b200_run_device + rscm_b200_member_quantiles, "
                          "%d quantiles x %d rows x %d scenarios back" % (len(q), rows_q, S)}
    ens.select_outputs(syn.COUPLED_OUTPUTS)

    # ---- secondary configs (rank-local for 2 and 4; config 5's loop runs over all ranks) -------------------------
    secondary = None
    if not args.no_secondary:
        del d_params, d_pq, d_res
        secondary = run_secondary(args, d_out, flush, barrier, max_over_ranks, rank, world, local)

    if rank == 0:
        peak_tf = _ffi.C.c_double(0.0)
        _ffi.check(_ffi.lib.rscm_b200_measure_fma_peak(local, 0, _ffi.C.byref(peak_tf)))
        hbm_peak, hbm_src = measured_peaks()
        per_launch_my = runs * YEARS
        achieved_tf = FLOP_PER_MEMBER_YEAR * per_launch_my / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else None
        achieved_gbs = BYTES_PER_MEMBER_YEAR * per_launch_my / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
        traffic, ncu_pipe, traffic_note = None, None, "profiles/traffic.json missing"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                prof = json.load(open(tp))
                if prof.get("sources_sha16") == sources_sha16():
                    traffic = prof["dram_bytes_per_member_year"] * per_launch_my
                    traffic_note = "ncu --set full capture of this kernel source (hash matches), scaled to this launch"
                    ncu_pipe = {"fp64_pipe_active": prof.get("ncu_fp64_pipe_active"), "issue_active": prof.get("ncu_issue_active"),
                                "source": prof.get("source")}
                else:
                    traffic_note = "stale: kernel sources changed since the ncu capture in profiles/traffic.json (hash %s != %s)" % (
                        prof.get("sources_sha16"), sources_sha16())
            except Exception:
                traffic = None
        line = {
            "metric": "ensemble member-years/sec, coupled carbon+two-layer", "value": value, "unit": "member-years/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "member-years/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "rscm_b200_run_host (ModelRunner.run_batch): pinned host params [M][cols] in, selected output series out",
                    "host_output_equals_device_output": e2e_same, "numa": numa},
            "e2e_summary": e2e_summary,
            "parity": par,
            "strong": strong,
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf.value, "unit": "TFLOP/s",
                         "frac": (achieved_tf / peak_tf.value) if (achieved_tf and peak_tf.value) else None, "traffic": traffic,
                         "traffic_note": traffic_note,
                         "kernel": "ensemble_kernel<double, coupled, write>", "kernel_ms": kernel_ms,
                         "algorithmic": "1710 flop per member-year (SURVEY.md 8d) x %d member-years per launch" % per_launch_my,
                         "peak_source": "DFMA micro-benchmark in this run (rscm_b200_measure_fma_peak); MEASURED_PEAKS.json has no FP64 entry",
                         "note": "frac > 1 is possible: the kernel executes fewer FP64 operations than the reference's expression count "
                                 "(DESIGN.md 2.2); the committed ncu capture gives the pipe's own utilisation", "ncu": ncu_pipe},
            "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (achieved_gbs / hbm_peak) if achieved_gbs else None, "peak_source": hbm_src,
                             "algorithmic": "56 B written per member-year"},
            "secondary": secondary,
        }
        if world == 1 and not args.no_cpu_baseline:
            Ms = args.cpu_sample_members or auto_cpu_sample(args)
            v, cores, dt = cpu_leg(args, Ms, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": "member-years/s", "cores": cores, "kind": "port",
                                    "sample": f"{Ms} members x {S} scenarios x {YEARS} yr, 1 pass ({dt:.1f} s), OpenMP static over members"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_secondary(args, d_buf, flush, barrier, max_over_ranks, rank, world, local):
    """BASELINE configs[1] (two-layer 1M), [3] (MAGICC boxes, 100k, fp64 and fp32; and the eleven-box emissions-driven chain) and [4] (log-posterior of 1M members and
    the sampler loop with its all-gather).  Configs 2 and 4 do not communicate: at N > 1 every rank runs the same shape
    (weak) and the line reports N x the slowest rank.  Each entry carries kernel time and parity against the CPU oracle."""
    import torch

    from rscm_b200 import _ffi, synthetic as syn

    dev = torch.device("cuda", local)
    res = {}

    def one(name, builder, binds, params, scen, outputs, dtype, n_sub, tol, years=YEARS):
        ens = builder.build_ensemble(dtype=dtype, device=local).bind_parameters(binds)
        ens.select_outputs(outputs)
        sc_host = ens.pack_scenarios(scen)
        Mc, Sc = params.shape[0], len(scen)
        d_p = torch.from_numpy(np.ascontiguousarray(params.T)).to(dev)
        d_s = torch.from_numpy(sc_host).to(dev)
        d_o = d_buf.view(-1)[: ens.output_rows * Sc * Mc].view(ens.output_rows, Sc * Mc)
        ms = max_over_ranks(time_device(lambda: ens.run_device(d_p, d_s, d_o, layout=0), 3, 3, flush, barrier)) / 3
        kms = ens.kernel_ms(reset=True)
        par = oracle_check(builder, binds, ens, d_o, params, sc_host, outputs, Mc, Sc, n_sub, tol)
        par.pop("per_series")
        res[name] = {"value": world * Mc * Sc * years / (ms * 1e-3), "unit": "member-years/s", "ms_per_step": ms, "kernel_ms": kms,
                     "members_per_gpu": Mc, "scenarios": Sc, "years": years, "dtype": dtype, "jit": ens.program_is_jit(), "parity": par}
        ens.close()

    b, binds, params, scen = syn.config2(M=1 << 20)
    one("config2_two_layer_1M", b, binds, params, scen, ["Surface Temperature", "Deep Ocean Temperature"], "f64", 4096, 1e-9)
    b4, binds4, params4, scen4 = syn.config4(M=100_000)
    one("config4_magicc_boxes_100k_f64", b4, binds4, params4, scen4, syn.CONFIG4_OUTPUTS, "f64", 256, 1e-9)
    one("config4_magicc_boxes_100k_f32", b4, binds4, params4, scen4, syn.CONFIG4_OUTPUTS, "f32", 256, 1e-4)
    # configs[3] at its widest: all eleven rscm-magicc boxes in one emissions-driven graph (124 variables, run-time compiled)
    bf, bindsf, paramsf, scenf = syn.full_chain(M=100_000)
    one("config4_full_chain_11_boxes_100k_f64", bf, bindsf, paramsf, scenf, syn.FULL_CHAIN_OUTPUTS, "f64", 64, 1e-9, years=250)

    # ---- config 5: log-posterior of 1M two-layer members, then the sampler loop ------------------------------------
    from tests.helpers import oracle_bindings, oracle_from_builder
    ens = b.build_ensemble(device=local).bind_parameters(binds)
    sc_host = ens.pack_scenarios(scen)
    truth = dict(syn.TWO_LAYER_DEFAULTS, lambda0=1.1, efficacy=1.3, a=0.05)
    ens.select_outputs(["Surface Temperature"])
    t_true = ens.run(np.array([[truth[k] for k in syn.TWO_LAYER_RANGES]]), sc_host)[:, 0]
    obs = syn.config5_observations(t_true, syn.time_axis().values())
    priors = [(_ffi.PRIOR_UNIFORM, lo, hi) for lo, hi in syn.TWO_LAYER_RANGES.values()]
    ens.set_target(obs).set_priors(priors)
    Mc = params.shape[0]
    d_p = torch.from_numpy(np.ascontiguousarray(params.T)).to(dev)
    d_s = torch.from_numpy(sc_host).to(dev)
    d_lp = torch.empty(Mc, dtype=torch.float64, device=dev)
    d_sum = torch.zeros(5, dtype=torch.float64, device=dev)
    ms = max_over_ranks(time_device(lambda: ens.log_posterior_device(d_p, d_s, d_lp, d_sum, layout=0), 3, 3, flush, barrier)) / 3
    kms = ens.kernel_ms(reset=True)
    idx = np.arange(0, Mc, 257)
    m = oracle_from_builder(b)
    ref = m.log_posterior_batch(oracle_bindings(b, binds), params[idx], ens.exogenous_names, sc_host, priors, obs, n_threads=host_threads())
    got = d_lp.cpu().numpy()[idx]
    fin = np.isfinite(ref)
    err = float(np.max(np.abs(got[fin] - ref[fin]) / np.abs(ref[fin]))) if fin.any() else 0.0
    res["config5_logpost_1M"] = {"value": world * Mc * YEARS / (ms * 1e-3), "unit": "member-years/s", "ms_per_step": ms, "kernel_ms": kms,
                                 "members_per_gpu": Mc, "observations": len(obs), "bytes_out_per_member": 8,
                                 "parity": {"max_rel_err": err, "inf_positions_match": bool(np.array_equal(np.isinf(got), np.isinf(ref))),
                                            "n": int(idx.size), "tol": 1e-9, "ok": bool(err <= 1e-9)}}
    ens.close()
    res["config5_sampler_loop"] = sampler_loop(world, local)
    return res


def sampler_loop(world, local, W=1 << 20, iters=10):
    """BASELINE configs[4] as the loop it is: stretch-move iterations over W walkers, walker state resident in HBM, the
    log-posterior evaluation of each half-update sharded by member over the N GPUs with the all-gather of the 8-byte costs
    (strong scaling: W is fixed)."""
    import torch

    from rscm_b200 import synthetic as syn
    from rscm_b200.calibrate import DeviceEnsembleSampler, GaussianLikelihood, ModelRunner, ParameterSet, Target, Uniform, WalkerInit

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
    distributed = world > 1
    s.run(3, WalkerInit.from_prior(), n_walkers=W, thin=1000, seed=5, distributed=distributed)   # warm-up (NCCL, allocations)
    wall = []
    for k in (iters, 3 * iters):   # two run lengths: the difference removes walker initialisation and the final read-back
        torch.cuda.synchronize()
        if distributed:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        s.run(k, WalkerInit.from_prior(), n_walkers=W, thin=100000, seed=6, distributed=distributed)
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
    dt = torch.tensor([(wall[1] - wall[0]) / (2 * iters)], dtype=torch.float64, device="cuda")
    if distributed:
        torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
    dt = float(dt.item())
    return {"value": W * YEARS / dt, "unit": "member-years/s", "s_per_iteration": dt, "walkers": W, "n_gpus": world, "scaling": "strong",
            "acceptance_rate": s.acceptance_rate, "collective": getattr(s, "collective", "torch.distributed all_gather_into_tensor (NCCL)") if distributed else None,
            "note": "per iteration: 2 half-updates = W log-posterior evaluations (+ 2 all-gathers at N > 1); steady state from two run lengths"}


def emit(line: dict) -> None:
    """The JSON line goes to the process's real stdout; everything else that writes to fd 1
    (e.g. NCCL's version banner) was redirected to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
