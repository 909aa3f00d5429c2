#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the headline kernel.

    python tools/update_traffic.py gpurun_out/prof_coupled_r2.ncu-rep --members 262144 --scenarios 8 --source "<how it was captured>"

Writes the DRAM bytes per member-year, the FP64-pipe and issue utilisation, and the hash of the device sources the kernel
was built from (`bench.sources_sha16`).  bench.py reports `roofline.traffic` only while that hash matches the sources it
runs, so a stale capture cannot be quoted for a changed kernel.
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--members", type=int, default=1 << 18)
    ap.add_argument("--scenarios", type=int, default=8)
    ap.add_argument("--years", type=int, default=350)
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    import bench

    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    k = [r for r in rows[2:] if "ensemble_kernel" in r[h.index("Kernel Name")]]
    if not k:
        raise SystemExit("no ensemble_kernel launch in the report")
    r = k[-1]

    def val(name):
        v = float(r[h.index(name)].replace(",", ""))
        unit = rows[1][h.index(name)].lower()
        return v * {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "s": 1e3}.get(unit, 1.0)  # times in ms

    my = a.members * a.scenarios * a.years
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out = {
        "kernel": "ensemble_kernel<double, coupled carbon+two-layer, write>",
        "source": a.source or os.path.basename(a.report),
        "sources_sha16": bench.sources_sha16(),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "member_years_in_capture": my,
        "dram_bytes_per_member_year": (rd + wr) / my, "algorithmic_bytes_per_member_year": 56.16,
        "kernel_ms_under_ncu": val("gpu__time_duration.sum") if "gpu__time_duration.sum" in h else None,
        "ncu_fp64_pipe_active": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
        "ncu_issue_active": val("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
        "registers_per_thread": val("launch__registers_per_thread"),
        "note": "bench.py multiplies dram_bytes_per_member_year by the member-years of its own launch and quotes it only while "
                "sources_sha16 matches the device sources it runs",
    }
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=2)
    print(json.dumps(out, indent=2))


if __name__ == "__main__":
    main()
