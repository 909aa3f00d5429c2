#!/usr/bin/env python
"""Why the sampler loop stops at ~70 % strong-scaling efficiency on 8 GPUs: the same per-GPU load (2 x 65 536 members per
iteration) on ONE GPU, no communication at all.  python tools/prof_sampler_small.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

for W in (1 << 20, 1 << 17):
    r = bench.sampler_loop(1, 0, W=W, iters=10)
    print(json.dumps({"walkers": W, "members_per_evaluation": W // 2, "s_per_iteration": r["s_per_iteration"],
                      "member_years_per_s": r["value"]}), flush=True)
