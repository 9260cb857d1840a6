#!/usr/bin/env python
"""Print the per-kind time table of a bench.py JSON line: python tools/print_kinds.py gpurun_out/bench.log"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  launches/step {d['launches_per_step']}  "
      f"conv {r['achieved']} TFLOP/s ({r['frac']})  network eager {r['network_ms_eager']} ms")
for k, v in r["by_kind"].items():
    print(f"  {k:20s} {v['ms']:8.3f} ms  {v['launches']:4d} launches  {v['share'] * 100:5.1f} %  {v['algo_GBps']} GB/s algorithmic")
