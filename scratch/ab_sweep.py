"""A/B of the sweep kernel alone (hdb_time_last_query, kernel-only loop) between two builds of the library, alternating
processes on the same box:  python scratch/ab_sweep.py <pkgdir_a> <pkgdir_b> [workload]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, sys.argv[1])
import torch, hyperdb_b200 as hb
sys.path.insert(0, %r)
import bench
import os; assert os.path.abspath(sys.argv[1]) in hb.__file__, hb.__file__
w = bench.WORKLOADS[sys.argv[2]]
dev = torch.device("cuda", 0)
rows = bench.gen_rows_torch(0, w["n"], w["d"], w["dtype"], dev)
m = hb.DeviceMatrix(rows)
q = bench.gen_queries(1, w["d"], w["dtype"])[0]
m.query(q, w["k"], w["metric"])
res = []
for rep in range(3):
    res.append(m.time_last_query(0, 100))
print(json.dumps(res))
''' % ROOT
a, b = sys.argv[1], sys.argv[2]
wl = sys.argv[3] if len(sys.argv) > 3 else "c3_cosine_b1"
for rnd in range(3):
    for tag, pkg in (("A", a), ("B", b)):
        out = subprocess.run([sys.executable, "-c", CHILD, pkg, wl], capture_output=True, text=True)
        print(tag, pkg, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:], flush=True)
