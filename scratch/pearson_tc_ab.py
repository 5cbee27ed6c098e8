"""A/B of batched pearson_correlation: tensor-core path (automatic) against the multi-query streaming sweep (set_path(2), what
every pearson batch took before), same matrix, same queries, host queries in -> host top-10 out (DeviceMatrix.query).
Writes gpurun_out/r02_pearson_tc.json.  Usage: python scratch/pearson_tc_ab.py [rows] [dim] [batch]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "local-hyperdb_b200")]
import numpy as np
import torch
import hyperdb_b200 as hb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(7)
V = torch.empty(n, d, dtype=torch.float16, device=dev)
for r0 in range(0, n, 1_000_000):
    r1 = min(n, r0 + 1_000_000)
    x = torch.randn(r1 - r0, d, generator=g, device=dev)
    V[r0:r1] = (x / x.norm(dim=1, keepdim=True)).half()
Q = torch.randn(B, d, generator=g, device=dev)
Q = (Q / Q.norm(dim=1, keepdim=True)).half().cpu().numpy()
m = hb.DeviceMatrix(V)
out = {"rows": n, "dim": d, "batch": B, "dtype": "float16", "metric": "pearson_correlation", "k": 10}
res = {}
for name, mode in (("tensor", 0), ("sweep_mq", 2)):
    m.set_path(mode)
    r = m.query(Q, 10, "pearson_correlation")          # warm-up (builds the pearson columns, allocates workspaces)
    torch.cuda.synchronize()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        r = m.query(Q, 10, "pearson_correlation")
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    res[name] = r
    flags = r[3]
    out[name] = {"ms_per_batch": dt * 1e3, "queries_per_s": B / dt, "tensor_flag": int(sum(1 for f in flags if f & 4)),
                 "exact_fallbacks": int(sum(1 for f in flags if f & 1))}
out["identical"] = bool(np.array_equal(res["tensor"][0], res["sweep_mq"][0]) and np.array_equal(res["tensor"][1], res["sweep_mq"][1]))
out["speedup"] = out["sweep_mq"]["ms_per_batch"] / out["tensor"]["ms_per_batch"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r02_pearson_tc.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
m.close()
