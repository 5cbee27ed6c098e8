"""One dense single-query sweep of an n x d matrix (default 20M x 384 fp16) for ncu: python scratch/staged_probe.py [n] [d] [dtype]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import hyperdb_b200 as hb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
dt = sys.argv[3] if len(sys.argv) > 3 else "float16"
dev = torch.device("cuda", 0)
rows = bench.gen_rows_torch(0, n, d, dt, dev)
m = hb.DeviceMatrix(rows)
q = bench.gen_queries(1, d, dt)[0]
for i in range(3):
    m.query(q, 10, "cosine_similarity")
t = min(m.time_last_query(0, 20) for _ in range(3))
print(f"n={n} d={d} {dt}: sweep {t:.3f} ms = {n * d * bench.ITEM[dt] / t / 1e6:.0f} GB/s", flush=True)
