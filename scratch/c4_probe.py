"""Kernel-only timing of the C4-shaped sweep (100M x 384 fp16 by default) under different row subsets: dense, random 50 % mask,
the same documents clustered by category.  usage: python scratch/c4_probe.py [rows]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
import hyperdb_b200 as hb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
d, dev = 384, torch.device("cuda", 0)
q = bench.gen_queries(1, d, "float16")[0]
def run(tag, m, k, bias):
    m.query(q, k, "cosine_similarity", bias)
    t = [m.time_last_query(0, 20) for _ in range(3)]
    print(f"{tag:40s} k={k:3d} sweep ms {min(t):.3f}  kept {m.n_kept}  kept-GB/s {m.n_kept * d * 2 / min(t) / 1e6:.0f}", flush=True)
rows = bench.gen_rows_torch(0, n, d, "float16", dev)
m = hb.DeviceMatrix(rows)
run("dense", m, 10, 0.0); run("dense", m, 100, 0.0)
ts = 1.7e9 + 3600.0 * torch.rand(n, device=dev, dtype=torch.float64)
m.set_timestamps(ts); m.refresh_decay()
run("dense + decay", m, 100, 0.3)
keep = bench.gen_keep_torch(0, n, dev)
m.set_mask(bench.pack_keep_bits(keep)); m.refresh_decay()
run("random mask + decay", m, 100, 0.3)
m.set_timestamps(None)
run("random mask", m, 10, 0.0)
half = torch.zeros(n, dtype=torch.bool, device=dev); half[: n // 2] = True
m.set_mask(bench.pack_keep_bits(half))
run("first half kept (one run)", m, 10, 0.0)
m.set_mask(None); m.set_range(0, n // 2)
run("range [0, n/2)", m, 10, 0.0)
m.set_range(0, n)
cat = bench.gen_category_torch(0, n, dev)
cl = torch.zeros(n, dtype=torch.bool, device=dev)
srt = torch.sort(cat, stable=True)
cl = keep[srt.indices]
m.set_mask(bench.pack_keep_bits(cl))
run("clustered mask (4 runs), no order", m, 10, 0.0)
m.set_row_order(srt.indices.to(torch.int32))
run("clustered mask + row order", m, 10, 0.0)
m.set_timestamps(ts); m.refresh_decay()
run("clustered mask + row order + decay", m, 100, 0.3)
