import sys, numpy as np, torch
sys.path[:0]=['.','local-hyperdb_b200']
import hyperdb_b200 as hb, bench
dev = torch.device("cuda", 0)
n, d, B = 4_000_000, 768, 4096
V = bench.gen_rows_torch(0, n, d, "float16", dev, seed=0)
m = hb.DeviceMatrix(V)
Q = torch.as_tensor(bench.gen_queries(B, d, "float16", seed=5)).to(dev)
idx = torch.empty((B, 10), dtype=torch.int64, device=dev); sc = torch.empty((B, 10), dtype=torch.float64, device=dev)
cnt = torch.empty(B, dtype=torch.int64, device=dev); fl = torch.zeros(B, dtype=torch.int32, device=dev)
for it in range(2):
    m.query_device(Q, 10, "cosine_similarity", 0.0, idx, sc, cnt, fl)
torch.cuda.synchronize()
m.profile_enable(8)
for it in range(3):
    m.query_device(Q, 10, "cosine_similarity", 0.0, idx, sc, cnt, fl)
nl, kms = m.profile_read()
print(f"contraction {kms/nl:.3f} ms = {2.0*n*d*B/(kms/nl*1e-3)/1e12:.1f} TFLOP/s")
