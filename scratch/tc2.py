import sys, time, numpy as np, torch
sys.path[:0]=['.','local-hyperdb_b200']
import hyperdb_b200 as hb, bench
from hyperdb_b200 import _native as N
dev = torch.device("cuda", 0)
n, d = 10_000_000, 768
V = bench.gen_rows_torch(0, n, d, "float16", dev, seed=0)
m = hb.DeviceMatrix(V)
for B in (64, 1024, 4096):
    Q = torch.as_tensor(bench.gen_queries(B, d, "float16", seed=5)).to(dev)
    idx = torch.empty((B, 10), dtype=torch.int64, device=dev); sc = torch.empty((B, 10), dtype=torch.float64, device=dev)
    cnt = torch.empty(B, dtype=torch.int64, device=dev); fl = torch.zeros(B, dtype=torch.int32, device=dev)
    for it in range(2):
        m.query_device(Q, 10, "cosine_similarity", 0.0, idx, sc, cnt, fl)
    torch.cuda.synchronize()
    m.profile_enable(16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    e0.record()
    for it in range(iters):
        m.query_device(Q, 10, "cosine_similarity", 0.0, idx, sc, cnt, fl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nl, kms = m.profile_read(); m.profile_enable(0)
    flops = 2.0 * n * d * B
    print(f"B={B}: {ms:.3f} ms/batch -> {B/ms*1e3:.0f} q/s; contraction {kms/nl:.3f} ms = {flops/(kms/nl*1e-3)/1e12:.1f} TFLOP/s; flags {set(fl.cpu().tolist())}; GB/s matrix {n*d*2/(kms/nl*1e-3)/1e9:.0f}")
    # correctness vs sweep for a few queries
    m.set_path(2)
    i0, s0, c0, f0 = m.query(Q[:3].cpu().numpy(), 10, "cosine_similarity")
    m.set_path(0)
    print("   first 3 queries equal to sweep:", np.array_equal(i0, idx[:3].cpu().numpy()), np.array_equal(s0, sc[:3].cpu().numpy()))
m.close()
