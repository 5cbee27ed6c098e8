"""Fixed cost of a sweep launch: the sweep alone (hdb_time_last_query what=1) on shards so small that every warp sees 1, 2, 4 ... windows."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "local-hyperdb_b200")]
import hyperdb_b200 as hb
W = 296 * 8 * 32          # rows of one window per warp
for metric, dtype, d in (("hamming_distance", torch.float32, 1024), ("cosine_similarity", torch.float16, 768), ("cosine_similarity", torch.float32, 384)):
    for mult in (1, 2, 4, 8, 16):
        n = W * mult
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        V = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32).to(dtype)
        m = hb.DeviceMatrix(V)
        q = torch.randn(d, generator=g, device="cuda", dtype=torch.float32).to(dtype).cpu().numpy()
        m.query(q, 10, metric)
        t = min(m.time_last_query(1, 300) for _ in range(3))
        b = n * ((d + 127) // 128) * 16 if metric == "hamming_distance" else n * d * V.element_size()
        print(f"{metric:18s} {str(dtype):14s} d={d:5d} windows/warp={mult:3d} n={n:8d} sweep {t*1e3:8.2f} us  stream-only {b/7.0e6:7.2f} us", flush=True)
        m.close(); del V
