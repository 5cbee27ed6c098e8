"""Fixed cost vs streaming rate of the sweep kernels: time the sweep alone (hdb_time_last_query what=1) at several shard sizes."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "local-hyperdb_b200")]
import hyperdb_b200 as hb

def run(metric, dtype, d, sizes, k=10):
    out = []
    for n in sizes:
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        V = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)
        V /= V.norm(dim=1, keepdim=True)
        V = V.to(dtype)
        m = hb.DeviceMatrix(V)
        q = torch.randn(d, generator=g, device="cuda", dtype=torch.float32).to(dtype).cpu().numpy()
        m.query(q, k, metric)
        ts = [m.time_last_query(1, 200) for _ in range(3)]
        bytes_ = n * ((d + 127) // 128) * 16 if metric in ("hamming_distance", "jaccard_similarity") else n * d * V.element_size()
        out.append((n, min(ts), bytes_))
        m.close(); del V
    # least squares t = t0 + bytes / bw
    A = np.array([[1.0, b] for _, _, b in out]); y = np.array([t for _, t, _ in out])
    (t0, inv_bw), *_ = np.linalg.lstsq(A, y, rcond=None)
    print(json.dumps({"metric": metric, "dtype": str(dtype), "d": d, "points": [(n, round(t * 1e3, 2), round(b / t / 1e6, 1)) for n, t, b in out],
                      "t0_us": round(t0 * 1e3, 2), "bw_GBs": round(1.0 / inv_bw / 1e6, 1)}), flush=True)

tag = os.environ.get("HDB_HAMMING_COOPERATIVE", "0")
print("cooperative =", tag)
run("hamming_distance", torch.float32, 1024, [625_000, 1_250_000, 2_500_000, 5_000_000])
run("cosine_similarity", torch.float16, 768, [250_000, 500_000, 1_000_000, 2_000_000, 4_000_000])
run("cosine_similarity", torch.float32, 384, [250_000, 500_000, 1_000_000, 2_000_000])
