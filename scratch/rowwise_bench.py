"""Times the row-wise passes (csrc/rowwise.cu) at the BASELINE shapes: ingest (row_stats), pearson_stats, hdb_scores
(full similarity vector), the exact full-vector path.  CUDA events on the default stream; prints one JSON line per item."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "local-hyperdb_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
import hyperdb_b200 as hb
from hyperdb_b200 import _native as N

dev = torch.device("cuda", 0)


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for name, n, d, dtype in (("C3", 10_000_000, 768, "float16"), ("C2", 1_000_000, 384, "float32"), ("C5", 5_000_000, 1024, "float32"),
                          ("smoke", 20_000, 768, "float16")):
    rows = bench.gen_rows_torch(0, n, d, dtype, dev, seed=0)
    nbytes = n * d * bench.ITEM[dtype]
    m = hb.DeviceMatrix(rows)
    ms = timed(lambda: N.check(N.lib().hdb_matrix_finalize(m._h)))
    print(json.dumps({"shape": name, "op": "hdb_matrix_finalize (row_stats + kept count)", "ms": ms, "GBps": nbytes / ms / 1e6}))
    q = bench.gen_queries(1, d, dtype)[0]
    out = torch.empty(n, dtype=torch.float64, device=dev)
    got = C.c_int()
    qd = torch.as_tensor(q).to(dev)
    qdt = {"float16": 0, "float32": 1}[dtype]
    for metric in ("cosine_similarity", "dot_product", "euclidean_metric", "manhattan_distance", "pearson_correlation", "hamming_distance"):
        fn = lambda: N.check(N.lib().hdb_scores(m._h, N.METRIC_IDS[metric], C.c_void_p(qd.data_ptr()), qdt, N.HDB_DEVICE,
                                                 C.c_void_p(out.data_ptr()), N.HDB_DEVICE, C.byref(got)))
        fn()
        ms = timed(fn)
        print(json.dumps({"shape": name, "op": f"hdb_scores {metric}", "ms": ms, "GBps": nbytes / ms / 1e6}))
    m.set_path(1)
    for metric in ("cosine_similarity", "euclidean_metric"):
        m.query(q, 10, metric)
        ms = timed(lambda: m.query(q, 10, metric))
        print(json.dumps({"shape": name, "op": f"exact path top-10 {metric} (scores + CUB sort, host out)", "ms": ms}))
    m.close()
    del rows, out
    torch.cuda.empty_cache()
