"""Small driver for ncu captures of the row-wise kernels: 1M x 768 fp16 and 1M x 384 fp32."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "local-hyperdb_b200")):
    sys.path.insert(0, p)
import torch
import bench
import hyperdb_b200 as hb
from hyperdb_b200 import _native as N

dev = torch.device("cuda", 0)
for n, d, dtype in ((1_000_000, 768, "float16"), (1_000_000, 384, "float32")):
    rows = bench.gen_rows_torch(0, n, d, dtype, dev, seed=0)
    m = hb.DeviceMatrix(rows)
    q = torch.as_tensor(bench.gen_queries(1, d, dtype)[0]).to(dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    got = C.c_int()
    for metric in ("cosine_similarity", "euclidean_metric", "pearson_correlation"):
        N.check(N.lib().hdb_scores(m._h, N.METRIC_IDS[metric], C.c_void_p(q.data_ptr()), {"float16": 0, "float32": 1}[dtype], N.HDB_DEVICE,
                                   C.c_void_p(out.data_ptr()), N.HDB_DEVICE, C.byref(got)))
    torch.cuda.synchronize()
    m.close()
