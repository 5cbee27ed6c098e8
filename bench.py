#!/usr/bin/env python
"""bench.py -- queries/sec of the brute-force ranking hot path on synthetic unit-norm vectors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic queries (B queries against the
whole stored matrix, top-k out).  Default workload `c3_cosine_b1`: BASELINE.json config C3, the one
north_star's roofline target is quoted on (10M x 768 fp16, cosine top-10, single query).
  value     whole-job queries/s with matrix, queries and results resident in HBM (CUDA events)
  e2e       the same through the public host API: every step's query starts in pinned HOST memory and its
            top-k lands in HOST memory.  `e2e.value` keeps up to 3 steps in flight through
            ShardedMatrix.submit / collect (hdb_query_submit / hdb_query_collect); `e2e.sync_value` is one
            blocking DeviceMatrix.query / ShardedMatrix.query call after the other
  roofline  the dominant kernel: algorithmic bytes (or flops) per launch / its mean duration, measured
            live with CUDA event pairs around every launch inside the timed region.  With a row mask the
            bytes are those of the KEPT rows (the full N*D figure is given beside it)
  cpu_baseline   the NumPy port of the reference (oracle/reference_port.py) on this box's host cores, on a
            bounded row subsample (median of 3), scaled linearly to the full row count
  parity_check   outside the timed region: two planted queries (a stored row must come back first) and one
            random query checked against an independent chunked torch fp32 scoring of every shard
  extra     short passes of the other BASELINE.json configurations (batched tensor-core path f16 / tf32, bit-packed
            hamming, the multi-query sweep, the 100M-row masked + decayed config C4 plain and clustered), each with its own
            e2e and roofline
N > 1 (torchrun): the matrix is row-sharded (strong scaling), one exchange of candidates per step.
`--impl reference` times the reference's CPU path (the NumPy port; the reference is pure Python and
/root/reference does not travel to the GPU box) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "local-hyperdb_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: rows, dim, dtype, metric, top_k, batch, decay/mask
    "c3_cosine_b1": dict(n=10_000_000, d=768, dtype="float16", metric="cosine_similarity", k=10, b=1),
    "c3_dot_b1": dict(n=10_000_000, d=768, dtype="float16", metric="dot_product", k=10, b=1),
    "c3_pearson_b1": dict(n=10_000_000, d=768, dtype="float16", metric="pearson_correlation", k=10, b=1),
    "c3_cosine_b64": dict(n=10_000_000, d=768, dtype="float16", metric="cosine_similarity", k=10, b=64),
    "c3_cosine_b4096": dict(n=10_000_000, d=768, dtype="float16", metric="cosine_similarity", k=10, b=4096),
    "c3_dot_b4096": dict(n=10_000_000, d=768, dtype="float16", metric="dot_product", k=10, b=4096),
    "c2_cosine_b1024": dict(n=1_000_000, d=384, dtype="float32", metric="cosine_similarity", k=10, b=1024),
    "c2_cosine_b1": dict(n=1_000_000, d=384, dtype="float32", metric="cosine_similarity", k=10, b=1),
    "c5_euclid_b1": dict(n=5_000_000, d=1024, dtype="float32", metric="euclidean_metric", k=10, b=1),
    "c5_euclid_b1024": dict(n=5_000_000, d=1024, dtype="float32", metric="euclidean_metric", k=10, b=1024),
    "c5_manhattan_b1": dict(n=5_000_000, d=1024, dtype="float32", metric="manhattan_distance", k=10, b=1),
    "c5_hamming_b1": dict(n=5_000_000, d=1024, dtype="float32", metric="hamming_distance", k=10, b=1),
    # small batches: ONE read of the matrix serves up to 8 queries (multi-query sweep)
    "c5_manhattan_b8": dict(n=5_000_000, d=1024, dtype="float32", metric="manhattan_distance", k=10, b=8),
    "c5_euclid_b8": dict(n=5_000_000, d=1024, dtype="float32", metric="euclidean_metric", k=10, b=8),
    "c5_hamming_b8": dict(n=5_000_000, d=1024, dtype="float32", metric="hamming_distance", k=10, b=8),
    "c3_cosine_b8": dict(n=10_000_000, d=768, dtype="float16", metric="cosine_similarity", k=10, b=8),
    "c3_pearson_b8": dict(n=10_000_000, d=768, dtype="float16", metric="pearson_correlation", k=10, b=8),
    # pearson batches on the tensor cores: V . (q - mean q), 1 / (std_v d) in the epilogue, 1 / std_q in the certify step
    "c3_pearson_b1024": dict(n=10_000_000, d=768, dtype="float16", metric="pearson_correlation", k=10, b=1024),
    "c4_decay_mask_k100": dict(n=100_000_000, d=384, dtype="float16", metric="cosine_similarity", k=100, b=1,
                               decay=True, mask=True),
    # the same store CLUSTERED by the filtered metadata key at ingest (hdb_matrix_set_row_order): same ids, ties and answers,
    # but the kept rows form contiguous runs
    "c4_decay_mask_k100_clustered": dict(n=100_000_000, d=384, dtype="float16", metric="cosine_similarity", k=100, b=1,
                                         decay=True, mask=True, cluster=True),
}
# the short passes attached to the default line as "extra" (same matrix as the headline first, then config C5's)
EXTRAS = ["c3_cosine_b8", "c3_cosine_b64", "c3_cosine_b4096", "c3_pearson_b1024", "c5_hamming_b1", "c5_manhattan_b8", "c2_cosine_b1024",
          "c4_decay_mask_k100", "c4_decay_mask_k100_clustered"]
CHUNK = 262_144          # rows per generator chunk: chunk c of the GLOBAL matrix is seeded with (seed, c)
ITEM = {"float16": 2, "float32": 4, "float64": 8}
# metadata of the synthetic documents (config C4): one category per row, the filter keeps 4 of the 8 categories
MASK_CATEGORIES, MASK_KEPT = 8, (1, 3, 5, 6)


def matrix_bytes(w, rows):
    if w["metric"] in ("hamming_distance", "jaccard_similarity"):
        return rows * ((w["d"] + 127) // 128) * 16          # bit-packed rows, 16-byte granules
    return rows * w["d"] * ITEM[w["dtype"]]


# ------------------------------------------------------------------------------------------------
# synthetic data: unit-norm Gaussian rows, fixed seeds, any shard reproducible
# ------------------------------------------------------------------------------------------------
def gen_rows_torch(lo, hi, d, dtype, device, seed=0, dest=None):
    """Rows lo..hi of the global synthetic matrix.  dest (int64 CUDA tensor [hi - lo]): row lo + i is stored at position
    dest[i] (a clustered physical layout) instead of position i."""
    import torch
    tdt = getattr(torch, dtype)
    out = torch.empty((hi - lo, d), dtype=tdt, device=device)
    c0, c1 = lo // CHUNK, (hi + CHUNK - 1) // CHUNK
    for c in range(c0, c1):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1_000_003 + c)
        x = torch.randn((CHUNK, d), generator=g, device=device, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True).clamp_min(1e-30)
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        if dest is None:
            out[a - lo:b - lo] = x[a - c * CHUNK:b - c * CHUNK].to(tdt)
        else:
            out[dest[a - lo:b - lo]] = x[a - c * CHUNK:b - c * CHUNK].to(tdt)
        del x
    return out


def gen_queries(b, d, dtype, seed=1):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((b, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(dtype)


def gen_rows_numpy(n, d, dtype, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, d), dtype=np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(dtype)


def gen_category_torch(lo, hi, device):
    """category of document i = splitmix64(i) % 8 (int64 arithmetic wraps): reproducible per global id, no visible period"""
    import torch
    z = torch.arange(lo, hi, device=device, dtype=torch.int64) + 0x1234567
    z = (z ^ ((z >> 30) & 0x3FFFFFFFF)) * -4658895280553007687          # 0xbf58476d1ce4e5b9
    z = (z ^ ((z >> 27) & 0x1FFFFFFFFF)) * -7723592293110705685         # 0x94d049bb133111eb
    z = z ^ ((z >> 31) & 0x1FFFFFFFF)
    return (z & 0x7FFFFFFF) % MASK_CATEGORIES


def gen_keep_torch(lo, hi, device):
    """The metadata predicate of config C4 on synthetic documents: document i carries category
    hash(i) % 8 (a device-resident int8 column, reproducible per global row id) and the filter keeps the
    documents whose category is in MASK_KEPT (`filters=[('metadata', {'category': [...]})]` in the reference,
    hyperdb/hyperdb.py:1218-1257).  Returns the bool keep column of rows [lo, hi)."""
    import torch
    cat = gen_category_torch(lo, hi, device)
    keep = torch.zeros(hi - lo, dtype=torch.bool, device=device)
    for c in MASK_KEPT:
        keep |= cat == c
    return keep


def pack_keep_bits(keep):
    import torch
    n = keep.numel()
    pad = (-n) % 32
    if pad:
        keep = torch.cat([keep, torch.zeros(pad, dtype=torch.bool, device=keep.device)])
    w = keep.view(-1, 32).to(torch.int64) << torch.arange(32, device=keep.device, dtype=torch.int64)
    bits = w.sum(dim=1)
    bits = torch.where(bits >= 2**31, bits - 2**32, bits)
    return bits.to(torch.int32).contiguous()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` captures (profiles/r02_traffic.json, else r01's); None if this workload was not captured."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name))).get(workload)
            if t is not None:
                return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
        except Exception:
            pass
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


_TF32_PEAK = None


def tf32_peak_tflops(dev):
    """No tf32 figure is driver-written: measure torch.matmul (cuBLAS, allow_tf32) 8192^3 back to back for ~1 s,
    the same recipe MEASURED_PEAKS.json states for its sustained bf16 figure."""
    global _TF32_PEAK
    if _TF32_PEAK is not None:
        return _TF32_PEAK
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device=dev)
        b = torch.randn(8192, 8192, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(400):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        _TF32_PEAK = 400 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return _TF32_PEAK


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_time_per_query(w, sample_rows, repeats=1):
    """-> (median seconds of `repeats` full ranking calls, median seconds of the bare metric kernel -- np.dot for the
    dot/cosine family -- on the same sample)."""
    from oracle import reference_port as P
    v = gen_rows_numpy(sample_rows, w["d"], w["dtype"], seed=0)
    qs = gen_queries(repeats, w["d"], w["dtype"])
    ts = None
    if w.get("decay"):
        ts = 1.7e9 + np.random.default_rng(2).uniform(0, 3600, sample_rows)
    full, kern = [], []
    for r in range(repeats):
        t0 = time.perf_counter()
        P.rank(v, qs[r], w["k"], w["metric"], ts, 0.3 if ts is not None else 0, canonical=False)
        full.append(time.perf_counter() - t0)
    if w["metric"] in ("cosine_similarity", "dot_product", "pearson_correlation"):
        for r in range(repeats):
            t0 = time.perf_counter()
            np.dot(v, qs[r])
            kern.append(time.perf_counter() - t0)
    return float(np.median(full)), (float(np.median(kern)) if kern else None)


def cpu_sample_rows(w, repeats):
    # sized for roughly 10-30 s of host work in total on the survey machine
    per_row_us = {"float16": 60.0, "float32": 4.5, "float64": 6.0}[w["dtype"]] * w["d"] / 768
    if w["metric"] == "hamming_distance":
        per_row_us = 60.0 * w["d"] / 1024
    if w["metric"] == "dot_product":
        per_row_us *= 0.2
    return int(max(10_000, min(w["n"], 15e6 / per_row_us / repeats)))


def run_reference_arm(args, w):
    """`--impl reference`: the reference's CPU path (NumPy port) on this box's host cores, EXACTLY --steps timed steps
    after --warmup untimed ones; each step ranks one query against a row sample sized so the whole run takes ~2 min,
    and the per-step time is scaled linearly to the full row count (rows are independent in the reference)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per_row_s = cpu_time_per_query(w, 20_000)[0] / 20_000                 # calibration, untimed
    sample = int(max(2_000, min(w["n"], 120.0 / (steps + warm) / per_row_s)))
    v = gen_rows_numpy(sample, w["d"], w["dtype"], seed=0)
    qs = gen_queries(steps + warm, w["d"], w["dtype"])
    ts = 1.7e9 + np.random.default_rng(2).uniform(0, 3600, sample) if w.get("decay") else None
    from oracle import reference_port as P
    for i in range(warm):
        P.rank(v, qs[i], w["k"], w["metric"], ts, 0.3 if ts is not None else 0, canonical=False)
    t0 = time.perf_counter()
    for i in range(steps):
        P.rank(v, qs[warm + i], w["k"], w["metric"], ts, 0.3 if ts is not None else 0, canonical=False)
    t_step = (time.perf_counter() - t0) / steps
    t_full = t_step * (w["n"] / sample) * w["b"]                       # a batch is B independent calls in the reference
    qps = w["b"] / t_full
    cores = len(os.sched_getaffinity(0))
    line = {
        "impl": "reference", "metric": "queries/sec @top-%d" % w["k"], "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
        "config": {"workload": args.workload, "rows": w["n"], "dim": w["d"], "metric": w["metric"], "top_k": w["k"], "batch": w["b"]},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"each step = 1 query on {sample} of {w['n']} rows ({t_step * 1e3:.1f} ms), scaled linearly by "
                                   f"{w['n'] / sample:.1f}x (and by the batch size); oracle/reference_port.rank = the reference's NumPy "
                                   f"calls (without its per-call NaN scan of the matrix: slightly faster than the reference), "
                                   f"NumPy {np.__version__}, all threads NumPy/OpenBLAS chooses to use"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# one resident matrix (a row shard per rank) and the workloads measured on it
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args, world, rank, dev):
        self.args, self.world, self.rank, self.dev = args, world, rank, dev
        self.shape = None
        self.rows = self.m = self.eng = self.sm = self.perm = self.keep = self.ts = None

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def close(self):
        import gc
        import torch
        if self.m is not None:
            torch.cuda.synchronize()
            if self.sm.xchg is not None:
                self.barrier()
                self.eng.enable_pipeline(False)
                self.eng.attach_exchange(None)
                self.sm.xchg.close()
                self.sm.xchg = None
            self.m.close()
        self.rows = self.m = self.eng = self.sm = self.perm = self.keep = self.ts = None
        self.shape = None
        gc.collect()
        torch.cuda.empty_cache()

    def prepare(self, w, max_b, max_k):
        """Generate (or keep) the matrix of workload `w` and configure mask / decay / exchange for it."""
        import torch
        import hyperdb_b200 as hb
        from hyperdb_b200.sharded import CudaEngine, ShardedMatrix, shard_bounds
        shape = (w["n"], w["d"], w["dtype"], bool(w.get("mask")), bool(w.get("decay")), bool(w.get("cluster")))
        if shape != self.shape:
            self.close()
            self.lo, self.hi = shard_bounds(w["n"], self.world, self.rank)
            self.perm = None
            if w.get("cluster"):
                # ingest clusters the shard by the documents' category (stable): physical row p holds document lo + perm[p]
                cat = gen_category_torch(self.lo, self.hi, self.dev)
                self.perm = torch.sort(cat, stable=True).indices
                dest = torch.empty_like(self.perm)
                dest[self.perm] = torch.arange(self.hi - self.lo, device=self.dev, dtype=torch.int64)
                self.rows = gen_rows_torch(self.lo, self.hi, w["d"], w["dtype"], self.dev, seed=0, dest=dest)
                del cat, dest
            else:
                self.rows = gen_rows_torch(self.lo, self.hi, w["d"], w["dtype"], self.dev, seed=0)
            self.m = hb.DeviceMatrix(self.rows, row_offset=self.lo)
            if self.perm is not None:
                self.m.set_row_order(self.perm.to(torch.int32))
            self.keep = None
            if w.get("mask"):
                self.keep = gen_keep_torch(self.lo, self.hi, self.dev)
                if os.environ.get("HDB_BENCH_MASK") == "random":          # A/B only: round 1's mask (independent random bits)
                    g = torch.Generator(device=self.dev)
                    g.manual_seed(3_000_003 + self.rank)
                    self.keep = torch.rand(self.hi - self.lo, generator=g, device=self.dev) < 0.5
                if self.perm is not None:
                    self.keep = self.keep[self.perm]                  # per-row inputs travel in physical order
                self.m.set_mask(pack_keep_bits(self.keep))
            self.eng = CudaEngine(self.m)
            self.sm = ShardedMatrix(self.eng, w["n"])
            if self.world > 1 and self.args.exchange == "peer":
                # CUDA IPC between the ranks' processes; if the box forbids it, every rank fails here alike and the run
                # continues on the NCCL all-gather (reported in config.exchange)
                try:
                    self.sm.enable_peer_exchange(max_batch=max_b, max_k=max_k)
                except Exception as e:                                   # noqa: BLE001
                    print(f"bench.py: peer-memory exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
                    self.sm.xchg = None
            self.ts = None
            if w.get("decay"):
                g = torch.Generator(device=self.dev)
                g.manual_seed(2_000_003 + self.rank)
                self.ts = 1.7e9 + 3600.0 * torch.rand(self.hi - self.lo, generator=g, device=self.dev, dtype=torch.float64)
                if self.perm is not None:
                    self.ts = self.ts[self.perm]
                self.m.set_timestamps(self.ts)
                self.sm.refresh_decay()
            self.shape = shape
        if self.args.path:
            self.eng.set_path(self.args.path)
            self.eng.set_path = lambda mode: None              # keep the forced path for the whole run
        self.m.set_max_group(self.args.max_group)
        pipelined = not self.args.no_pipeline and w["b"] < 2
        self.eng.enable_pipeline(pipelined)      # certify/exchange/merge of query i overlap the sweep of query i+1
        if pipelined:
            self.m.set_sweep_overlap(not self.args.no_overlap)     # ... and the head of sweep i+1 fills the tail of sweep i
        return 0.3 if w.get("decay") else 0.0

    # ---- the measurement of one workload -------------------------------------------------------------------------
    def measure(self, name, w, steps, warmup, max_b, max_k, want_clocks=True, graph=False):
        import torch
        import torch.distributed as dist
        from hyperdb_b200 import _native as N
        args, world, rank, dev = self.args, self.world, self.rank, self.dev
        bias = self.prepare(w, max_b, max_k)
        m, eng, sm = self.m, self.eng, self.sm
        lo, hi = self.lo, self.hi
        b, k = w["b"], w["k"]
        # a different query (batch) every step; large batches cycle through a pool of 6 batches
        pool = (warmup + steps + 1) if b == 1 else min(warmup + steps + 1, 6)
        q_host = gen_queries(pool * b, w["d"], w["dtype"], seed=1 + b)
        q_dev = torch.as_tensor(q_host).to(dev)
        q_pin = torch.as_tensor(q_host).pin_memory()

        def qslice(t, i):
            j = (i % pool) * b
            return t[j:j + b]

        # ---- device-resident arm (value) ------------------------------------------------------------
        graphed = None
        if graph and world == 1 and b <= 64 and eng.post is None:
            from hyperdb_b200.sharded import GraphedQuery
            graphed = GraphedQuery(sm, qslice(q_dev, 0), k, w["metric"], bias)
        step = (lambda q: graphed.replay(q)) if graphed else (lambda q: sm.query_async(q, k, w["metric"], bias))
        outs, flag_log = [], []
        for i in range(warmup):
            outs.append(step(qslice(q_dev, i)))
        self.barrier()
        m.profile_enable(steps * b + 8)
        N.lib().hdb_launch_count(1)
        sampler = ClockSampler(dev.index)
        if rank == 0 and want_clocks:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for i in range(steps):
            o = step(qslice(q_dev, warmup + i))
            if graphed:
                flag_log.append(o[3].clone())            # the static outputs are overwritten by the next replay
            else:
                outs.append(o)
        sm.wait_results()                                # pipelined mode: the last certify/exchange/merge run on the post stream
        e1.record()
        self.barrier()
        launches = N.lib().hdb_launch_count(0)
        clocks = sampler.stop() if (rank == 0 and want_clocks) else None
        ms_total = e0.elapsed_time(e1)
        n_sweeps, sweep_ms = m.profile_read()
        if graphed:
            # Kernels replayed from a graph cannot be bracketed by event pairs and are not counted by the launch
            # counter: run K more un-graphed steps of the same workload for the per-launch duration of the dominant
            # kernel (roofline) and the launch count per step (the graph replays exactly these launches).
            N.lib().hdb_launch_count(1)
            for i in range(steps):
                sm.query_async(qslice(q_dev, warmup + i), k, w["metric"], bias)
            self.barrier()
            n_sweeps, sweep_ms = m.profile_read()
            launches = N.lib().hdb_launch_count(0)
        m.profile_enable(0)
        if graphed:
            uncertified = sum(int(bool((f & N.FLAG_UNCERTIFIED).any())) for f in flag_log)
        else:
            uncertified = sum(int(bool((o[3] & N.FLAG_UNCERTIFIED).any())) for o in outs[warmup:])
        tensor_path = any(int(f) & N.FLAG_TENSOR for o in outs[-1:] for f in o[3].flatten().tolist()) if outs else False
        del outs, flag_log
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = t.item() / steps
        value = b / (ms_step * 1e-3)

        # ---- end-to-end arm through the host API ------------------------------------------------------
        e2e_steps = max(3, min(steps, 20))
        q_np = q_pin.numpy()
        use_tickets = world == 1 or sm.xchg is not None
        # (a) one blocking call after the other: DeviceMatrix.query on one GPU, ShardedMatrix.query on several
        host_query = (lambda q: m.query(q, k, w["metric"], bias)) if world == 1 else (lambda q: sm.query(q, k, w["metric"], bias))
        for i in range(2):
            host_query(qslice(q_np, i))
        self.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            host_query(qslice(q_np, warmup + i))
        torch.cuda.synchronize()
        t_sync = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_sync, op=dist.ReduceOp.MAX)
        sync_qps = b * e2e_steps / t_sync.item()
        # (b) the asynchronous host API with up to 3 steps in flight (host query in, host top-k out, every step)
        e2e_qps, depth = sync_qps, 1
        if use_tickets:
            depth = 3
            warm = [sm.submit(qslice(q_np, i), k, w["metric"], bias) for i in range(4)]      # every ticket slot: its pinned block exists
            for t in warm:
                sm.collect(t)
            self.barrier()
            t0 = time.perf_counter()
            tickets = []
            for i in range(e2e_steps):
                tickets.append(sm.submit(qslice(q_np, warmup + i), k, w["metric"], bias))
                if len(tickets) == depth:
                    sm.collect(tickets.pop(0))
            while tickets:
                sm.collect(tickets.pop(0))
            t_pipe = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_pipe, op=dist.ReduceOp.MAX)
            e2e_qps = b * e2e_steps / t_pipe.item()

        res = None
        if rank == 0:
            peaks, peak_src = measured_peaks()
            n_kept = m.n_kept
            full_bytes = matrix_bytes(w, hi - lo)
            shard_bytes = matrix_bytes(w, n_kept) if w.get("mask") else full_bytes
            side_bytes = 0
            if w.get("mask"):
                side_bytes += (hi - lo) // 8                                      # the keep bits of every row
            if w["metric"] in ("cosine_similarity", "pearson_correlation"):
                side_bytes += n_kept * 4 * (2 if w["metric"] == "pearson_correlation" else 1)
            if w.get("decay"):
                side_bytes += n_kept * 8
            sweep_avg_ms = sweep_ms / max(1, n_sweeps)
            overlapped = eng.post is not None and not args.no_overlap and b < 2
            if overlapped and n_sweeps:
                # Overlapping sweeps: an event pair around a launch also spans the time the kernel waited for the previous
                # query's CTAs to leave the SMs, so the per-launch average is taken as timed region / launches (an upper
                # bound of the kernel's own duration: the region also holds the small kernels), whichever is smaller.
                sweep_avg_ms = min(sweep_avg_ms, ms_total / n_sweeps)
            achieved = shard_bytes / (sweep_avg_ms * 1e-3) / 1e9 if n_sweeps else None
            # tensor-bound only when the contraction, not the row stream, limits the step: flops per matrix byte above
            # the machine balance (peak flop/s / peak byte/s)
            shard_flops = 2.0 * (hi - lo) * w["d"] * b
            kernel = ("batched_tc_kernel (sample + select passes)" if tensor_path else
                      ("sweep_hamming_kernel" if w["metric"] in ("hamming_distance", "jaccard_similarity") else "sweep_kernel") +
                      (f" ({min(b, 8)} queries per pass)" if b > 1 else ""))
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": (achieved / peaks["hbm_gbs"]) if achieved else None,
                    "traffic": ncu_traffic(name) if world == 1 else None, "kernel": kernel,
                    "launches_timed": n_sweeps, "avg_launch_ms": sweep_avg_ms,
                    "algorithmic_bytes_per_launch": shard_bytes, "peak_source": peak_src,
                    "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None}
            if w.get("mask"):
                roof["bytes_counted"] = "kept rows only (n_kept * d * sizeof); side columns (keep bits, inverse norms, decay) excluded"
                roof["n_kept"] = int(n_kept)
                roof["full_matrix_bytes"] = full_bytes
                roof["side_column_bytes"] = int(side_bytes)
                roof["frac_if_dropped_rows_counted"] = full_bytes / (sweep_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if n_sweeps else None
            if tensor_path and n_sweeps and b >= 256:
                tf = shard_flops / (sweep_avg_ms * 1e-3) / 1e12
                if w["dtype"] == "float32":
                    peak_tf = tf32_peak_tflops(dev)
                    src = "measured here: torch.matmul fp32 with allow_tf32, 8192^3 x 400 back to back (no tf32 figure in MEASURED_PEAKS.json)"
                    kind = "kind::tf32"
                else:
                    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
                    src = peak_src + " bf16_tflops_sustained (kernel timed inside a long step)"
                    kind = "kind::f16"
                roof = {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                        "traffic": ncu_traffic(name) if world == 1 else None, "kernel": kernel + " " + kind,
                        "launches_timed": n_sweeps, "avg_launch_ms": sweep_avg_ms, "algorithmic_flops_per_launch": shard_flops,
                        "peak_source": src}
                if w["dtype"] != "float32":
                    roof["frac_of_burst"] = tf / peaks.get("bf16_tflops", 1667.5)
                    roof["frac_of_nominal_2250"] = tf / 2250.0
            res = {
                "workload": name, "value": value, "unit": "queries/s", "ms_per_step": ms_step, "steps": steps, "warmup": warmup,
                "uncertified_steps": uncertified, "gpu_launches": int(launches), "cuda_graph": bool(graphed),
                "pipelined": eng.post is not None, "sweep_overlap": bool(overlapped), "clocks": clocks,
                "shard_gb": full_bytes / 1e9,
                "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(b * w["d"] * ITEM[w["dtype"]]),
                        "d2h_bytes_per_step": int(b * k * 16 + b * 8 + world * b * 4), "steps": e2e_steps,
                        "api": (f"ShardedMatrix.submit/collect (hdb_query_submit/hdb_query_collect), {depth} steps in flight, "
                                "uncertified steps repaired inside collect") if use_tickets else "ShardedMatrix.query (NCCL all-gather)",
                        "sync_value": sync_qps,
                        "sync_api": "DeviceMatrix.query" if world == 1 else "ShardedMatrix.query"},
                "roofline": roof,
            }
        return res

    # ---- answers checked outside the timed region ------------------------------------------------------------------
    def parity_check(self, w):
        """Two planted queries (stored rows of two different shards: they must come back first) and one random query
        against an independent scoring of every shard with plain torch fp32 (chunked matmul + topk), merged over the
        ranks by (score desc, id asc).  Only for the dot / cosine family; the sharded host API is the path checked."""
        import torch
        import torch.distributed as dist
        from hyperdb_b200.sharded import shard_bounds
        if w["metric"] not in ("cosine_similarity", "dot_product"):
            return None
        world, rank, dev = self.world, self.rank, self.dev
        bias = 0.3 if w.get("decay") else 0.0
        ts_max = None
        if w.get("decay"):
            # the decay reference is the newest timestamp over the KEPT rows of ALL shards (hyperdb/ranking_algorithm.py:183)
            kept_ts = self.ts if self.keep is None else self.ts[self.keep]
            ts_max = kept_ts.max() if kept_ts.numel() else torch.tensor(float("-inf"), dtype=torch.float64, device=self.dev)
            if world > 1:
                dist.all_reduce(ts_max, op=dist.ReduceOp.MAX)
        sm, rows, lo, hi = self.sm, self.rows, self.lo, self.hi
        n, d, k = w["n"], w["d"], w["k"]
        plants = [n // 3, (2 * n) // 3 + 1]
        if self.keep is not None:                        # a planted row must be a kept one: move to the next kept row (rank-local)
            plants = plants[:0]
        qs = []
        for pid in plants:
            row = torch.zeros(d, dtype=rows.dtype, device=dev)
            if lo <= pid < hi:
                row = rows[pid - lo].clone()
            if world > 1:
                src = [r for r in range(world) if shard_bounds(n, world, r)[0] <= pid < shard_bounds(n, world, r)[1]][0]
                dist.broadcast(row, src=src)
            qs.append(row.cpu().numpy())
        qs.append(gen_queries(1, d, w["dtype"], seed=777)[0])
        out = {"queries": len(qs), "planted_first": True, "indices_equal": True, "max_rel_score_err": 0.0, "tolerance_ties": 0}
        for qi, q in enumerate(qs):
            idx, sc, cnt = sm.query(q, k, w["metric"], bias)[:3]
            idx, sc = idx[0], sc[0]
            # independent check
            qf = torch.as_tensor(q.astype(np.float32)).to(dev)
            qn = qf / qf.norm().clamp_min(1e-30) if w["metric"] == "cosine_similarity" else qf
            best_s = torch.empty(0, dtype=torch.float64, device=dev)
            best_i = torch.empty(0, dtype=torch.int64, device=dev)
            step = 1 << 20
            for a in range(0, hi - lo, step):
                v = rows[a:a + step].float()
                s = v @ qn
                if w["metric"] == "cosine_similarity":
                    nv = v.norm(dim=1)
                    s = s / torch.where(nv == 0, torch.ones_like(nv), nv)
                s = s.double()
                if ts_max is not None:
                    s = s + bias * torch.exp(self.ts[a:a + step] - ts_max)
                if self.keep is not None:
                    s = torch.where(self.keep[a:a + step], s, torch.full_like(s, float("-inf")))
                kk = min(2 * k, s.numel())
                ts_, ti_ = torch.topk(s, kk)
                ti_ = ti_ + a
                if self.perm is not None:
                    ti_ = self.perm[ti_]                          # clustered storage: physical row -> the document's local index
                best_s = torch.cat([best_s, ts_])
                best_i = torch.cat([best_i, ti_ + lo])
                del v, s
            kk = min(2 * k, best_s.numel())
            ts_, sel = torch.topk(best_s, kk)
            cand = torch.stack([ts_.double(), best_i[sel].double()])
            if world > 1:
                allc = [torch.empty_like(cand) for _ in range(world)]
                dist.all_gather(allc, cand)
                cand = torch.cat(allc, dim=1)
            cs, ci = cand[0].cpu().numpy(), cand[1].cpu().numpy().astype(np.int64)
            order = np.lexsort((ci, -cs))[:k]
            want_i, want_s = ci[order], cs[order]
            if qi < len(plants) and idx[0] != plants[qi]:
                out["planted_first"] = False
            rel = np.abs(sc[:len(want_s)] - want_s) / np.maximum(np.abs(want_s), 1e-30)
            out["max_rel_score_err"] = max(out["max_rel_score_err"], float(rel.max()) if len(rel) else 0.0)
            if not np.array_equal(idx, want_i):
                # fp32 torch scores vs the reference's arithmetic: a swap of two rows closer than the tolerance is a tie
                tol = 1e-3 if w["dtype"] == "float16" else 1e-5
                if set(idx.tolist()) == set(want_i.tolist()) or float(rel.max()) <= tol:
                    out["tolerance_ties"] += 1
                else:
                    out["indices_equal"] = False
        tol = 1e-3 if w["dtype"] == "float16" else 1e-5
        out["ok"] = bool(out["planted_first"] and out["indices_equal"] and out["max_rel_score_err"] <= tol)
        out["against"] = ("torch fp32 chunked matmul (+ float64 decay, keep mask, row order) + topk per shard, merged by (score desc, id asc); "
                          "planted rows %s" % plants)
        return out


# ------------------------------------------------------------------------------------------------
def hyperdb_query_extra(dev, steps=200, warmup=5):
    """End to end through the REFERENCE-FACING call, `HyperDB.query(query_vector, top_k=10)` (hyperdb/hyperdb.py:1584 -> the
    brute-force branch :1429-1582) of the drop-in shim, on config C2 (1M x 384 fp32, one document dict per row): a host query
    vector in, a list of (document, score, index) out, every step; query cache off (every query is new anyway).  Second figure:
    the same with a metadata predicate (`filters=[("metadata", {"category": 3})]`, 1 of 8 categories: the mask is compiled from the
    documents once and cached, hyperdb/hyperdb.py:1218-1257).  One GPU only; the shim's INFO lines go to a buffer, not to stdout."""
    import contextlib
    import io
    from hyperdb_b200.hyperdb import HyperDB
    w = WORKLOADS["c2_cosine_b1"]
    n, d = w["n"], w["d"]
    out = {"workload": "hyperdb_query_c2", "unit": "queries/s", "steps": steps, "warmup": warmup,
           "api": "hyperdb_b200.hyperdb.HyperDB.query(query_vector, top_k=10, metric='cosine_similarity'): host vector in, "
                  "[(document, score, index)] out; 1M documents, query cache off"}
    with contextlib.redirect_stdout(io.StringIO()):
        rows = gen_rows_torch(0, n, d, w["dtype"], dev, seed=0).cpu().numpy()
        cat = gen_category_torch(0, n, dev).cpu().numpy()
        docs = [{"id": i, "category": int(c)} for i, c in enumerate(cat)]
        db = HyperDB(documents=docs, vectors=rows, metadata_keys=["category"], fp_precision=w["dtype"], cache_size=0,
                     device=dev.index, sharded=False)
        try:
            qs = gen_queries(2 * (steps + warmup), d, w["dtype"], seed=99)
            hit = db.query(rows[n // 8], top_k=1)                  # a stored row as the query must come back first
            out["planted_first"] = bool(hit and hit[0][0]["id"] == n // 8)
            for name, kw, off in (("value", {}, 0), ("filtered_value", {"filters": [("metadata", {"category": 3})]}, steps + warmup)):
                for i in range(warmup):
                    db.query(qs[off + i], top_k=10, **kw)
                t0 = time.perf_counter()
                for i in range(steps):
                    res = db.query(qs[off + warmup + i], top_k=10, **kw)
                out[name] = steps / (time.perf_counter() - t0)
                assert len(res) == 10 and (not kw or all(doc["category"] == 3 for doc, _s, _i in res))
            out["filter"] = "metadata category == 3 (%d of %d documents kept)" % (int((cat == 3).sum()), n)
        finally:
            db.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3_cosine_b1", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debugging only; invalidates the line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="keep every kernel of a query on one stream")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU candidate exchange: NVLink peer-memory kernels (default) or the NCCL all-gather")
    ap.add_argument("--no-overlap", action="store_true", help="pipelined mode: do not let consecutive sweeps overlap")
    ap.add_argument("--graph", action="store_true", help="capture the step in a CUDA graph (single GPU only)")
    ap.add_argument("--path", type=int, default=0, help="A/B: hdb_matrix_set_path mode (2 = streaming sweep only, no tensor-core path)")
    ap.add_argument("--max-group", type=int, default=0, help="A/B: cap on the queries per sweep pass (1 = one pass per query)")
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "all"],
                    help="short passes of the other configurations attached as `extra` (auto: with the default workload)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.rows:
        w["n"] = args.rows
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (hyperdb_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    extras = []
    if args.extras == "all" or (args.extras == "auto" and args.workload == "c3_cosine_b1" and not args.rows):
        extras = list(EXTRAS)
    same = [e for e in extras if (WORKLOADS[e]["n"], WORKLOADS[e]["d"], WORKLOADS[e]["dtype"]) == (w["n"], w["d"], w["dtype"])]
    max_b = max([w["b"]] + [WORKLOADS[e]["b"] for e in same])
    max_k = max([w["k"]] + [WORKLOADS[e]["k"] for e in same])

    bench = Bench(args, world, rank, dev)
    head = bench.measure(args.workload, w, args.steps, args.warmup, max_b, max_k, graph=args.graph)
    parity = bench.parity_check(w)
    extra_lines = []
    for e in extras:
        we = dict(WORKLOADS[e])
        b_same = [x for x in extras if (WORKLOADS[x]["n"], WORKLOADS[x]["d"], WORKLOADS[x]["dtype"]) == (we["n"], we["d"], we["dtype"])]
        mb = max([w["b"]] + [WORKLOADS[x]["b"] for x in b_same])
        mk = max([w["k"]] + [WORKLOADS[x]["k"] for x in b_same])
        try:
            r = bench.measure(e, we, 5 if we["b"] > 1 else 20, 3, mb, mk, want_clocks=False)
        except Exception as ex:                                   # noqa: BLE001
            r = {"workload": e, "error": str(ex)[:300]} if rank == 0 else None
        if we.get("mask") or we.get("decay"):
            try:
                pc = bench.parity_check(we)                      # the masked / decayed / clustered configs carry their own check
            except Exception as ex:                               # noqa: BLE001
                pc = {"ok": False, "error": str(ex)[:200]}
            if r is not None and pc is not None:
                r["parity_check"] = pc
        if r is not None:
            r.pop("clocks", None)
            extra_lines.append(r)
    if extras and world == 1:
        # the reference-facing API itself (one GPU): HyperDB.query on config C2, plain and with a metadata predicate
        try:
            extra_lines.append(hyperdb_query_extra(dev))
        except Exception as ex:                                   # noqa: BLE001
            extra_lines.append({"workload": "hyperdb_query_c2", "error": str(ex)[:300]})

    if rank == 0:
        line = {
            "metric": "queries/sec @top-%d" % w["k"], "value": head["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"float16": "f16", "float32": "f32", "float64": "f64"}[w["dtype"]] + " storage, f32 accumulate"
            if w["dtype"] != "float64" else "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "rows": w["n"], "dim": w["d"], "metric": w["metric"], "top_k": w["k"], "batch": w["b"],
                       "sharding": f"rows/{world}", "l2": "inputs larger than L2 (shard %.2f GB per GPU, a new query every step)"
                       % head["shard_gb"], "uncertified_steps": head["uncertified_steps"], "cuda_graph": head["cuda_graph"],
                       "exchange": ("peer-memory" if bench.sm is not None and bench.sm.xchg is not None else "nccl all-gather")
                       if world > 1 else None,
                       "pipelined": head["pipelined"], "sweep_overlap": head["sweep_overlap"]},
            "clocks": head["clocks"],
            "e2e": head["e2e"],
            "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"],
        }
        if parity is not None:
            line["parity_check"] = parity
        if not args.no_cpu_baseline and world == 1:
            reps = 3
            sample = cpu_sample_rows(w, reps)
            tq, tk = cpu_time_per_query(w, sample, repeats=reps)
            line["cpu_baseline"] = {
                "value": 1.0 / (tq * w["n"] / sample), "unit": "queries/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                "sample": f"median of {reps} queries on {sample} of {w['n']} rows ({tq:.2f} s each), scaled linearly by {w['n'] / sample:.1f}x; "
                          f"oracle/reference_port.rank = the reference's NumPy calls without its per-call NaN scan of the matrix "
                          f"(hyperdb/ranking_algorithm.py:150), i.e. slightly faster than the reference itself; NumPy {np.__version__}"}
            if tk is not None:
                line["cpu_baseline"]["metric_kernel_only"] = {
                    "value": 1.0 / (tk * w["n"] / sample), "unit": "queries/s",
                    "what": f"np.dot(vectors, q) alone on the same sample ({tk:.2f} s), scaled the same way"}
        if extra_lines:
            line["extra"] = extra_lines
        print(json.dumps(line))
    bench.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
