#!/usr/bin/env python
"""bench.py -- queries/sec of the brute-force ranking hot path on synthetic unit-norm vectors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic queries (B queries against the
whole stored matrix, top-k out).  Default workload `c3_cosine_b1`: BASELINE.json config C3, the one
north_star's roofline target is quoted on (10M x 768 fp16, cosine top-10, single query).
  value     whole-job queries/s with matrix, queries and results resident in HBM (CUDA events)
  e2e       the same through the public host API (DeviceMatrix.query / ShardedMatrix.query): the query
            starts in pinned HOST memory and the top-k lands in HOST memory every step
  roofline  the streaming sweep kernel: algorithmic bytes per launch (N*D*sizeof) / its mean duration,
            measured live with CUDA event pairs around every launch inside the timed region
  cpu_baseline   the NumPy port of the reference (oracle/reference_port.py) on this box's host cores,
            on a bounded row subsample, scaled linearly to the full row count
N > 1 (torchrun): the matrix is row-sharded (strong scaling), one all-gather of candidates per step.
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
    "c4_decay_mask_k100": dict(n=100_000_000, d=384, dtype="float16", metric="cosine_similarity", k=100, b=1,
                               decay=True, mask=True),
}
CHUNK = 262_144          # rows per generator chunk: chunk c of the GLOBAL matrix is seeded with (seed, c)
ITEM = {"float16": 2, "float32": 4, "float64": 8}


def algorithmic_bytes(w, rows):
    if w["metric"] == "hamming_distance":
        return rows * ((w["d"] + 127) // 128) * 16          # bit-packed rows, 16-byte granules
    return rows * w["d"] * ITEM[w["dtype"]]


# ------------------------------------------------------------------------------------------------
# synthetic data: unit-norm Gaussian rows, fixed seeds, any shard reproducible
# ------------------------------------------------------------------------------------------------
def gen_rows_torch(lo, hi, d, dtype, device, seed=0):
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
        out[a - lo:b - lo] = x[a - c * CHUNK:b - c * CHUNK].to(tdt)
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
    `ncu --set full` capture of this workload (profiles/r01_traffic.json); None if it was not captured."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        t = json.load(open(path)).get(workload)
        return None if t is None else float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_time_per_query(w, sample_rows, repeats=1):
    from oracle import reference_port as P
    v = gen_rows_numpy(sample_rows, w["d"], w["dtype"], seed=0)
    q = gen_queries(1, w["d"], w["dtype"])[0]
    ts = None
    if w.get("decay"):
        ts = 1.7e9 + np.random.default_rng(2).uniform(0, 3600, sample_rows)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        P.rank(v, q, w["k"], w["metric"], ts, 0.3 if ts is not None else 0, canonical=False)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_sample_rows(w):
    # sized for roughly 10-30 s of host work on the survey machine
    per_row_us = {"float16": 60.0, "float32": 4.5, "float64": 6.0}[w["dtype"]] * w["d"] / 768
    if w["metric"] == "hamming_distance":
        per_row_us = 60.0 * w["d"] / 1024
    if w["metric"] == "dot_product":
        per_row_us *= 0.2
    return int(max(10_000, min(w["n"], 12e6 / per_row_us)))


def run_reference_arm(args, w):
    """`--impl reference`: the reference's CPU path (NumPy port) on this box's host cores, EXACTLY --steps timed steps
    after --warmup untimed ones; each step ranks one query against a row sample sized so the whole run takes ~2 min,
    and the per-step time is scaled linearly to the full row count (rows are independent in the reference)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    per_row_s = cpu_time_per_query(w, 20_000) / 20_000                 # calibration, untimed
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
                                   f"calls, NumPy {np.__version__}, all threads NumPy/OpenBLAS chooses to use"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
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
    import hyperdb_b200 as hb
    from hyperdb_b200 import _native as N
    from hyperdb_b200.sharded import CudaEngine, ShardedMatrix, shard_bounds

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

    lo, hi = shard_bounds(w["n"], world, rank)
    rows = gen_rows_torch(lo, hi, w["d"], w["dtype"], dev, seed=0)
    m = hb.DeviceMatrix(rows, row_offset=lo)
    if w.get("mask"):
        g = torch.Generator(device=dev)
        g.manual_seed(3_000_003 + rank)
        keep_bits = torch.randint(-2**31, 2**31 - 1, ((hi - lo + 31) // 32,), generator=g, device=dev, dtype=torch.int32)
        m.set_mask(keep_bits)
    eng = CudaEngine(m)
    sm = ShardedMatrix(eng, w["n"])
    if world > 1 and args.exchange == "peer":
        # CUDA IPC between the ranks' processes; if the box forbids it, every rank fails here alike and the run
        # continues on the NCCL all-gather (reported in config.exchange)
        try:
            sm.enable_peer_exchange(max_batch=w["b"], max_k=w["k"])
        except Exception as e:                                   # noqa: BLE001
            print(f"bench.py: peer-memory exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
            sm.xchg = None
    if not args.no_pipeline and w["b"] < 2:
        eng.enable_pipeline()            # certify/exchange/merge of query i overlap the sweep of query i+1
        m.set_sweep_overlap(not args.no_overlap)     # ... and the head of sweep i+1 fills the tail of sweep i
    bias = 0.0
    if w.get("decay"):
        g = torch.Generator(device=dev)
        g.manual_seed(2_000_003 + rank)
        ts = 1.7e9 + 3600.0 * torch.rand(hi - lo, generator=g, device=dev, dtype=torch.float64)
        m.set_timestamps(ts)
        sm.refresh_decay()
        bias = 0.3
    b, k = w["b"], w["k"]
    # a different query (batch) every step; large batches cycle through a pool of 6 batches
    pool = (args.warmup + args.steps + 1) if b == 1 else min(args.warmup + args.steps + 1, 6)
    q_host = gen_queries(pool * b, w["d"], w["dtype"])
    q_dev = torch.as_tensor(q_host).to(dev)
    q_pin = torch.as_tensor(q_host).pin_memory()

    def qslice(t, i):
        j = (i % pool) * b
        return t[j:j + b]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm (value) ------------------------------------------------------------
    # one captured CUDA graph per step when the step is short enough for host launch overhead to matter
    graphed = None
    if args.graph and world == 1 and b <= 64 and eng.post is None:      # opt-in: NCCL collectives captured in a graph stalled at teardown here
        from hyperdb_b200.sharded import GraphedQuery
        graphed = GraphedQuery(sm, qslice(q_dev, 0), k, w["metric"], bias)
    step = (lambda q: graphed.replay(q)) if graphed else (lambda q: sm.query_async(q, k, w["metric"], bias))
    outs = []
    flag_log = []
    for i in range(args.warmup):
        outs.append(step(qslice(q_dev, i)))
    barrier()
    m.profile_enable(args.steps * b + 8)
    N.lib().hdb_launch_count(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        o = step(qslice(q_dev, args.warmup + i))
        if graphed:
            flag_log.append(o[3].clone())            # the static outputs are overwritten by the next replay
        else:
            outs.append(o)
    sm.wait_results()                                # pipelined mode: the last certify/exchange/merge run on the post stream
    e1.record()
    barrier()
    launches = N.lib().hdb_launch_count(0)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    n_sweeps, sweep_ms = m.profile_read()
    launches_per_step = launches / max(1, args.steps)
    if graphed:
        # Kernels replayed from a graph cannot be bracketed by event pairs and are not counted by the launch
        # counter: run K more un-graphed steps of the same workload for the per-launch duration of the dominant
        # kernel (roofline) and the launch count per step (the graph replays exactly these launches).
        N.lib().hdb_launch_count(1)
        for i in range(args.steps):
            sm.query_async(qslice(q_dev, args.warmup + i), k, w["metric"], bias)
        barrier()
        n_sweeps, sweep_ms = m.profile_read()
        launches_per_step = N.lib().hdb_launch_count(0) / max(1, args.steps)
        launches = int(round(launches_per_step * args.steps))
    m.profile_enable(0)
    if graphed:
        uncertified = sum(int(bool((f & N.FLAG_UNCERTIFIED).any())) for f in flag_log)
    else:
        uncertified = sum(int(bool((o[3] & N.FLAG_UNCERTIFIED).any())) for o in outs)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = b / (ms_step * 1e-3)

    # ---- end-to-end arm through the host API ------------------------------------------------------
    e2e_steps = max(3, min(args.steps, 20))
    # the call a user makes: DeviceMatrix.query on one GPU, ShardedMatrix.query on several
    host_query = (lambda q: m.query(q, k, w["metric"], bias)) if world == 1 else (lambda q: sm.query(q, k, w["metric"], bias))
    q_np = q_pin.numpy()
    for i in range(2):
        host_query(qslice(q_np, i))
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        host_query(qslice(q_np, args.warmup + i))
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_qps = b * e2e_steps / t_e2e.item()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        shard_bytes = algorithmic_bytes(w, hi - lo)
        sweep_avg_ms = sweep_ms / max(1, n_sweeps)
        overlapped = eng.post is not None and not args.no_overlap and b < 2
        if overlapped and n_sweeps:
            # Overlapping sweeps: an event pair around a launch also spans the time the kernel waited for the previous
            # query's CTAs to leave the SMs, so the per-launch average is taken as timed region / launches (an upper
            # bound of the kernel's own duration: the region also holds the small kernels), whichever is smaller.
            sweep_avg_ms = min(sweep_avg_ms, ms_total / n_sweeps)
        achieved = shard_bytes / (sweep_avg_ms * 1e-3) / 1e9 if n_sweeps else None
        tensor_bound = b >= 256 and any(int(f) & N.FLAG_TENSOR for o in outs[-1:] for f in o[3].flatten().tolist())
        shard_flops = 2.0 * (hi - lo) * w["d"] * b
        line = {
            "metric": "queries/sec @top-%d" % k, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"float16": "f16", "float32": "f32", "float64": "f64"}[w["dtype"]] + " storage, f32 accumulate"
            if w["dtype"] != "float64" else "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "rows": w["n"], "dim": w["d"], "metric": w["metric"], "top_k": k, "batch": b,
                       "sharding": f"rows/{world}", "l2": "inputs larger than L2 (shard %.2f GB per GPU, a new query every step)"
                       % (shard_bytes / 1e9), "uncertified_steps": uncertified, "cuda_graph": bool(graphed),
                       "exchange": ("peer-memory" if sm.xchg is not None else "nccl all-gather") if world > 1 else None,
                       "pipelined": eng.post is not None, "sweep_overlap": bool(eng.post is not None and not args.no_overlap and b < 2)},
            "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(b * w["d"] * ITEM[w["dtype"]]),
                    "d2h_bytes_per_step": int(b * k * 16 + b * 8), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": (achieved / peaks["hbm_gbs"]) if achieved else None,
                         "traffic": ncu_traffic(args.workload) if world == 1 else None,
                         "kernel": "sweep_kernel" if b < 2 else "batched_tc_kernel (sample + select passes)",
                         "launches_timed": n_sweeps, "avg_launch_ms": sweep_avg_ms,
                         "algorithmic_bytes_per_launch": shard_bytes, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None},
        }
        if tensor_bound and n_sweeps:
            tf = shard_flops / (sweep_avg_ms * 1e-3) / 1e12
            peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
            line["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                                "traffic": ncu_traffic(args.workload) if world == 1 else None, "kernel": "batched_tc_kernel (sample + select passes)",
                                "launches_timed": n_sweeps, "avg_launch_ms": sweep_avg_ms,
                                "algorithmic_flops_per_launch": shard_flops,
                                "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                                "frac_of_burst": tf / peaks.get("bf16_tflops", 1667.5), "frac_of_nominal_2250": tf / 2250.0}
        if not args.no_cpu_baseline and world == 1:
            sample = cpu_sample_rows(w)
            tq = cpu_time_per_query(w, sample)
            line["cpu_baseline"] = {
                "value": 1.0 / (tq * w["n"] / sample), "unit": "queries/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
                "sample": f"1 query on {sample} of {w['n']} rows ({tq:.2f} s), scaled linearly by {w['n'] / sample:.1f}x; "
                          f"oracle/reference_port.rank = the reference's NumPy calls; NumPy {np.__version__}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
