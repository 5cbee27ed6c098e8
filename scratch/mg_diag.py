"""Where does a sharded step go?  torchrun --nproc-per-node 2 scratch/mg_diag.py  (hamming 5M x 1024 bits, split over the ranks)"""
import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "local-hyperdb_b200")]
import hyperdb_b200 as hb
from hyperdb_b200.sharded import CudaEngine, ShardedMatrix, shard_bounds, packed_len
os.environ.setdefault("NCCL_DEBUG", "WARN")
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
n, d, metric, k = 5_000_000, 1024, sys.argv[1] if len(sys.argv) > 1 else "hamming_distance", 10
lo, hi = shard_bounds(n, world, rank)
g = torch.Generator(device=dev); g.manual_seed(rank)
V = torch.randn((hi - lo, d), generator=g, device=dev)
m = hb.DeviceMatrix(V, row_offset=lo)
eng = CudaEngine(m); sm = ShardedMatrix(eng, n); eng.enable_pipeline()
Q = torch.randn((400, d), device=dev)
def timed(fn, steps=200, warm=20):
    for i in range(warm): fn(i)
    sm.wait_results(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(steps): fn(warm + i)
    t_host = time.perf_counter() - t0
    sm.wait_results(); e1.record(); torch.cuda.synchronize()
    return round(t_host / steps * 1e6, 1), round(e0.elapsed_time(e1) / steps * 1e3, 1)
res = {}
res["full query_async (host us, device us)"] = timed(lambda i: sm.query_async(Q[i:i + 1], k, metric, 0.0))
res["local_topk only"] = timed(lambda i: eng.local_topk(Q[i:i + 1], k, metric, 0.0))
buf = torch.zeros(packed_len(1, k), dtype=torch.int64, device=dev)
gathered = torch.empty(world * buf.numel(), dtype=torch.int64, device=dev)
if world > 1:
    res["all_gather only (main stream)"] = timed(lambda i: dist.all_gather_into_tensor(gathered, buf))
    def lt_ag(i):
        mine = eng.local_topk(Q[i:i + 1], k, metric, 0.0)
        with torch.cuda.stream(eng.post):
            dist.all_gather_into_tensor(gathered, mine)
    res["local_topk + all_gather on post"] = timed(lt_ag)
res["merge only"] = timed(lambda i: eng.merge(gathered.view(world, -1), 1, k))
if rank == 0: print(json.dumps(res, indent=1))
if world > 1: dist.barrier(); dist.destroy_process_group()
