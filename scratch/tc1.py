import sys, time, numpy as np, torch
sys.path[:0]=['.','local-hyperdb_b200']
import hyperdb_b200 as hb
torch.manual_seed(0)
n, d, B = 600_000, 128, 40
dev = torch.device("cuda", 0)
V = torch.randn(n, d, device=dev)
V = (V / V.norm(dim=1, keepdim=True)).half()
Q = torch.randn(B, d, device=dev)
Q = (Q / Q.norm(dim=1, keepdim=True)).half()
m = hb.DeviceMatrix(V)
q_np = Q.cpu().numpy()
for metric in ("dot_product", "cosine_similarity"):
    m.set_path(2)
    i0, s0, c0, f0 = m.query(q_np, 10, metric)
    m.set_path(0)
    t0 = time.time()
    i1, s1, c1, f1 = m.query(q_np, 10, metric)
    print(metric, "tensor flags", set(f1.tolist()), "sweep flags", set(f0.tolist()), "time", time.time() - t0)
    print("  idx equal", np.array_equal(i0, i1), "scores equal", np.array_equal(s0, s1))
    if not np.array_equal(i0, i1):
        bad = np.flatnonzero((i0 != i1).any(axis=1))
        print("  bad queries", bad[:10]); b = bad[0]; print(i0[b], i1[b]); print(s0[b], s1[b])
    # direct check of tensor results vs torch
    ref = (V.float() @ Q.float().T)          # [n, B]
    top = ref.topk(10, dim=0)
    print("  torch top idx match", np.array_equal(np.sort(top.indices.T.cpu().numpy(), axis=1), np.sort(i1, axis=1)))
m.close()
