#!/bin/bash
O=gpurun_out
run() { # tag args...
  tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 "$@" > $O/r02m_$tag.json 2> $O/r02m_$tag.err
}
run default --steps 200 --warmup 10
run hamming --workload c5_hamming_b1 --steps 200 --warmup 10 --no-cpu-baseline
run c3_nccl --steps 100 --warmup 10 --exchange nccl --extras none
for f in $O/r02m_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f launches=%d unc=%s parity=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['gpu_launches'],d['config']['uncertified_steps'],d.get('parity_check',{}).get('ok')))
    for e in d.get('extra',[]):
        if 'error' in e: print("   extra", e); continue
        print("   extra %-30s value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f unc=%s"%(e['workload'],e['value'],e['e2e']['value'],e['e2e']['sync_value'],e['ms_per_step'],e['roofline']['frac'] or 0,e['uncertified_steps']))
except Exception as e:
    print(sys.argv[1],"ERR",e)
PY
done
tail -n 5 $O/r02m_*.err
