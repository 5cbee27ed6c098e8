#!/bin/bash
O=gpurun_out
( time timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus ${NGPU:-8} --steps 200 --warmup 10 > $O/default${NGPU:-8}.json 2> $O/default${NGPU:-8}.err ) 2>&1 | tail -3
python - "$O/default${NGPU:-8}.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("HEAD value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f launches=%d unc=%s parity=%s"%(d['value'],d['e2e']['value'],d['e2e']['sync_value'],d['ms_per_step'],d['roofline']['frac'] or 0,d['gpu_launches'],d['config']['uncertified_steps'],d.get('parity_check',{}).get('ok')))
for e in d.get('extra',[]):
    if 'error' in e or 'roofline' not in e: print("   extra", e); continue
    print("   extra %-30s value=%.1f e2e=%.1f sync=%.1f ms=%.4f frac=%.3f unc=%s"%(e['workload'],e['value'],e['e2e']['value'],e['e2e']['sync_value'],e['ms_per_step'],e['roofline']['frac'] or 0,e['uncertified_steps']))
PY
tail -n 4 $O/default${NGPU:-8}.err
