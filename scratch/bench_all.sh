#!/bin/bash
# every single-GPU workload of bench.py -> <out>/bench_<workload>.json (one JSON line each)
OUT=${1:-gpurun_out/r02/final}
mkdir -p $OUT
for w in c3_cosine_b1 c3_dot_b1 c3_pearson_b1 c3_cosine_b64 c3_cosine_b4096 c2_cosine_b1 c2_cosine_b1024 c5_euclid_b1 c5_manhattan_b1 c5_hamming_b1 c5_euclid_b1024 c4_decay_mask_k100; do
  extra="--no-cpu-baseline"
  [ "$w" = "c3_cosine_b1" ] && extra=""
  [ "$w" = "c2_cosine_b1" ] && extra=""
  timeout 400 python bench.py --workload $w $extra > $OUT/bench_$w.json 2> $OUT/bench_$w.err || echo "FAILED $w"
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_c3.json 2> $OUT/bench_reference_c3.err
OUTDIR=$OUT python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.path.join(os.environ.get("OUTDIR","gpurun_out/r02/final"), "bench_*.json"))):
    try:
        j = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j.get("roofline") or {}
    print(os.path.basename(f)[6:-5], round(j["value"], 4), "ms/step", round(j["ms_per_step"], 4), r.get("bound"), r.get("achieved") and round(r["achieved"], 1),
          r.get("frac") and round(r["frac"], 3), "e2e", round(j["e2e"]["value"], 4), "unc", j["config"].get("uncertified_steps"), (j.get("clocks") or {}).get("sm_mhz"))
PY
