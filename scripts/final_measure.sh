#!/bin/bash
# GPU box: the measurement set committed under profiles/ at the end of a round.
#   bash scripts/final_measure.sh gpurun_out/<run>
set -u
out=$1
mkdir -p $out
timeout 400 python bench.py 2> $out/bench.err | tail -1 > $out/bench.json
timeout 400 python bench.py --impl reference 2>> $out/bench.err | tail -1 > $out/bench_reference.json
for w in single30k kitti120k micro1m; do
  timeout 300 python bench.py --workload $w 2>> $out/bench.err | tail -1 > $out/bench_$w.json
done
# launch list of the bench command itself (2 steps): per-launch durations, cold caches, serialised
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
bash scripts/ncu_all_kernels.sh $out
python - <<PY
import json
for n in ("bench", "bench_reference", "bench_single30k", "bench_kitti120k", "bench_micro1m"):
    try:
        d = json.load(open("$out/%s.json" % n))
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"),
              (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "FAILED", e)
PY
