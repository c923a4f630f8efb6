#!/bin/bash
# Bench lines + GPU test log of one round on ONE box (run under gpurun): gpurun_out/<tag>_*.json / .log
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/${tag}_gpu_tests.log 2>&1
tail -3 $out/${tag}_gpu_tests.log
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
for w in C2 C4 C5; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_$(echo $w | tr A-Z a-z).json 2> $out/${tag}_bench_$w.err
done
python tools/bench_vae.py 8 $out/${tag}_vae_decode_breakdown.json > $out/${tag}_vae_decode_breakdown.txt 2>&1
python - <<PY
import json
for w in ("c3", "c2", "c4", "c5"):
    try:
        d = json.load(open("$out/${tag}_bench_%s.json" % w))
        print(w, round(d["ms_per_step"], 2), "ms", d["clocks"]["sm_mhz"], "MHz", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(w, "failed", e)
PY
