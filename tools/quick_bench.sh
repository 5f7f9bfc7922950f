#!/bin/bash
# usage: tools/quick_bench.sh tag [lib]   -> one-line summary of a short bench run
tag=$1; lib=$2
SWO_LIB=$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python - "$tag" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print(sys.argv[1], "kernel G/s %.2f  ms %.2f  e2e G/s %.3f  frac %.4f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["roofline"]["frac"]))
PY
