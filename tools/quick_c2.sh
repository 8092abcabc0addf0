#!/bin/bash
# quick GPU check of a Chamfer kernel change: bit-exact tests + stage timings.  Usage: gpurun -- 'bash tools/quick_c2.sh TAG'
TAG=${1:-x}
timeout 600 python -m pytest tests -m gpu -x -q -k "chamfer" 2>&1 | tail -3
for w in c2 c5; do python bench.py --workload $w --no-cpu-baseline --steps 10 > gpurun_out/bench_${w}_$TAG.log 2>&1; echo $w rc=$?; done
python - <<PY
import json
for w in ("c2","c5"):
    l=[x for x in open(f"gpurun_out/bench_{w}_$TAG.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print(w, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k: round(v,3) for k,v in d["roofline"]["forward_total"]["stages_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
