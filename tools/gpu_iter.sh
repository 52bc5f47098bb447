#!/bin/bash
# Development iteration on one B200 (through gpurun): GPU parity suite, then a short bench line without the CPU leg.
# usage: tools/gpu_iter.sh <tag> [extra bench args]
tag=${1:-iter}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_tests.log
python bench.py --steps 30 --warmup 5 --no-cpu "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_tests.log
python - <<PY
import json
for ln in open("gpurun_out/${tag}_bench.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("ms/frame %.3f  fps %.1f  e2e %.1f (%.3f ms)  phases %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["phase_ms"].items()}))
PY
tail -3 gpurun_out/${tag}_bench.err
