#!/bin/bash
# A/B timing of library variants (tools/build_variant.sh) on ONE box: tools/ab_probe.sh <variant> [<variant> ...]
# "lib" is the in-tree build. Prints per variant: device ms/frame, end-to-end frames/s and the phase split.
mkdir -p gpurun_out
for v in "$@"; do
  d=swraster-viewer_b200/$v
  SWR_LIB_DIR=$PWD/$d python bench.py --steps ${AB_STEPS:-40} --warmup 5 --no-cpu ${AB_ARGS:-} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<PY
import json, sys
v = sys.argv[1]
for ln in open(f"gpurun_out/ab_{v}.json"):
    if ln.startswith("{"):
        d = json.loads(ln)
        print("%-12s ms/frame %.3f  e2e %.1f fps (%.3f ms, sync %.1f)  phases %s" % (v, d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("synchronous_value", 0), {k[3:]: round(x, 3) for k, x in d["roofline"]["phase_ms"].items()}))
PY
  grep "swr dbg" gpurun_out/ab_$v.err | tail -1
done
