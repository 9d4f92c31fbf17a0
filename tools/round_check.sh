#!/bin/bash
# Standard verification pass for ONE gpurun call on a single B200 (about 5 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round_check.sh'
# Logs and JSON lines land in gpurun_out/ and are merged back.  Pass "c4" as first argument to add
# the 100 M-DOF emi_3d solve (about 2.5 more minutes).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c3.json"))
print("C3", round(d["value"]), "DOF/s", round(d["ms_per_step"], 1), "ms/solve", d["iterations"], "its",
      "vcycle", round(d["vcycle_ms"], 2), "ms", "e2e", round(d["e2e"]["value"]))
print({k: (v["ms"], v["share"], v["alg_GBs"]) for k, v in d["kernels"].items()})
print(d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), d["clocks"], d["host"])
PY
if [ "${1:-}" = "c4" ]; then
  timeout 700 python bench.py --workload emi_3d -n 464 --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
  echo "c4 rc=$?"
  python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c4.json"))
print("C4", round(d["value"]), "DOF/s", round(d["ms_per_step"], 1), "ms/solve", d["iterations"], "its",
      "vcycle", round(d["vcycle_ms"], 2), "ms")
print({k: (v["ms"], v["share"], v["alg_GBs"]) for k, v in d["kernels"].items()})
print(d["gs_ms_by_level"][:8], d["host"])
PY
fi
