#!/bin/bash
# round 2, last session: full GPU test suite, smoke, default bench line (N=1) with the per-layer CSV
mkdir -p gpurun_out
SECONDS=0; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -3 gpurun_out/r02b_gpu_tests.log
SECONDS=0; timeout 600 python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$? in ${SECONDS}s"; tail -5 gpurun_out/r02b_smoke.log
SECONDS=0; timeout 900 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/r02b_layers.csv > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$? in ${SECONDS}s"; tail -c 600 gpurun_out/r02b_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02b_bench_n1.json"))
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"], "launches", d["gpu_launches"])
print("input_stage", d["quality"].get("input_stage"))
print("parity", d["parity"])
for k,v in d["configs"].items(): print(k, round(v["value"],1), round(v["e2e"]["value"],1), v.get("parity",{}).get("symbol_mismatches_outside_band"), (v.get("quality") or {}).get("input_stage"))
PY
