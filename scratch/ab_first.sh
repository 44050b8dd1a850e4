#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/ab_tests.log
for f in 1 0; do
  CIC_FIRST_TC=$f timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --cpu-tiles 8 > gpurun_out/ab_c1_$f.json 2> gpurun_out/ab_c1_$f.err
  CIC_FIRST_TC=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-extra-configs --cpu-tiles 8 --profile-csv gpurun_out/ab_layers_$f.csv > gpurun_out/ab_c2_$f.json 2> gpurun_out/ab_c2_$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_c1_$f.json"))
print("first_tc=$f c1 value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["roofline"]["by_kernel"], d["parity"])
d=json.load(open("gpurun_out/ab_c2_$f.json"))
print("first_tc=$f c2 value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"], {k:v["ms"] for k,v in d["roofline"]["by_kernel"].items()}, d["parity"]["symbol_mismatches"], d["parity"]["symbol_mismatches_outside_band"])
PY
  grep "rd/" gpurun_out/ab_layers_$f.csv
done
