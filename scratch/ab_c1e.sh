#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_ops.py tests/test_gpu_models.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-extra-configs --cpu-tiles 8 --profile-csv gpurun_out/ab_layers_$i.csv > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$i.json"))
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"], "parity", d["parity"]["symbol_mismatches"], d["parity"]["symbol_mismatches_outside_band"])
PY
  grep "enc_conv1\|rd/conv1\|hq_enc/conv3" gpurun_out/ab_layers_$i.csv
done
