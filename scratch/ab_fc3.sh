#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
timeout 300 python bench.py --config c1 --steps 20 --warmup 5 --cpu-tiles 8 > gpurun_out/ab_c1.json 2> gpurun_out/ab_c1.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra-configs --cpu-tiles 8 --profile-csv gpurun_out/ab_layers.csv > gpurun_out/ab_c2.json 2> gpurun_out/ab_c2.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_c1.json"))
print("c1 value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), {k:round(v["ms"],4) for k,v in d["roofline"]["by_kernel"].items()}, d["parity"]["u8_max_abs_lsb"])
d=json.load(open("gpurun_out/ab_c2.json"))
print("c2 value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"], d["parity"]["symbol_mismatches"], d["parity"]["symbol_mismatches_outside_band"])
PY
grep "rd/conv1\|enc_conv1\|hq_enc/conv3" gpurun_out/ab_layers.csv
