#!/bin/bash
# A/B of the RD side-stream fork and the new conv1 builder (knob build of libcic.so): tests first, then bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_tc_ops.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/ab_tests.log
for f in 1 0 1 0; do
  CIC_RD_FORK=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-extra-configs --cpu-tiles 0 --profile-csv gpurun_out/ab_layers_f$f.csv > gpurun_out/ab_f$f.json 2> gpurun_out/ab_f$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_f$f.json"))
print("fork=$f value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), "single", round(d["e2e_single_call"]["value"],1), "clk", d["clocks"]["sm_mhz"], "conv1", d["roofline"]["by_kernel"]["conv1_tc_kernel"]["ms"])
PY
done
