#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "stream or phased or pipelined" > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra-configs > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'floor', round(e['host_copy_floor_ms'],2), 'roof frac', round(e['frac_of_roof'],3))
print(d['clocks'])
PY
