#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra-configs --cpu-tiles 0 > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2f_bench.err
grep "stream batch" gpurun_out/r2f_bench.log | awk '{print NR": "$0}' | sed -n '1,200p' | awk 'NR%1==0' | head -80
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.log').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'], d['clocks'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'floor', round(e['host_copy_floor_ms'],2), 'roof frac', round(e['frac_of_roof'],3))
PY
