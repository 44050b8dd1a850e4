#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "rate_sweep or stream or phased" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra-configs --cpu-tiles 8 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3), d['clocks'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'roof frac', round(e['frac_of_roof'],3))
print(json.dumps(d['quality']['entropy_coded'])[:500])
PY
