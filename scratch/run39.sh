timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "phased" 2>&1 | tail -4
timeout 300 python bench.py --steps 8 --warmup 3 --cpu-tiles 0 > gpurun_out/bench_u8.log 2> gpurun_out/err_u8.log; tail -3 gpurun_out/err_u8.log
python -c "
import json
l=[x for x in open('gpurun_out/bench_u8.log') if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', d['e2e'], 'u8', d['e2e_u8_io'])"
