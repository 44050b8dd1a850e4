#!/bin/bash
# usage: bash scratch/r2_scale_jpeg.sh N   (headline config only, under torchrun on N GPUs; adds the JPEG-out leg)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-extra-configs > gpurun_out/r02_bench_n${N}_jpeg.json 2> gpurun_out/r02_bench_n${N}_jpeg.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n${N}_jpeg.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads(open(f'gpurun_out/r02_bench_n{n}_jpeg.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],3), d['clocks'])
for k in ('e2e','e2e_single_call','e2e_f32_io','e2e_u8_io_no_dt','e2e_u8_in_jpeg_out'):
    e=d[k]; print(k, round(e['value'],1), 'ms', round(e['ms_per_step'],3), 'floor', round(e['host_copy_floor_ms'],2), 'd2h', e['d2h_bytes_per_step'], 'host GB/s', round(e['host_gbs_achieved'],1))
PY
