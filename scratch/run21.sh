timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print(' ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['quality']['psnr_db'], d['quality']['ssim'], 'metrics ms', round(d['roofline']['hbm_kernels']['metrics_psnr_ssim_f32']['ms'],4))
"; tail -3 gpurun_out/err_$name.log; }
for rep in 1 2; do
run mp0_$rep CIC_METRICS_PIPE=0
run mp1_$rep CIC_METRICS_PIPE=1
done
