for n in 8 16 64; do
timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 --images $n --e2e-chunks 1 --profile-csv gpurun_out/layers_img$n.csv > gpurun_out/bench_img$n.log 2> gpurun_out/err_img$n.log
python -c "
import json
l=[x for x in open('gpurun_out/bench_img$n.log') if x.startswith('{')]
d=json.loads(l[-1]); print($n, 'ms', round(d['ms_per_step'],3), 'layers', round(d['layers_ms_per_step'],3), 'metrics', round(d['roofline']['hbm_kernels']['metrics_psnr_ssim_f32']['ms'],3))"
tail -2 gpurun_out/err_img$n.log
done
