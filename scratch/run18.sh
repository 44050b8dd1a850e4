timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
CIC_TC_DC2=1 timeout 600 python -m pytest tests/test_gpu_tc_ops.py -m gpu -x -q -k "transpose" 2>&1 | tail -2
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --cpu-tiles 0 --profile-csv gpurun_out/layers_$name.csv > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; grep -E "deconv[34]" gpurun_out/layers_$name.csv | cut -d, -f2 | tr '\n' ' '; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print(' ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['quality']['psnr_db'])
"; tail -3 gpurun_out/err_$name.log; }
for rep in 1 2; do
run sv_$rep CIC_TC_DC2=0
done
B="python bench.py --steps 3 --warmup 3 --cpu-tiles 0"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 260 -c 110 --csv --log-file gpurun_out/r01_launches_dram.csv $B > gpurun_out/ncu_ld.log 2>&1; echo "ncu dram list rc $?"
CIC_TC_DC2=1 ncu --set full --clock-control none --import-source on -k regex:tc_deconv2 -s 2 -c 1 -o gpurun_out/prof_r01_dc2 $B > gpurun_out/ncu_dc2.log 2>&1; echo "ncu dc2 rc $?"
ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 10 -c 1 -o gpurun_out/prof_r01_deconv4_sv $B > gpurun_out/ncu_dc4sv.log 2>&1; echo "ncu dc4 rc $?"
