#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_full_size.py tests/test_gpu_tc_ops.py -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2l_tests.log
for v in "CIC_GEN_TAIL=1" "CIC_GEN_TAIL=0"; do
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-extra-configs --cpu-tiles 8 --profile-csv gpurun_out/r2l_layers_${v//=/_}.csv > gpurun_out/r2l_${v//=/_}.json 2> gpurun_out/r2l_err.log
  python - "$v" <<'PY'
import csv,sys,json
v=sys.argv[1].replace('=','_')
rows=list(csv.DictReader(open(f'gpurun_out/r2l_layers_{v}.csv')))
d=json.loads(open(f'gpurun_out/r2l_{v}.json').read().strip().splitlines()[-1])
sel=[r for r in rows if ('conv_out' in r['layer'] or r['layer'] in ('roi_blend','gen_tail','hq_gen/deconv4','hq_gen/deconv3'))]
print(v, 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), ' '.join(f"{r['layer']}={float(r['ms']):.3f}" for r in sel), 'parity', d['parity']['symbol_mismatches'], d['parity']['recon_max_abs_01'], d['parity']['psnr_delta_db_max'], d['parity']['dt_max_abs'], d['parity']['hq_ratio_delta_max'])
PY
done
tail -3 gpurun_out/r2l_err.log
