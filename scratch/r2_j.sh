#!/bin/bash
mkdir -p gpurun_out
for v in "base" "CIC_TC_DC2=1" "CIC_TC_DC2=2" "CIC_TC_MERGE=0"; do
  if [ "$v" = "base" ]; then envs=""; else envs="$v"; fi
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-extra-configs --cpu-tiles 0 --profile-csv gpurun_out/r2j_layers_${v//=/_}.csv > gpurun_out/r2j_${v//=/_}.json 2> gpurun_out/r2j_err.log
  python - "$v" <<'PY'
import csv,sys,json
v=sys.argv[1].replace('=','_')
rows=list(csv.DictReader(open(f'gpurun_out/r2j_layers_{v}.csv')))
d=json.loads(open(f'gpurun_out/r2j_{v}.json').read().strip().splitlines()[-1])
sel=[r for r in rows if r['layer'] in ('hq_gen/deconv3','hq_gen/deconv4','hq_gen/deconv2','rd/conv1','rd/conv2')]
print(v, 'step', round(d['ms_per_step'],3), ' '.join(f"{r['layer']}={float(r['ms']):.3f}({r['kernel'][:9]})" for r in sel))
PY
done
