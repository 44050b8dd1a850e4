timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "metrics" 2>&1 | tail -8
python - <<'PY'
import numpy as np, sys
sys.path.insert(0,'.')
import cic_b200 as cic
from oracle import metrics
rng=np.random.default_rng(1)
worst=0
for (n,h,w,s) in [(2,256,256,0.08),(2,256,256,0.01),(2,512,512,0.2),(1,200,311,0.03)]:
    a=cic.synth.to_signed_range(cic.synth.synth_images_u8(n,h,w,seed=h)).astype(np.float32)
    b=np.clip(a+rng.standard_normal(a.shape).astype(np.float32)*np.float32(s),-1,1).astype(np.float32)
    f=cic.ops.metrics_f32(a,b,signed_range=True,fast=True).cpu().numpy(); e=cic.ops.metrics_f32(a,b,signed_range=True).cpu().numpy()
    for i in range(n):
        m=metrics.compute_metrics(a[i],b[i])
        print(h,w,s,'fast-oracle',f[i,1]-m['ssim'],'exact-oracle',e[i,1]-m['ssim'])
PY
run() { name=$1; shift; timeout 300 python bench.py --steps 8 --warmup 3 --cpu-tiles 0 "$@" > gpurun_out/bench_$name.log 2> gpurun_out/err_$name.log; echo -n "$name: "; python -c "
import json
l=[x for x in open('gpurun_out/bench_$name.log') if x.startswith('{')]
d=json.loads(l[-1]); print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value'],1), d['quality']['ssim'], {k:(round(v['ms'],3), round(v['frac_of_hbm_peak'],3)) for k,v in d['roofline']['hbm_kernels'].items()})
"; tail -3 gpurun_out/err_$name.log; }
run ssim_fast --ssim fast
run ssim_exact --ssim exact
