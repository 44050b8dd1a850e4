#!/bin/bash
# round 2, last build: smoke, default bench line (N=1), reference arm, ncu launch lists of the same build
mkdir -p gpurun_out
SECONDS=0; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gpu_tests.log 2>&1; echo "tests rc=$? in ${SECONDS}s"; tail -2 gpurun_out/r02e_gpu_tests.log
SECONDS=0; timeout 600 python __graft_entry__.py smoke > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$? in ${SECONDS}s"; tail -2 gpurun_out/r02e_smoke.log
SECONDS=0; timeout 900 python bench.py --steps 20 --warmup 5 --profile-csv gpurun_out/r02e_layers.csv > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench rc=$? in ${SECONDS}s"; tail -c 300 gpurun_out/r02e_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02e_bench_n1.json"))
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"], "launches", d["gpu_launches"])
print({k:round(v["ms"],3) for k,v in d["roofline"]["by_kernel"].items()}, "frac", round(d["roofline"]["frac"],3))
for k,v in d["configs"].items(): print(k, round(v["value"],1), round(v["e2e"]["value"],1))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/r02e_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra-configs --cpu-tiles 0 > gpurun_out/r02e_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02e_ncu_launches_step.csv python scratch/one_step.py 3 > gpurun_out/r02e_ncu_step.log 2>&1; echo "ncu step rc=$?"
du -sh gpurun_out
