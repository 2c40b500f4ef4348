#!/bin/bash
# Run 23: round-1 evidence pass (lean): GPU tests, smoke, bench (+CPU and eager-GPU baselines), launch list,
# ncu --set full of the hot-path kernels (report kept on the box, CSV pages brought back).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none \
    -k regex:"warp_corr|adaptive_eval|conv_nhwc" -o /tmp/native_full python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/native_full.ncu-rep --page raw --csv > gpurun_out/native_full_raw.csv 2>/dev/null
tail -4 gpurun_out/pytest_gpu.log | cut -c1-160; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
try:
    b=json.load(open("gpurun_out/bench.json"))
    print('value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'launches/step',b['gpu_launches_per_step'],'roofline',{k:(round(v,3) if isinstance(v,float) else v) for k,v in b['roofline'].items() if k in('achieved','frac','traffic','us_per_launch','best_launch_frac','worst_launch_frac')})
    print('cpu_baseline',b['cpu_baseline']['value'],b['cpu_baseline']['cores']); print('gpu_eager',b['gpu_eager_reference']['value']); print('clocks',b['clocks'])
    for r in b['roofline_detail']: print('   in-step',r['entry'],r['shape'],round(r['us'],1),round(r['frac'],3))
    for r in b['roofline_detail_cold_isolated']: print('   isolated',r['entry'],r['shape'],round(r['us'],1),round(r['frac'],3))
except Exception as e: print('bench ERR',e)
PY
tail -n 2 gpurun_out/bench.err; du -sh gpurun_out
