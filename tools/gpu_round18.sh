#!/bin/bash
# Run 18: hybrid conv policy (native where memory-bound, cuDNN where FLOP-bound): full GPU tests, bench A/B/C, launch list.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_hybrid.json 2> gpurun_out/bench_hybrid.err
PMB200_NATIVE_CONVS=all timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_allnative.json 2> gpurun_out/bench_allnative.err
PMB200_NATIVE_CONVS=0 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_cudnn.json 2> gpurun_out/bench_cudnn.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("bench_hybrid","bench_allnative","bench_cudnn"):
    try:
        b=json.load(open(f"gpurun_out/{f}.json")); print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'launches/step',b['gpu_launches_per_step'],'ka frac',round(b['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
tail -n 2 gpurun_out/bench_hybrid.err
