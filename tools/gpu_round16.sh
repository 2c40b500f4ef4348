#!/bin/bash
# Run 16: native channels-last tensor-core convs -- parity, per-layer micro-benchmark, bench A/B, launch list.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_conv.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_conv.log 2>&1
echo "pytest conv exit $?" >> gpurun_out/pytest_conv.log
timeout 600 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err
echo "convbench exit $?" >> gpurun_out/convbench.err
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_conv.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err
echo "bench native exit $?" >> gpurun_out/bench_native.err
PMB200_NATIVE_CONVS=0 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_cudnn.json 2> gpurun_out/bench_cudnn.err
echo "bench cudnn exit $?" >> gpurun_out/bench_cudnn.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
tail -5 gpurun_out/pytest_conv.log; tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/convbench.err
python - <<'PY'
import json
for f in ("bench_native","bench_cudnn"):
    try:
        b=json.load(open(f"gpurun_out/{f}.json")); print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'launches/step',b['gpu_launches_per_step'],'ka frac',round(b['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
try:
    c=json.load(open("gpurun_out/convbench.json")); print(c["sum_cold_us"])
    for r in c["layers"]: print(r["layer"], r["shape"], 'cudnn', r["cudnn_tf32_us"]["cold"], 'native', r.get("native_p1_mt0_us"), 'x', r.get("speedup_vs_cudnn_cold"), 'GB/s', r.get("native_tf32_gbs"))
except Exception as e: print('convbench ERR', e)
PY
tail -3 gpurun_out/bench_native.err gpurun_out/bench_cudnn.err
