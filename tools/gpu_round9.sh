#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=12 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --height 1184 --width 1600 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1600.json 2> gpurun_out/bench_1600.err
echo "bench1600 exit $?" >> gpurun_out/bench_1600.err
timeout 600 python bench.py --batch 8 --steps 10 --warmup 3 --no-cpu-baseline --slots 1 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err
echo "benchb8 exit $?" >> gpurun_out/bench_b8.err
grep -A16 "slowest" gpurun_out/pytest_gpu.log | cut -c1-150; tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ("bench_1600","bench_b8"):
    try:
        b=json.load(open(f"gpurun_out/{f}.json")); print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],2),'ka frac',round(b['roofline']['frac'],3),'best',round(b['roofline']['best_launch_frac'],3))
        for r in b['roofline_detail']: print('   ',r['entry'],r['shape'],round(r['us'],1),round(r['frac'],3))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/bench_1600.err gpurun_out/bench_b8.err
