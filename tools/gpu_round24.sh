#!/bin/bash
# Run 24 (short budget: ~5 GPU-minutes left in the round): GPU tests with the new rows first (f4 geometric filter,
# f5 map I/O), then the whole GPU suite, the bench line, default kernel timings and the f4 tool.  Every step writes its
# own file so that whatever finishes before the limit is kept.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
echo "build $(( $(date +%s) - t0 )) s" >> gpurun_out/build.log
timeout 200 python -m pytest tests/test_geo.py tests/test_mapio.py tests -m gpu -x -q --tb=short -p no:cacheprovider --durations=6 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 150 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/bench.err
KB_SWEEP_EVAL=0 KB_SWEEP_KA=0 timeout 60 python tools/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err
echo "kbench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/kbench.err
timeout 60 python tools/geobench.py > gpurun_out/geobench.json 2> gpurun_out/geobench.err
echo "geobench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/geobench.err
timeout 60 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python - <<'PY'
import json
try:
    b=json.load(open("gpurun_out/bench.json"))
    print('value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'roofline',{k:(round(v,3) if isinstance(v,float) else v) for k,v in b['roofline'].items() if k in('achieved','frac','us_per_launch')})
    print('cpu_baseline',b['cpu_baseline']['value'],b['cpu_baseline']['cores'],'clocks',b['clocks'])
except Exception as e: print('bench ERR',e)
try:
    for r in json.load(open("gpurun_out/kbench.json"))['rows']: print('  ',r['call'],r['default_us'])
except Exception as e: print('kbench ERR',e)
try:
    for r in json.load(open("gpurun_out/geobench.json"))['rows']: print('  geo',r['shape'],r['gpu_us'],r['frac_of_hbm_peak'],r['speedup_e2e'],r['mask_count_mismatch_px'])
except Exception as e: print('geobench ERR',e)
PY
tail -n 2 gpurun_out/bench.err gpurun_out/smoke.log
