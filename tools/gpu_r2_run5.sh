#!/bin/bash
# Round 2 / run 5: K-D5 with dedicated epilogue warps and two TMEM accumulators (parity first, short timeout), the whole GPU
# suite (fusion half of f4 included), convbench, the bench line with the measured tc5 allowlist and with every supported
# layer on tc5, launch list.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bench_mode_parity.json
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 240 python -m pytest tests/test_conv.py -m gpu -x -q --tb=short -p no:cacheprovider -k "tc5" > gpurun_out/pytest_tc5.log 2>&1
echo "pytest tc5 exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_tc5.log
tail -5 gpurun_out/pytest_tc5.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=6 -s --deselect tests/test_conv.py::test_gpu_tc5_conv_matches_cudnn_fp32 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
grep -h "bench-mode parity" gpurun_out/pytest_gpu.log | cut -c1-260
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/bench.err
PMB200_TC5=all timeout 300 python bench.py --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_tc5_all.json 2> gpurun_out/bench_tc5_all.err
PMB200_TC5=0 timeout 300 python bench.py --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_no_tc5.json 2> gpurun_out/bench_no_tc5.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/convbench.json"))
    for r in j['layers']: print('  conv',r['layer'],r['shape'],'p1',r['native_p1_mt0_us']['cold'],'p3',r['native_p3_mt0_us']['cold'],'tc5',(r.get('tc5_3xtf32_us') or {}).get('cold') if isinstance(r.get('tc5_3xtf32_us'),dict) else r.get('tc5_3xtf32_us'), r.get('tc5_tflops_3x'))
except Exception as e: print('convbench ERR',e)
for f in ("bench.json","bench_tc5_all.json","bench_no_tc5.json"):
    try:
        b=json.load(open("gpurun_out/"+f))
        print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'roofline frac',round(b['roofline']['frac'],3), b['native_kernels_per_step'])
        for k in ('value_tf32','cfg3_1600x1184','batch8_640x512'):
            v=b.get(k)
            if v: print(' ',k,{kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','e2e_value','ms_per_step','error')}, (v.get('roofline') or {}).get('frac'))
    except Exception as e: print(f,'ERR',e)
PY
echo "done at $(( $(date +%s) - t0 )) s"; du -sh gpurun_out
