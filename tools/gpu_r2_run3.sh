#!/bin/bash
# Round 2 / run 3: first hardware run of K-D5 (tcgen05 / TMEM / TMA conv) -- its parity tests alone first, under a short
# timeout (a descriptor mistake must cost seconds, not the box) -- then the whole GPU suite, K-A generation 4 with the
# pipelined producer (kbench sweep incl. ring depth; also at batch 8 where run 2 showed a 7 ms launch), convbench with the
# tc5 column, the bench line.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bench_mode_parity.json
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 240 python -m pytest tests/test_conv.py -m gpu -x -q --tb=short -p no:cacheprovider -k "tc5" > gpurun_out/pytest_tc5.log 2>&1
echo "pytest tc5 exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_tc5.log
tail -12 gpurun_out/pytest_tc5.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=6 -s --deselect tests/test_conv.py::test_gpu_tc5_conv_matches_cudnn_fp32 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
grep -h "bench-mode parity" gpurun_out/pytest_gpu.log | cut -c1-300
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err
timeout 500 python tools/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err
KB_B=8 KB_SWEEP_EVAL=0 timeout 500 python tools/kbench.py > gpurun_out/kbench_b8.json 2> gpurun_out/kbench_b8.err
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/bench.err
PMB200_TC5=0 timeout 300 python bench.py --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_no_tc5.json 2> gpurun_out/bench_no_tc5.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/convbench.json"))
    for r in j['layers']: print('  conv',r['layer'],r['shape'],'cudnn_tf32',r['cudnn_tf32_us']['cold'],'p1',r['native_p1_mt0_us']['cold'],'p3',r['native_p3_mt0_us']['cold'],'tc5',(r.get('tc5_3xtf32_us') or {}).get('cold') if isinstance(r.get('tc5_3xtf32_us'),dict) else r.get('tc5_3xtf32_us'), r.get('tc5_tflops_3x'))
except Exception as e: print('convbench ERR',e)
for f in ("bench.json","bench_no_tc5.json"):
    try:
        b=json.load(open("gpurun_out/"+f))
        print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'roofline frac',round(b['roofline']['frac'],3), b['native_kernels_per_step'])
        for k in ('value_tf32','cfg3_1600x1184','batch8_640x512'):
            v=b.get(k)
            if v: print(' ',k,{kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','e2e_value','ms_per_step','error')}, (v.get('roofline') or {}).get('frac'))
    except Exception as e: print(f,'ERR',e)
for f in ("kbench.json","kbench_b8.json"):
    try:
        for r in json.load(open("gpurun_out/"+f))['rows']:
            if 'warp_corr' in r['call'] or 'adaptive_eval' in r['call']:
                print('  ',f,r['call'],'default',r['default_us'])
                for k,v in r.items():
                    if isinstance(v,list) and k!='default_us': print('        ',k,v)
    except Exception as e: print(f,'ERR',e)
PY
echo "done at $(( $(date +%s) - t0 )) s"; du -sh gpurun_out
