#!/bin/bash
# Round 2 / run 2: K-A generation 4 (TMA-staged windows) on hardware for the first time -- parity first (kernel-level A/B
# against generation 3 and the oracle, then the whole GPU suite and the fp32-default bench-mode parity), then timing:
# kbench sweep (gen 4 consumer warps x resident CTAs vs gen 3), the bench line in the new default (fp32-accurate convs,
# all native), convbench with the conversion-free 3xTF32 split, launch list + ncu --set full of the PatchMatch kernels.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bench_mode_parity.json
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -p no:cacheprovider -k "warp_corr or fused_heads" > gpurun_out/pytest_ka.log 2>&1
echo "pytest K-A exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_ka.log
tail -3 gpurun_out/pytest_ka.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider --durations=8 -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
grep -h "bench-mode parity" gpurun_out/pytest_gpu.log | cut -c1-420
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
KB_SWEEP_EVAL=0 timeout 600 python tools/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err
echo "kbench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/kbench.err
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/bench.err
timeout 400 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"warp_corr|adaptive_eval" -o gpurun_out/native_full python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/native_full.ncu-rep --page raw --csv > gpurun_out/native_full_raw.csv 2>/dev/null
python - <<'PY'
import json
try:
    b=json.load(open("gpurun_out/bench.json"))
    print('value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'roofline frac',round(b['roofline']['frac'],3),'repeats',b['repeats']['ms_per_step_all'])
    print('roofline detail', [(r['shape'], round(r['us'],1), round(r['frac'],3)) for r in b['roofline_detail']])
    print('latency',b['latency_single_request'])
    for k in ('value_tf32','cfg3_1600x1184','batch8_640x512'):
        v=b.get(k); print(' ',k,{kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in (v or {}).items() if kk in ('value','e2e_value','ms_per_step','error')}, (v or {}).get('roofline'))
    print('clocks',b['clocks'], 'launches', b['gpu_launches_per_step'], b['native_kernels_per_step'])
except Exception as e: print('bench ERR',e)
try:
    for r in json.load(open("gpurun_out/kbench.json"))['rows']:
        if 'warp_corr' in r['call']:
            print('  ',r['call'],'default',r['default_us'])
            for k,v in r.items():
                if isinstance(v,list) and k!='default_us': print('        ',k,v)
        else: print('  ',r['call'],'default',r['default_us'])
except Exception as e: print('kbench ERR',e)
try:
    j=json.load(open("gpurun_out/convbench.json"))
    for r in j['layers']: print('  conv',r['layer'],'cudnn_tf32',r['cudnn_tf32_us']['cold'],'p1',r['native_p1_mt0_us']['cold'],'p3',r['native_p3_mt0_us']['cold'])
except Exception as e: print('convbench ERR',e)
PY
echo "done at $(( $(date +%s) - t0 )) s"; du -sh gpurun_out
