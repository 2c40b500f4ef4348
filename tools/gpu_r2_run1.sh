#!/bin/bash
# Round 2 / run 1: parity of the timed configuration first (tests/test_gpu_bench_mode.py), then the whole GPU suite, the new
# bench line (repeats, fp32 value, cfg-3 / batch-8 sub-records, latency), the batch x slots grid, kernel sweeps of what was
# changed blind at the end of round 1 (K-A 32-bit tap indices / MINB=5 / DC_VW=16, K-D column staging, transposed-operand
# conv) and fresh ncu evidence.  Every step writes its own file so that whatever finishes before the limit is kept.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/bench_mode_parity.json
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bench_mode.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_bench_mode.log 2>&1
echo "pytest bench-mode exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_bench_mode.log
grep -h "bench-mode parity\|passed\|failed" gpurun_out/pytest_bench_mode.log | cut -c1-400
timeout 700 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider --durations=8 --deselect tests/test_gpu_bench_mode.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/bench.err
for cfg in "1 1" "2 2" "4 1" "4 2" "8 1"; do
  set -- $cfg
  timeout 200 python bench.py --batch $1 --slots $2 --repeats 3 --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_b$1_s$2.json 2> gpurun_out/bench_b$1_s$2.err
done
timeout 600 python tools/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err          # full K-A / K-B sweeps
timeout 400 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err  # per-layer native vs cuDNN
PMB200_CONV_T=1 timeout 400 python tools/convbench.py > gpurun_out/convbench_transposed.json 2> gpurun_out/convbench_transposed.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"warp_corr|adaptive_eval|init_propagate|offset_corr" -o /tmp/native_full python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/native_full.ncu-rep --page raw --csv > gpurun_out/native_full_raw.csv 2>/dev/null
python - <<'PY'
import json
try:
    b=json.load(open("gpurun_out/bench.json"))
    print('value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'roofline frac',round(b['roofline']['frac'],3),'repeats',b['repeats']['ms_per_step_all'])
    print('latency',b['latency_single_request'],'numa',b['config']['numa'])
    for k in ('value_fp32','cfg3_1600x1184','batch8_640x512'):
        v=b.get(k); print(' ',k,{kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in (v or {}).items() if kk in ('value','e2e_value','ms_per_step','error')}, (v or {}).get('roofline'))
    print('clocks',b['clocks'])
except Exception as e: print('bench ERR',e)
for tag in ("b1_s1","b2_s2","b4_s1","b4_s2","b8_s1"):
    try:
        b=json.load(open(f"gpurun_out/bench_{tag}.json")); print('  batch/slots',tag,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1))
    except Exception as e: print('  ',tag,'ERR',e)
try:
    for r in json.load(open("gpurun_out/kbench.json"))['rows']:
        best=min(((v[0],k) for k,v in r.items() if isinstance(v,list)),default=None)
        print('  ',r['call'],'default',r['default_us'],'best',best)
except Exception as e: print('kbench ERR',e)
PY
echo "done at $(( $(date +%s) - t0 )) s"; du -sh gpurun_out
