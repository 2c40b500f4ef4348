#!/bin/bash
# Round 2 evidence pass on ONE B200: build, the whole GPU suite, smoke, the default bench line (sub-records, CPU baseline),
# config-5 training step at N = 1, the ncu launch lists (one eager forward, and the bench command itself) and the
# `--set full` capture of the PatchMatch kernels that profiles/r2_warp_corr_traffic.json is written from.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/bench.err
timeout 400 python tools/train_step.py --batch 2 --steps 10 > gpurun_out/train_n1.json 2> gpurun_out/train_n1.err
echo "train exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/train_n1.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_cmd.csv \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu lists done at $(( $(date +%s) - t0 )) s"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"warp_corr|adaptive_eval|conv5|conv_stem|refine_full" -o gpurun_out/native_full -f python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - t0 )) s"
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
    print('cpu_baseline',b.get('cpu_baseline'), 'gpu eager ref', b.get('gpu_eager_reference'))
except Exception as e: print('bench ERR',e)
try:
    j=json.loads(open("gpurun_out/train_n1.json").read().strip().splitlines()[-1]); print('train N1',j)
except Exception as e: print('train ERR',e)
PY
echo "done at $(( $(date +%s) - t0 )) s"; du -sh gpurun_out
