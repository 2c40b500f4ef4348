#!/bin/bash
# Run 22: round-1 evidence pass -- GPU tests, sanitizer, smoke, bench (+CPU and eager-GPU baselines), reference arm,
# launch list, ncu --set full of the hot-path kernels.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_conv.py -m gpu -q -x --tb=short -p no:cacheprovider -k "matches_cudnn or transposed or fused_upsample" > gpurun_out/sanitizer_conv.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer_conv.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none \
    -k regex:"warp_corr|adaptive_eval|conv_nhwc" -o gpurun_out/native_full python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
tail -12 gpurun_out/pytest_gpu.log | cut -c1-160; tail -3 gpurun_out/sanitizer_conv.log; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
try:
    b=json.load(open("gpurun_out/bench.json"))
    print('value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'launches/step',b['gpu_launches_per_step'],'roofline',{k:(round(v,3) if isinstance(v,float) else v) for k,v in b['roofline'].items() if k in('achieved','frac','traffic','us_per_launch','best_launch_frac','worst_launch_frac')})
    print('cpu_baseline',b['cpu_baseline']); print('gpu_eager',b['gpu_eager_reference']); print('clocks',b['clocks'])
    for r in b['roofline_detail']: print('   ',r['entry'],r['shape'],round(r['us'],1),round(r['frac'],3))
except Exception as e: print('bench ERR',e)
try:
    r=json.load(open("gpurun_out/bench_ref.json")); print('ref arm value',r['value'],r['cpu_baseline'])
except Exception as e: print('ref ERR',e)
PY
tail -n 2 gpurun_out/bench.err; ls -la gpurun_out/native_full.ncu-rep
