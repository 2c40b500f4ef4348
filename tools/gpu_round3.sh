#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
PMB200_WARP_CORR_V1=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v1kernel.json 2> gpurun_out/bench_v1kernel.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"warp_corr|adaptive_eval|aggregate_score" \
    -o gpurun_out/native_full python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line -p no:cacheprovider \
    -k "warp_corr or fused_heads or init_propagate or adaptive_eval" > gpurun_out/sanitizer.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; head -c 1200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -4 gpurun_out/sanitizer.log
