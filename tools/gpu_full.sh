#!/bin/bash
# Full evidence pass: GPU tests, smoke, bench (+CPU baseline), reference arm, launch list, ncu --set full of the native kernels.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"warp_corr|adaptive_eval|aggregate_score|init_propagate|init_only|offset_corr|relative_projection|upsample2x" \
    -o gpurun_out/native_full python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; head -c 500 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err; head -c 300 gpurun_out/bench_ref.json
