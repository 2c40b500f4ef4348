#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for s in 1 2 3 4 6; do
  timeout 600 python bench.py --steps 24 --warmup 6 --no-cpu-baseline --slots $s > gpurun_out/bench_slots$s.json 2> gpurun_out/bench_slots$s.err
done
tail -5 gpurun_out/pytest_gpu.log
for s in 1 2 3 4 6; do python -c "
import json,sys
b=json.load(open('gpurun_out/bench_slots$s.json')); print('slots',$s,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1))"; done
