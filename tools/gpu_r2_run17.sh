#!/bin/bash
# Round 2 / run 17: K-S (fused conv0 -> conv1, exact fp32 FFMA, weights in the constant bank): its parity test, the whole GPU
# suite, convbench (stem row), the bench line with and without it.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 240 python -m pytest tests/test_conv.py -m gpu -x -q --tb=short -p no:cacheprovider -k "stem or refinement" > gpurun_out/pytest_stem.log 2>&1
echo "pytest stem exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_stem.log
tail -4 gpurun_out/pytest_stem.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err
timeout 400 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
PMB200_REFINE=0 timeout 300 python bench.py --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_stem_off.json 2> gpurun_out/bench_stem_off.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/convbench.json")); print('stem', j['stem']); print('refinement', j['refinement'])
except Exception as e: print('convbench ERR',e, open("gpurun_out/convbench.err").read()[-500:])
for f in ("bench_default.json","bench_stem_off.json"):
    try:
        b=json.load(open("gpurun_out/"+f))
        print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'frac',round(b['roofline']['frac'],3),'traffic',b['roofline']['traffic'], b['native_kernels_per_step'])
        for k in ('value_tf32','cfg3_1600x1184','batch8_640x512'):
            if k in b: print('   ',k,{kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in b[k].items() if kk in ('value','e2e_value','ms_per_step','error')})
    except Exception as e: print(f,'ERR',e, open("gpurun_out/"+f.replace('.json','.err')).read()[-600:])
PY
echo "done at $(( $(date +%s) - t0 )) s"
