#!/bin/bash
# Round 2 / run 11: K-D5h with the resident filter + next-halo-first producer, dispatch lists set from run 10: parity tests,
# the whole GPU suite, convbench, the bench line with the lists (default) next to PMB200_TC5H=0 (A/B).
set -u
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 240 python -m pytest tests/test_conv.py -m gpu -x -q --tb=short -p no:cacheprovider -k "tc5" > gpurun_out/pytest_tc5h.log 2>&1
echo "pytest tc5 exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_tc5h.log
tail -4 gpurun_out/pytest_tc5h.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
PMB200_TC5H=0 timeout 300 python bench.py --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_tc5h_off.json 2> gpurun_out/bench_tc5h_off.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/convbench.json"))
    g=lambda r,k: (r.get(k) or {}).get('cold') if isinstance(r.get(k),dict) else r.get(k)
    for r in j['layers']: print('  conv',r['layer'],r['shape'],'p3',r['native_p3_mt0_us']['cold'],'tc5',g(r,'tc5_3xtf32_us'),'tc5h',g(r,'tc5h_3xtf32_us'))
except Exception as e: print('convbench ERR',e)
for f in ("bench_default.json","bench_tc5h_off.json"):
    try:
        b=json.load(open("gpurun_out/"+f))
        print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'frac',b['roofline']['frac'], b['native_kernels_per_step'])
        for k in ('value_tf32','cfg3_1600x1184','batch8_640x512'):
            if k in b: print('   ',k,b[k])
    except Exception as e: print(f,'ERR',e)
PY
echo "done at $(( $(date +%s) - t0 )) s"
