#!/bin/bash
# Round 2 / run 21: serving loop with packed small inputs / one image upload / pooled events: the whole GPU suite (engine
# tests included), the bench line, the e2e probe.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 400 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 200 python tools/e2e_probe.py > gpurun_out/e2e_probe_fused.json 2> gpurun_out/e2e_probe.err
python - <<'PY'
import json
for f in ("bench_default.json",):
    try:
        b=json.load(open("gpurun_out/"+f))
        print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'frac',round(b['roofline']['frac'],3), b['repeats'])
        print(b['latency_single_request'])
        for k in ('value_tf32','cfg3_1600x1184','batch8_640x512'):
            if k in b: print('   ',k,{kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in b[k].items() if kk in ('value','e2e_value','ms_per_step','error')})
    except Exception as e: print(f,'ERR',e, open("gpurun_out/"+f.replace('.json','.err')).read()[-600:])
try: print([ (r['requests'], round(r['device_ms_per_request'],4)) for r in json.loads(open("gpurun_out/e2e_probe_fused.json").read().strip().splitlines()[-1])['rows']])
except Exception as e: print('probe ERR', e, open("gpurun_out/e2e_probe.err").read()[-500:])
PY
echo "done at $(( $(date +%s) - t0 )) s"
