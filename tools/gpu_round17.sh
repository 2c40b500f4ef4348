#!/bin/bash
# Run 17: persistent cp.async-pipelined conv kernel -- parity + per-layer micro-benchmark + bench A/B.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_conv.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_conv.log 2>&1
echo "pytest conv exit $?" >> gpurun_out/pytest_conv.log
timeout 600 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err
echo "convbench exit $?" >> gpurun_out/convbench.err
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err
echo "bench native exit $?" >> gpurun_out/bench_native.err
tail -5 gpurun_out/pytest_conv.log; tail -2 gpurun_out/convbench.err
python - <<'PY'
import json
for f in ("bench_native",):
    try:
        b=json.load(open(f"gpurun_out/{f}.json")); print(f,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'launches/step',b['gpu_launches_per_step'],'ka frac',round(b['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e)
try:
    c=json.load(open("gpurun_out/convbench.json")); print(c["sum_cold_us"])
    for r in c["layers"]:
        g=lambda k: (r.get(k) or {}).get("cold") if isinstance(r.get(k),dict) else r.get(k)
        print(r["layer"], r["shape"], 'cudnn', r["cudnn_tf32_us"]["cold"], '| p1 auto/1/2/4/-1:', g("native_p1_mt0_us"), g("native_p1_mt1_us"), g("native_p1_mt2_us"), g("native_p1_mt4_us"), g("native_p1_mt-1_us"), '| warm', (r.get("native_p1_mt0_us") or {}).get("warm"), '| p3 auto', g("native_p3_mt0_us"), '| GB/s', r.get("native_tf32_gbs"))
except Exception as e: print('convbench ERR', e)
PY
tail -n 3 gpurun_out/bench_native.err
