#!/bin/bash
# Run 25 (first GPU call of the next round): time everything that was changed with the GPU budget spent
# (K-A 32-bit tap indices + launch bounds, K-D column staging, bench clock sampler), sweep the knobs that were built but
# left off (PMB200_KA_MINB=5, PMB200_KA_DC_VW=16), and refresh the ncu evidence.  Every step writes its own file.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/bench.err
# requests coalesced into one batch vs independent slots (run 9, older kernels: batch 8 gave 1,058 maps/s against 871 with 3 slots)
for cfg in "2 2" "4 1" "4 2" "8 1"; do
  set -- $cfg
  timeout 200 python bench.py --batch $1 --slots $2 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_b$1_s$2.json 2> gpurun_out/bench_b$1_s$2.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tools/kbench.py > gpurun_out/kbench.json 2> gpurun_out/kbench.err          # full K-A / K-B sweeps
timeout 600 python tools/convbench.py > gpurun_out/convbench.json 2> gpurun_out/convbench.err  # per-layer native vs cuDNN
PMB200_CONV_T=1 timeout 600 python tools/convbench.py > gpurun_out/convbench_transposed.json 2> gpurun_out/convbench_transposed.err
timeout 120 python tools/geobench.py > gpurun_out/geobench.json 2> gpurun_out/geobench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_forward.py > gpurun_out/launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"warp_corr|adaptive_eval|conv_nhwc" -o /tmp/native_full python tools/profile_forward.py > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/native_full.ncu-rep --page raw --csv > gpurun_out/native_full_raw.csv 2>/dev/null
python - <<'PY'
import json
try:
    b=json.load(open("gpurun_out/bench.json"))
    print('value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'ms',round(b['ms_per_step'],3),'roofline frac',round(b['roofline']['frac'],3),'clocks',b['clocks'])
except Exception as e: print('bench ERR',e)
for tag in ("b2_s2","b4_s1","b4_s2","b8_s1"):
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
