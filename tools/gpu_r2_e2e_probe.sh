mkdir -p gpurun_out
timeout 200 python tools/e2e_probe.py > gpurun_out/e2e_probe_fused.json 2> gpurun_out/e2e_probe.err
PMB200_REFINE=0 timeout 200 python tools/e2e_probe.py > gpurun_out/e2e_probe_unfused.json 2>> gpurun_out/e2e_probe.err
tail -1 gpurun_out/e2e_probe_fused.json; tail -1 gpurun_out/e2e_probe_unfused.json; tail -3 gpurun_out/e2e_probe.err
