#!/bin/bash
# Round 2 / run 8: why are the tcgen05 convs 3-4x off their model?  (a) tcgen05.mma cost vs N / accumulators / layout
# (tools/umma_rate.cu), (b) ncu --set full with source-level stall samples of K-D5h and K-D5 on conv6/7 and conv3/4.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate.cu > gpurun_out/umma_build.log 2>&1
timeout 120 /tmp/umma_rate > gpurun_out/umma_rate.json 2> gpurun_out/umma_rate.err
echo "umma_rate exit $? at $(( $(date +%s) - t0 )) s"
head -c 3000 gpurun_out/umma_rate.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv5 -s 4 -c 2 -f -o gpurun_out/conv5_conv67 \
    python tools/conv5_probe.py 5 32 32 3 1 1 1 128 160 > gpurun_out/ncu_conv67.log 2>&1
echo "ncu conv6/7 exit $? at $(( $(date +%s) - t0 )) s"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv5h -s 2 -c 1 -f -o gpurun_out/conv5h_conv34 \
    python tools/conv5_probe.py 5 16 16 3 1 1 1 256 320 > gpurun_out/ncu_conv34.log 2>&1
echo "ncu conv3/4 exit $? at $(( $(date +%s) - t0 )) s"
ls -la gpurun_out/*.ncu-rep
echo "done at $(( $(date +%s) - t0 )) s"
