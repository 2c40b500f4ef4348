#!/bin/bash
# multi-GPU check of the bench contract: launched exactly as the driver does
set -u
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?" >> gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "ref N=$N exit $?" >> gpurun_out/bench_ref_n$N.err
head -c 700 gpurun_out/bench_n$N.json; echo; tail -4 gpurun_out/bench_n$N.err; head -c 400 gpurun_out/bench_ref_n$N.json; tail -2 gpurun_out/bench_ref_n$N.err
