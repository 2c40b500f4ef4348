#!/bin/bash
# Round 2 / run 4 (8 GPUs): BASELINE config 5 -- the full training step (forward, reference loss, native backward, ONE
# flat-buffer NCCL all-reduce over gradient views, Adam) at N = 1 / 2 / 4 / 8, 2 reference views per GPU -- and the
# inference bench at N = 1 and N = 8 with the NUMA binding, for the e2e scaling efficiency.
set -u
mkdir -p gpurun_out
t0=$(date +%s)
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > gpurun_out/build.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 200 python tools/train_step.py --batch 2 --steps 10 > gpurun_out/train_n1.json 2> gpurun_out/train_n1.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      tools/train_step.py --batch 2 --steps 10 > gpurun_out/train_n$n.json 2> gpurun_out/train_n$n.err
  echo "train N=$n exit $? at $(( $(date +%s) - t0 )) s" >> gpurun_out/train_n$n.err
done
timeout 300 python bench.py --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 8 --no-sub --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        j=json.loads(open(f"gpurun_out/train_n{n}.json").read().strip().splitlines()[-1])
        print('train N',n,{k:(round(v,3) if isinstance(v,float) else v) for k,v in j.items() if k in ('value','ms_per_step','allreduce_ms_per_step','params_without_grad','grads_not_views_of_the_flat_buffer','collectives_per_step','loss')})
    except Exception as e: print('train',n,'ERR',e)
for n in (1,8):
    try:
        b=json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
        print('bench N',n,'value',round(b['value'],1),'e2e',round(b['e2e']['value'],1),'numa',b['config']['numa'])
    except Exception as e: print('bench',n,'ERR',e)
PY
echo "done at $(( $(date +%s) - t0 )) s"
