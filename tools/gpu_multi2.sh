#!/bin/bash
# usage: tools/gpu_multi2.sh <ngpus> <tag> [ppg]   -- slab tests + bench with both transports
N=${1:-2}; TAG=${2:-x}; PPG=${3:-8000000}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q > gpurun_out/pytest_slab_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_slab_$TAG.log; tail -15 gpurun_out/pytest_slab_$TAG.log
for EX in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --particles-per-gpu $PPG --exchange $EX > gpurun_out/bench_n${N}_${EX}_$TAG.json 2> gpurun_out/bench_n${N}_${EX}_$TAG.err; echo "rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n${N}_${EX}_$TAG.json').read().strip().splitlines()[-1])
print('$EX N',d['n_gpus'],'particles',d['config']['particles'],'ms',round(d['ms_per_step'],3),'G/s',round(d['value']/1e9,3),{k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e', round(d.get('e2e',{}).get('value',0)/1e9,3))
"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/bench_n${N}_${EX}_$TAG.err | tail -5
done
