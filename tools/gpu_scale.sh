#!/bin/bash
# Scaling visit on one 8-GPU box: N = 8 (both transports), 4, 2 decomposed; N = 1 at 1M (the
# driver's N=1 line) and at 8M (the per-GPU load of the decomposed runs).  usage: tools/gpu_scale.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 N',d['n_gpus'],'particles',d['config']['particles'],'ms',round(d['ms_per_step'],3),'G/s',round(d['value']/1e9,3),{k:round(v,3) for k,v in d['stage_ms'].items()}, 'e2e', round(d.get('e2e',{}).get('value',0)/1e9,3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
"; }
run() { N=$1; EX=$2; OUT=gpurun_out/bench_n${N}_${EX}_$TAG.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange $EX > $OUT 2> ${OUT%.json}.err; echo "rc=$?"; show $OUT $EX; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" ${OUT%.json}.err | tail -3; }
run 8 peer
run 8 nccl
run 4 peer
run 2 peer
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_1m_$TAG.json 2> gpurun_out/bench_n1_1m_$TAG.err; show gpurun_out/bench_n1_1m_$TAG.json single
python bench.py --particles 8000000 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_8m_$TAG.json 2> gpurun_out/bench_n1_8m_$TAG.err; show gpurun_out/bench_n1_8m_$TAG.json single
