#!/bin/bash
# Multi-GPU visit: NCCL parity test + slab bench.  usage: tools/gpu_multi.sh <ngpus> <tag>
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt
timeout 900 python -m pytest tests/test_slab_gpu.py -m gpu -x -q -k nccl > gpurun_out/pytest_nccl_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nccl_$TAG.log; tail -3 gpurun_out/pytest_nccl_$TAG.log
for PPG in 1000000 8000000; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --particles-per-gpu $PPG > gpurun_out/bench_n${N}_${PPG}_$TAG.json 2> gpurun_out/bench_n${N}_${PPG}_$TAG.err; echo "rc=$?"; cat gpurun_out/bench_n${N}_${PPG}_$TAG.json; tail -5 gpurun_out/bench_n${N}_${PPG}_$TAG.err
done
