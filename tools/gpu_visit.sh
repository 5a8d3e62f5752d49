#!/bin/bash
# One GPU visit: parity suite, bench lines, ncu launch list and full captures of the gathers.
# usage: tools/gpu_visit.sh <tag> [quick]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1m_$TAG.json 2> gpurun_out/bench_1m_$TAG.err; cat gpurun_out/bench_1m_$TAG.json; tail -3 gpurun_out/bench_1m_$TAG.err
python bench.py --particles 16000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_16m_$TAG.json 2> gpurun_out/bench_16m_$TAG.err; cat gpurun_out/bench_16m_$TAG.json; tail -3 gpurun_out/bench_16m_$TAG.err
if [ "$2" != "quick" ]; then
ncu --set full --clock-control none --import-source on -k 'regex:k_density_tile|k_update_tile' -s 6 -c 2 -o gpurun_out/prof_1m_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
fi
