#!/bin/bash
# Visit: GPU suite, default bench, ncu launch list of the bench command + full captures
# (1M: every kernel of one step; 16M: the two gathers).  usage: tools/gpu_visit3.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; cat gpurun_out/bench_default_$TAG.json; tail -3 gpurun_out/bench_default_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 28 --csv --log-file gpurun_out/launches_1m_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_density_tile|k_update_tile|k_hash_count|k_reorder|k_scatter_ids|k_scan|k_build_groups' -s 21 -c 7 -o gpurun_out/prof_1m_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none -k 'regex:k_density_tile|k_update_tile' -s 6 -c 2 -o gpurun_out/prof_16m_$TAG python bench.py --particles 16000000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full16_$TAG.log 2>&1
ls -la gpurun_out | tail -12
