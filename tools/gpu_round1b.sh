#!/bin/bash
# One GPU visit: parity suite, bench lines, ncu launch list and full captures of the gathers.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; cat gpurun_out/bench_1m.json
python bench.py --particles 16000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_16m.json 2> gpurun_out/bench_16m.err; cat gpurun_out/bench_16m.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_density_tile|k_update_tile|k_reorder|k_hash_count' -s 12 -c 4 -o gpurun_out/prof_1m python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
