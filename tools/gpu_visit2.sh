#!/bin/bash
# Visit: full GPU suite, default bench (with cpu baseline + reference arm), 16M, uniform-box sweep,
# ncu launch list + full capture.  usage: tools/gpu_visit2.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; cat gpurun_out/bench_default_$TAG.json; tail -3 gpurun_out/bench_default_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
python bench.py --particles 16000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_16m_$TAG.json 2> gpurun_out/bench_16m_$TAG.err; cat gpurun_out/bench_16m_$TAG.json; tail -3 gpurun_out/bench_16m_$TAG.err
for NB in 30 50 100 200; do
python bench.py --workload uniform_box --particles 8000000 --neighbours $NB --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_uniform_nb${NB}_$TAG.json 2> gpurun_out/bench_uniform_nb${NB}_$TAG.err; cat gpurun_out/bench_uniform_nb${NB}_$TAG.json; tail -3 gpurun_out/bench_uniform_nb${NB}_$TAG.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_1m_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_density_tile|k_update_tile|k_hash_count|k_reorder|k_scatter_ids|k_scan|k_build_groups' -s 28 -c 7 -o gpurun_out/prof_1m_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none -k 'regex:k_density_tile|k_update_tile' -s 6 -c 2 -o gpurun_out/prof_16m_$TAG python bench.py --particles 16000000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full16_$TAG.log 2>&1
