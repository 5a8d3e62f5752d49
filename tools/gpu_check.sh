#!/bin/bash
# Quick visit: full GPU suite, smoke, default bench (+cpu baseline), reference arm, 16M bench.
# usage: tools/gpu_check.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; cat gpurun_out/bench_default_$TAG.json; tail -3 gpurun_out/bench_default_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
python bench.py --particles 16000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_16m_$TAG.json 2> gpurun_out/bench_16m_$TAG.err; cat gpurun_out/bench_16m_$TAG.json; tail -3 gpurun_out/bench_16m_$TAG.err
