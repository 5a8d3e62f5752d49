#!/bin/bash
# One parameterised GPU visit (replaces the per-session gpu_*.sh wrappers).
#   tools/gpu_visit.sh <tag> [steps...]     steps: tests | tests:<pytest -k expr> | smoke | bench | quick | bench16 |
#       ref | launches | ncu1m | ncu16m | sanitize | variants | sweep | slabtests | scale:<N>[:nccl] | refN:<N>
# (scale:<N> / slabtests need `gpurun --gpus N`)
# Everything lands in gpurun_out/<name>_<tag>.*
TAG=${1:-x}; shift
STEPS=${@:-tests smoke bench}
mkdir -p gpurun_out
for S in $STEPS; do
  case $S in
    tests)   timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -15 gpurun_out/pytest_gpu_$TAG.log;;
    tests:*) timeout 2400 python -m pytest tests -m gpu -x -q -k "${S#tests:}" > gpurun_out/pytest_k_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k_$TAG.log; tail -25 gpurun_out/pytest_k_$TAG.log;;
    smoke)   python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log;;
    bench)   python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; cat gpurun_out/bench_default_$TAG.json; tail -3 gpurun_out/bench_default_$TAG.err;;
    quick)   python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick_$TAG.json 2> gpurun_out/bench_quick_$TAG.err; python tools/bench_brief.py gpurun_out/bench_quick_$TAG.json; tail -3 gpurun_out/bench_quick_$TAG.err;;
    bench16) python bench.py --particles 16000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_16m_$TAG.json 2> gpurun_out/bench_16m_$TAG.err; cat gpurun_out/bench_16m_$TAG.json; tail -3 gpurun_out/bench_16m_$TAG.err;;
    ref)     python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1m_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-16m > gpurun_out/ncu_launch_$TAG.log 2>&1; tail -12 gpurun_out/launches_1m_$TAG.csv;;
    ncu1m)   ncu --set full --clock-control none --import-source on -k 'regex:k_density|k_update|k_reorder|k_hash_count' -s 12 -c 6 -o gpurun_out/prof_1m_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-16m > gpurun_out/ncu_full_$TAG.log 2>&1; ls -la gpurun_out/prof_1m_$TAG.ncu-rep;;
    ncu16m)  ncu --set full --clock-control none --import-source on -k 'regex:k_density|k_update' -s 4 -c 2 -o gpurun_out/prof_16m_$TAG -f python bench.py --particles 16000000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full16_$TAG.log 2>&1; ls -la gpurun_out/prof_16m_$TAG.ncu-rep;;
    sanitize) bash tools/sanitize.sh $TAG;;
    variants) bash tools/gpu_variants.sh 2>&1 | tee gpurun_out/variants_$TAG.log;;
    sweep)   # BASELINE configs[4]: uniform random box, 8M particles, smoothing-length sweep
             for NB in 30 50 100 200; do python bench.py --workload uniform_box --particles 8000000 --neighbours $NB --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_${NB}_$TAG.json 2>/dev/null; echo "nb=$NB"; python tools/bench_brief.py gpurun_out/sweep_${NB}_$TAG.json; done;;
    slabtests) timeout 1200 python -m pytest tests/test_slab_gpu.py -m gpu -x -q > gpurun_out/pytest_slab_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_slab_$TAG.log; tail -5 gpurun_out/pytest_slab_$TAG.log;;
    scale:*) IFS=: read -r _ N EX <<< "$S"; EX=${EX:-peer}; OUT=gpurun_out/bench_n${N}_${EX}_$TAG.json
             timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange $EX > $OUT 2> ${OUT%.json}.err; echo "rc=$?"
             python tools/bench_brief.py $OUT; python -c "
import json; d=json.loads(open('$OUT').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('n_gpus','slab_parity','efficiency_vs_weak_baseline','gpu_launches_per_step_per_rank')}, 'e2e', d.get('e2e',{}).get('ms_per_step'))"
             grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" ${OUT%.json}.err | tail -5;;
    refN:*)  N=${S#refN:}; timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n${N}_$TAG.json 2> gpurun_out/bench_ref_n${N}_$TAG.err; tail -1 gpurun_out/bench_ref_n${N}_$TAG.json;;
    *) echo "unknown step $S";;
  esac
done
