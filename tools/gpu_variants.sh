#!/bin/bash
# Kernel-tuning visit: bench every variant_*.so (stage times only).
mkdir -p gpurun_out
for lib in watercube_b200/csrc/libwc_sph.so watercube_b200/csrc/variant_*.so; do
  for P in 1000000 16000000; do
    echo "== $lib $P"
    WC_SPH_LIB=$PWD/$lib python bench.py --particles $P --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['stage_ms'].items()})"
  done
done
