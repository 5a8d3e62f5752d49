#!/bin/bash
# Kernel-tuning visit: bench the default library and every variant_*.so (stage times only).
mkdir -p gpurun_out
for lib in watercube_b200/csrc/libwc_sph.so watercube_b200/csrc/variant_*.so; do
  echo "== $lib"
  WC_SPH_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > /tmp/v.json 2>/tmp/v.err || tail -3 /tmp/v.err
  python tools/bench_brief.py /tmp/v.json
done
