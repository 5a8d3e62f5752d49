#!/bin/bash
# BASELINE configs[4]: uniform random box, 8M particles, smoothing-length sweep (device-timed stage split).
for NB in 30 50 100 200; do
  python bench.py --workload uniform_box --particles 8000000 --neighbours $NB --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null > /tmp/sweep_$NB.json
  python - "$NB" <<'PY'
import sys, json
nb = sys.argv[1]
d = json.loads(open(f"/tmp/sweep_{nb}.json").read().strip().splitlines()[-1])
print("nb", nb, "G", d["config"]["grid_res"], "ms", round(d["ms_per_step"], 3), "G/s", round(d["value"] / 1e9, 3),
      {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
done
