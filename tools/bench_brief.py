"""Print the numbers of a bench.py line that matter when iterating on kernels."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    r = lambda x: round(x, 4)  # noqa: E731
    print(path, "ms/step", r(d["ms_per_step"]), {k: r(v) for k, v in d["stage_ms"].items()})
    if "roofline_16m" in d:
        m = d["roofline_16m"]
        print("   16M: ms/step", r(m["ms_per_step"]), {k: r(v) for k, v in m["stage_ms"].items()},
              "frac", r(m["frac"]))
    if "e2e" in d:
        print("   e2e ms/step", r(d["e2e"]["ms_per_step"]))
