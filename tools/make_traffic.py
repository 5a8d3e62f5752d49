"""profiles/traffic.json from an `ncu --set full` report (read here, no GPU):
    python tools/make_traffic.py <rep.ncu-rep> <particles> [<rep2> <particles2> ...]
Per kernel and particle count: DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum) and warp instructions per launch (smsp__inst_executed.sum) -- the
figures bench.py puts into roofline.traffic and issue_roofline."""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
out_path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
args = sys.argv[1:]
for rep, n in zip(args[0::2], args[1::2]):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[idx["Kernel Name"]])
        name = re.split(r"[<(]", name)[0].split("::")[-1]
        b = sum(float(r[idx[m]]) * UNIT[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out[f"{name}@{int(n)}"] = {
            "dram_bytes": b, "warp_inst": float(r[idx["smsp__inst_executed.sum"]]),
            "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]])
            * TIME_US[units[idx["gpu__time_duration.sum"]].replace("second", "s").replace("usecond", "us")
                      .replace("msecond", "ms").replace("nsecond", "ns")],
            "source": os.path.basename(rep)}
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
