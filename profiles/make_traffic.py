"""profiles/traffic_r02.json from an `ncu --page raw --csv` dump of the two store kernels (one launch each is used: the
first of every kernel name). Usage: python profiles/make_traffic.py raw.csv out.json workload replicas source-note"""
import csv, json, re, sys
raw, out, workload, replicas, note = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
kern = {}
for r in rows[2:]:
    m = re.search(r"(k_\w+)", r[col["Kernel Name"]])
    if not m or m.group(1) in kern:
        continue
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    dur = float(r[col["gpu__time_duration.sum"]])
    du = units[col["gpu__time_duration.sum"]]
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(du, 1.0)
    kern[m.group(1)] = {"dram_bytes_per_launch": int(rd + wr), "read_MB": round(rd / 1e6, 3), "write_MB": round(wr / 1e6, 3),
                        "duration_us": round(dur_us, 3), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                        "registers": int(float(r[col["launch__registers_per_thread"]])),
                        "dram_cycles_active_pct": float(r[col["dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]])
                        if "dram__cycles_active.avg.pct_of_peak_sustained_elapsed" in col else None,
                        "achieved_occupancy_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]])}
json.dump({"workload": workload, "replicas": replicas, "source": note, "kernels": kern}, open(out, "w"), indent=1)
print(json.dumps(kern, indent=1))
