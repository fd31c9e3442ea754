"""profiles/make_traffic.py <raw.csv> [tag] -> profiles/r2_traffic.json

Reads an `ncu --set full ... --page raw --csv` export of one bench step and writes the DRAM bytes per launch of the
emit / plan kernels (dram__bytes_read.sum + dram__bytes_write.sum) for bench.py's `roofline.traffic`."""
import csv
import json
import os
import sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def mb(row, name):
    v, u = float(row[ix[name]].replace(",", "")), units[ix[name]]
    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u]


out = {"source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, profiles/%s" % os.path.basename(src),
       "launches": []}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0]
    out["launches"].append({"kernel": name, "dram_read": int(mb(r, "dram__bytes_read.sum")), "dram_write": int(mb(r, "dram__bytes_write.sum")),
                            "duration_us": float(r[ix["gpu__time_duration.sum"]].replace(",", ""))})
k2 = [x for x in out["launches"] if x["kernel"] == "k_emit_nuc"]
if k2:
    out["k_emit_nuc_mean_bytes_per_launch"] = int(sum(x["dram_read"] + x["dram_write"] for x in k2) / len(k2))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_traffic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out)[:600])
