"""profiles/make_traffic.py <raw.csv> [fused_raw.csv one_launch_raw.csv] -> profiles/r2_traffic.json

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


def launches_of(path):
    rr = list(csv.reader(open(path)))
    h, u = rr[0], rr[1]
    jx = {k: i for i, k in enumerate(h)}
    res = []
    for r in rr[2:]:
        f = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        res.append((r[jx["Kernel Name"]].split("(")[0], float(r[jx["dram__bytes_read.sum"]].replace(",", "")) * f[u[jx["dram__bytes_read.sum"]]]))
    return res


if len(sys.argv) > 3:                                # DRAM reads of the fused forms against the launches they replace
    fused, one = launches_of(sys.argv[2]), launches_of(sys.argv[3])
    rd = lambda ls, name: [b for n, b in ls if n == name]   # noqa: E731
    k2 = sorted(rd(fused, "k_emit_nuc"))              # CDS launches (small), exon launches (large)
    k2_cds, k2_exon = k2[0], k2[-1]
    k3, k23, multi = rd(fused, "k_emit_prot")[0], rd(fused, "k_emit_nuc_prot")[0], rd(one, "k_emit_multi")[0]
    out["k23"] = {"dram_read_MB": round(k23 / 1e6, 1), "k2_plus_k3_dram_read_MB": round((k2_cds + k3) / 1e6, 1), "source": os.path.basename(sys.argv[2])}
    out["multi"] = {"dram_read_MB": round(multi / 1e6, 1), "three_launches_dram_read_MB": round((k2_cds + k2_exon + k3) / 1e6, 1), "source": os.path.basename(sys.argv[3])}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_traffic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out)[:600])
