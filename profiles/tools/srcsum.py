"""scratch: per-source-line instruction / stall totals from `ncu --page source --print-source cuda,sass --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; out = []; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and r[2] == "-":
        iI = hdr.index("Instructions Executed"); iS = hdr.index("Warp Stall Sampling (All Samples)"); iT = hdr.index("Thread Instructions Executed")
        out.append((int(r[iI]), int(r[iS]), int(r[iT]), cur, int(r[0]), r[1].strip()[:110]))
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total warp instr %d, stall samples %d" % (tot, tots))
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% stall  lanes %4.1f  %s:%d  %s" % (100.0 * o[0] / tot, 100.0 * o[1] / max(tots, 1), o[2] / max(o[0], 1), o[3], o[4], o[5]))
