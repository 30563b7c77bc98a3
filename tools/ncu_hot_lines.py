"""Hottest source lines of a kernel from an ncu report (needs --import-source on and -lineinfo):
   ncu -i REPORT --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_hot_lines.py x.csv [N]
Prints warp-stall samples and executed warp instructions per (file, line), with the dominant stall reasons."""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
total = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue                      # SASS rows (no line number) are folded into their source line's totals by ncu
    d = dict(zip(hdr[4:], r[4:]))
    try:
        samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
    except ValueError:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(Not Issued)" not in k and v.isdigit() and int(v) > 0}
    agg[(cur_file, int(r[0]))] = (samples, inst, stalls, r[1].strip())
    total += samples
print(f"total samples {total}")
print(f"{'file:line':28s} {'samples':>8s} {'%':>6s} {'warp inst':>10s}  top stalls | source")
for (f, l), (s, i, st, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    ts = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f + ':' + str(l):28s} {s:8d} {100.0 * s / max(total, 1):6.2f} {i:10d}  {ts} | {src[:90]}")
