"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
instructions executed + stall samples.  usage: python tools/ncu_lines.py report.ncu-rep [top=40]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
# the dump lists each CUDA line followed by its SASS rows; SASS rows have an Address
agg = collections.OrderedDict()
cur = None
iS, iI = col["# Samples"], col["Instructions Executed"]
first_src = hdr.index("Source")
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    if r[0] != "":      # CUDA line row
        cur = (r[0], r[first_src].strip())
        agg.setdefault(cur, [0, 0, 0])
        try:
            agg[cur][0] += int(r[iI] or 0); agg[cur][1] += int(r[iS] or 0)
        except ValueError:
            pass
tot_i = sum(v[0] for v in agg.values()) or 1; tot_s = sum(v[1] for v in agg.values()) or 1
print(f"total inst {tot_i}  samples {tot_s}")
for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:>5} inst {100*v[0]/tot_i:5.1f}%  samp {100*v[1]/tot_s:5.1f}%  {src[:110]}")
