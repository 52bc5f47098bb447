"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump: per CUDA source line, instructions and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
data, fname = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 8 and r[0] not in ("", "Line No") and r[2] == "-":
        try:
            data.append((int(r[7]), int(r[6]), fname, r[0], r[1]))
        except ValueError:
            pass
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total warp-instructions {ti:,}  samples {ts:,}")
for n, s, f, l, src in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{100*n/ti:5.1f}% inst {100*s/max(ts,1):5.1f}% stall  {f}:{l:>4}  {src.strip()[:110]}")
