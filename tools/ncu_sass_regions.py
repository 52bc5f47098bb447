"""From an `ncu --page source --print-source cuda,sass --csv` dump: SASS in address order, consecutive instructions with the
same execution count merged into regions (count, #instr, warp-instr total, stall samples, first/last opcode)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
ins = {}
for r in rows:
    if len(r) > 8 and r[2].startswith("0x"):
        ins[int(r[2], 16)] = (r[3].strip(), int(r[6]), int(r[7]), int(r[8]))
addrs = sorted(ins)
tot = sum(v[2] for v in ins.values()); ts = sum(v[1] for v in ins.values())
print(f"instructions {len(addrs)}  executed warp-instr {tot:,} samples {ts}")
base = addrs[0]
reg = []
for a in addrs:
    op, smp, n, tn = ins[a]
    if reg and reg[-1]["n"] == n:
        g = reg[-1]; g["k"] += 1; g["smp"] += smp; g["tn"] += tn; g["last"] = op; g["ops"].append(op.split()[0] if not op.startswith("@") else op.split()[1])
    else:
        reg.append({"a": a - base, "n": n, "k": 1, "smp": smp, "tn": tn, "first": op, "last": op, "ops": [op.split()[0] if not op.startswith("@") else op.split()[1]]})
for g in reg:
    w = g["n"] * g["k"]
    if 100.0 * w / tot >= minshare or 100.0 * g["smp"] / ts >= minshare:
        import collections
        c = collections.Counter(o.split(".")[0] for o in g["ops"]).most_common(6)
        print(f"+{g['a']:6x} x{g['n']:>10,} k={g['k']:4d} inst {100.0*w/tot:5.1f}% stall {100.0*g['smp']/ts:5.1f}% thr/inst {g['tn']/max(w,1):5.1f}  {c}  [{g['first'][:40]}]")
