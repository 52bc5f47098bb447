"""Turn an ncu launch list (csv) and a full .ncu-rep into the short text summaries kept under profiles/."""
import collections, csv, subprocess, sys
launch_csv, rep, out = sys.argv[1], sys.argv[2], sys.argv[3]
import re
def kname(full):
    """`void k_shade<1>(ShadeParams)` -> `k_shade`: templated kernels are printed with their return type and arguments"""
    n = full.split("(")[0].strip()
    n = n[5:] if n.startswith("void ") else n
    return re.sub(r"<.*>$", "", n)
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(kname(r[ki])[:60], []).append(float(r[vi].replace(",", "")))
lines = ["# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)", ""]
own = {k: v for k, v in agg.items() if k.startswith("k_")}
per_frame = sum(sum(v) / len(v) for v in own.values())
lines.append(f"{'kernel':24s} {'launches':>8s} {'avg us':>10s} {'share of frame':>15s}")
for k, v in agg.items():
    avg = sum(v) / len(v) / 1e3
    share = f"{100 * (sum(v) / len(v)) / per_frame:5.1f} %" if k in own else "(not ours)"
    lines.append(f"{k:24s} {len(v):8d} {avg:10.1f} {share:>15s}")
lines += ["", "# ncu --set full, selected metrics per kernel", ""]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rr[2:]:
    lines.append("## " + r[h.index("Kernel Name")])
    for w in want:
        if w in h:
            lines.append(f"  {w:66s} {r[h.index(w)]:>18s} {rr[1][h.index(w)]}")
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 4:  # also: per-kernel DRAM traffic and warp-instruction counts per launch, as bench.py reads them
    import json, os
    traffic, winst = {}, {}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rr[2:]:
        k = kname(r[h.index("Kernel Name")])
        rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        traffic[k] = int(float(r[rd].replace(",", "")) * scale[rr[1][rd]] + float(r[wr].replace(",", "")) * scale[rr[1][wr]])
        winst[k] = int(float(r[h.index("smsp__inst_executed.sum")].replace(",", "")))
    json.dump(traffic, open(os.path.join(sys.argv[4], "r02_traffic.json"), "w"), indent=1)
    json.dump(winst, open(os.path.join(sys.argv[4], "r02_warp_inst.json"), "w"), indent=1)
print("\n".join(lines))
