"""Turn the round-end captures in gpurun_out/ (tools/final_pass.sh) into the text evidence kept under profiles/:
   rNN_ncu_summary.txt   launch list shares + selected --set full metrics (tools/ncu_summary.py)
   rNN_stalls.txt        warp-stall breakdown (smsp__average_warps_issue_stalled_*) and pipe utilisation per kernel
   rNN_lines_<kernel>.txt  hottest CUDA source lines (instructions, stall samples) of the big kernels
   rNN_sass.txt          SASS mnemonic evidence per kernel: bulk async copies (UBLKCP), mbarrier (SYNCS), packed fp32 (FFMA2 ...)
usage: tools/profiles_round.py r02"""
import collections, csv, os, re, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
rep = os.path.join(ROOT, "gpurun_out", "final_prof.ncu-rep")
launches = os.path.join(ROOT, "gpurun_out", "final_launches.csv")
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), launches, rep, os.path.join(out, f"{tag}_ncu_summary.txt"), out], check=True, stdout=subprocess.DEVNULL)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
lines = ["# warp-stall breakdown per kernel (ncu --set full, --clock-control none): average warps stalled per issue-active cycle, by reason", ""]
pipes = ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread"]
seen = set()
for r in rr[2:]:
    k = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
    if k in seen:
        continue
    seen.add(k)
    lines.append(f"## {k}   ({r[h.index('gpu__time_duration.sum')]} us under ncu)")
    st = [(float(r[i].replace(",", "")), h[i]) for i in range(len(h)) if "issue_stalled" in h[i] and h[i].endswith(".ratio") and "not_issued" not in h[i]]
    for v, n in sorted(st, reverse=True)[:8]:
        lines.append("  %6.2f  %s" % (v, n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    for m in pipes:
        if m in h:
            lines.append(f"  {m:62s} {r[h.index(m)]}")
    lines.append("")
open(os.path.join(out, f"{tag}_stalls.txt"), "w").write("\n".join(lines) + "\n")
for kern in ("k_raster_tiles", "k_shade", "k_setup", "k_clip"):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{kern}"], capture_output=True, text=True).stdout
    tmp = f"/tmp/_{kern}.csv"
    open(tmp, "w").write(src)
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp, "40"], capture_output=True, text=True).stdout
    open(os.path.join(out, f"{tag}_lines_{kern}.txt"), "w").write(f"# {kern}: hottest CUDA source lines (share of warp instructions executed, share of stall samples), ncu source page\n" + txt)
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "swraster-viewer_b200", "lib", "libswr_b200.so")], capture_output=True, text=True).stdout
cnt = collections.defaultdict(collections.Counter)
name = None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and name:
        cnt[name][m.group(1)] += 1
want = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "ATOMS", "REDUX", "MATCH", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA"]
lines = ["# SASS mnemonic counts per kernel (cuobjdump -sass lib/libswr_b200.so, sm_100a). UBLKCP = cp.async.bulk (TMA engine, non-tensor form),",
         "# SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed fp32 pairs (sm_100+), ATOMS = shared-memory atomics (64-bit min = ATOMS.CAST.SPIN), MATCH/REDUX = warp match / reduce.",
         "# tcgen05 (UTC*MMA), TMEM (LDTM/STTM) and tensor-map TMA (UTMALDG/UTMASTG) are absent by design: nothing on the path is a dense contraction or a 2-D tile copy.", "",
         f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{w:>7s}" for w in want)]
for k in sorted(cnt, key=lambda k: -sum(cnt[k].values())):
    lines.append(f"{k[:58]:58s} {sum(cnt[k].values()):6d} " + " ".join(f"{cnt[k][w]:7d}" for w in want))
open(os.path.join(out, f"{tag}_sass.txt"), "w").write("\n".join(lines) + "\n")
print("wrote", [f for f in sorted(os.listdir(out)) if f.startswith(tag)])
