"""Summarise the text exports of an ncu --set full capture (details / raw / source pages) into profiles/:
   python tools/prof/summarize.py gpurun_out/prof_r2_fold profiles/r2_ncu_fold_kernel2_contra_L89"""
import csv
import sys

stem, out = sys.argv[1], sys.argv[2]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
rows = list(csv.reader(open(stem + "_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
idx = {h: i for i, h in enumerate(hdr)}
lines = [f"# selected raw counters of {stem.split('/')[-1]} (ncu --set full --clock-control none), kernel: {vals[idx['Kernel Name']]}"]
for k in KEYS:
    if k in idx:
        lines.append(f"{k} {vals[idx[k]]} {units[idx[k]]}")
stall = [(h, float(vals[i] or 0)) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
lines.append("# warp stall reasons, warps per issue-active cycle (smsp__average_warps_issue_stalled_*_per_issue_active.ratio)")
for h, v in sorted(stall, key=lambda x: -x[1])[:12]:
    lines.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
open(out + "_raw_selected.txt", "w").write("\n".join(lines) + "\n")
# details page as is (small)
open(out + "_details.csv", "w").write(open(stem + "_details.csv").read())
# source page: top SASS instructions by stall samples + totals per stall reason
rows = list(csv.reader(open(stem + "_source.csv")))
h2 = rows[1]
ix = {h: i for i, h in enumerate(h2)}
st = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(h2):
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    data.append((n, r[ix["Source"]], {s: int(r[ix[s]] or 0) for s in st}, int(r[ix["Instructions Executed"]] or 0)))
tot = sum(d[0] for d in data)
agg = sorted(((s, sum(d[2][s] for d in data)) for s in st), key=lambda x: -x[1])
L = [f"# source page of {stem.split('/')[-1]}: {len(data)} SASS instructions, {tot} warp-stall samples, {sum(d[3] for d in data)} warp-instructions executed",
     "# samples per stall reason: " + ", ".join(f"{k[6:]}={v} ({100 * v / max(tot, 1):.1f}%)" for k, v in agg[:9]),
     "# top instructions by samples: samples | executed | dominant stall | SASS"]
for n, src, sd, ex in sorted(data, key=lambda d: -d[0])[:40]:
    dom = max(sd.items(), key=lambda x: x[1])
    L.append(f"{n:8d} {ex:10d} {dom[0][6:]:14s} {src.strip()[:90]}")
open(out + "_source_hotspots.txt", "w").write("\n".join(L) + "\n")
print("\n".join(lines[:40]))
print(L[1])
