"""Turn gpurun_out/{launches,raw}_<tag>.csv (ncu exports) into the committed summaries under
profiles/: <round>_launches_summary.md, <round>_ncu_full_summary.md and traffic.json (per-kernel
DRAM bytes per launch, read by bench.py for roofline.traffic).
usage: python tools/summarise_ncu.py <tag> <round-label> "<what was profiled>" """
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, label, what = sys.argv[1], sys.argv[2], sys.argv[3]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def short(name):
    return name.split("(")[0].replace("void ", "").replace("ds::", "")


rows = list(csv.reader(l for l in open(f"{G}/launches_{tag}.csv") if l.startswith('"')))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[short(r[ik])].append(float(r[iv].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
with open(f"{P}/{label}_launches_summary.md", "w") as f:
    f.write(f"# {label} -- ncu launch list: {what}\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (tools/gpu_profile.sh). "
            "Times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    f.write("| kernel | launches | mean us | share of profiled GPU time |\n|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k[:110]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / tot * 100:.1f}% |\n")
os.replace(f"{G}/launches_{tag}.csv", f"{P}/{label}_launches.csv") if False else None

# one or several raw exports (raw_<tag>.csv, raw_<tag>_<mode>.csv): merged by column NAME
import glob
hdr, units, merged = [], {}, []
for path in sorted(glob.glob(f"{G}/raw_{tag}.csv") + glob.glob(f"{G}/raw_{tag}_*.csv")):
    rr = list(csv.reader(open(path)))
    for h, u in zip(rr[0], rr[1]):
        if h not in units:
            hdr.append(h)
            units[h] = u
    merged += [dict(zip(rr[0], r)) for r in rr[2:]]
assert merged, f"no raw_{tag}*.csv under {G}"
rows = [hdr, [units[h] for h in hdr]] + [[d.get(h, "") for h in hdr] for d in merged]
units = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
data = rows[2:]
names = [short(r[idx["Kernel Name"]]) for r in data]
with open(f"{P}/{label}_ncu_full_summary.md", "w") as f:
    f.write(f"# {label} -- `ncu --set full --clock-control none --import-source on`: {what}\n\n")
    f.write(f"Source report gpurun_out/prof_{tag}.ncu-rep (scratch, not committed); exported with "
            "`ncu -i ... --page raw --csv`.  Per launch.  DRAM writes are small because under ncu's serialised "
            "replay the written windows / panorama rows stay in the 126 MB L2.\n\n")
    f.write("| metric | unit | " + " | ".join(f"`{n[:48]}`" for n in names) + " |\n|---|---|" + "---|" * len(names) + "\n")
    f.write("| grid | | " + " | ".join(r[idx["Grid Size"]] for r in data) + " |\n")
    for w in want:
        if w in idx:
            f.write(f"| {w} | {units[idx[w]]} | " + " | ".join(r[idx[w]] for r in data) + " |\n")
traffic = {}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in data:
    t = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        t += float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]]
    traffic[short(r[idx["Kernel Name"]])] = {"dram_bytes_per_launch": t, "source": f"profiles/{label}_ncu_full_summary.md"}
tp = f"{P}/traffic.json"
old = json.load(open(tp)) if os.path.exists(tp) else {}
old.update(traffic)
json.dump(old, open(tp, "w"), indent=1, sort_keys=True)
print(open(f"{P}/{label}_launches_summary.md").read())
print(json.dumps(traffic, indent=1))
