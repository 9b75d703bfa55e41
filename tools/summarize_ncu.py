"""Turns ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.
   python tools/summarize_ncu.py <round tag> <full .ncu-rep> <launch list csv> [reads in the profiled run]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, launches = sys.argv[1], sys.argv[2], sys.argv[3]
n_reads = int(sys.argv[4]) if len(sys.argv) > 4 else 2_000_000
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

FAMILY = [("seed_kernel", "seed"), ("fill_kernel", "fill"), ("pack_reads", "fill"), ("align_init", "align_screen"), ("align_screen", "align_screen"),
          ("align_walk", "align_walk"), ("align_finish", "align_finish"), ("align_emit", "align_emit"), ("project_", "project"), ("peek_kernel", "scalars"), ("poke_kernel", "scalars"), ("zero_kernel", "scalars"),
          ("chunk_", "chunk glue"),
          ("sketch_kernel", "sketch(index)")]


def family(name):
    for key, fam in FAMILY:
        if key in name:
            return fam
    return "cub/other"


# ---- launch list -> per-family share of the step
rows = [l for l in open(launches) if not l.startswith("==")]
agg, order = {}, []
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    f = family(r["Kernel Name"])
    agg.setdefault(f, [0, 0.0])
    agg[f][0] += 1
    agg[f][1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for k, v in agg.items() if k != "sketch(index)")
lines = ["| kernel family | launches | total ms | share of step |", "|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    share = "%.1f %%" % (100 * v[1] / tot) if k != "sketch(index)" else "(index build)"
    lines.append("| %s | %d | %.3f | %s |" % (k, v[0], v[1] / 1e6, share))
share_md = "\n".join(lines)

# ---- full capture -> key metrics per kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rd = list(csv.reader(raw.splitlines()))
hdr, units = rd[0], rd[1]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
        ("launch__registers_per_thread", "registers/thread"), ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
seen, md, traffic = set(), [], {}
for r in rd[2:]:
    name = r[ix["Kernel Name"]]
    fam = family(name)
    if name in seen:                 # first launch of every distinct kernel (template instances count separately)
        continue
    seen.add(name)
    md.append("### %s  (`%s`)" % (fam, name[:90]))
    md.append("| metric | value |")
    md.append("|---|---|")
    for key, label in want:
        if key in ix:
            md.append("| %s | %s %s |" % (label, r[ix[key]], units[ix[key]]))
    stalls = []
    for h in hdr:
        if "issue_stalled" in h and "pcsamp" in h and not h.endswith("not_issued"):
            try:
                stalls.append((float(r[ix[h]].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    md.append("| top stall reasons (samples) | %s |" % ", ".join("%s %d" % (n, v) for v, n in sorted(stalls, reverse=True)[:5]))
    md.append("")

    def num(key):
        v, u = float(r[ix[key]].replace(",", "")), units[ix[key]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    try:   # a family's traffic = the sum over its distinct kernels of one batch (e.g. seed = prescreen + queued pass)
        t = traffic.setdefault(fam, {"dram_bytes_per_read": 0.0, "reads_in_profiled_launch": n_reads, "source": os.path.basename(rep), "kernels": []})
        t["dram_bytes_per_read"] += (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / n_reads
        t["warp_inst_per_read"] = t.get("warp_inst_per_read", 0.0) + float(r[ix["smsp__inst_executed.sum"]].replace(",", "")) / n_reads
        t["kernels"].append(name[:60])
    except Exception:
        pass

open(os.path.join(out_dir, "%s_ncu_summary.md" % tag), "w").write(
    "# %s — ncu summary (B200, `--set full --clock-control none`, %d reads per launch)\n\n"
    "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
    "## Launch list of one bench step (`gpu__time_duration.sum`)\n\n%s\n\n## Key metrics per kernel\n\n%s\n" % (tag, n_reads, share_md, "\n".join(md)))
json.dump(traffic, open(os.path.join(out_dir, "kernel_traffic.json"), "w"), indent=1)
print(share_md)
