"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/make_summary.py <launches.csv> <full.ncu-rep> <tag> [frames per launch of the full capture, default 8]

<launches.csv>: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list of bench.py
<full.ncu-rep>: `ncu --set full --clock-control none --import-source on` capture of the top kernels, or the CSV of its
                raw page (`ncu -i REP --page raw --csv`; the report itself can exceed what travels back from the GPU box)
Writes profiles/<tag>_launches.md, profiles/<tag>_kernels.md and profiles/<tag>_traffic.json
(per-launch DRAM traffic of each captured kernel; bench.py reads it for roofline.traffic).
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def short(name):
    n = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    return n.strip()


def launches(path, tag):
    rows = [l for l in open(path) if not l.startswith("==")]
    tot = collections.OrderedDict()
    for row in csv.DictReader(rows):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        k = short(row["Kernel Name"])
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    s = sum(v[1] for v in tot.values())
    out = [f"# {tag}: launch list of `bench.py` under ncu (cold-cache, serialised: compare SHARES)\n",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / s:.1f} % |")
    out.append(f"\ntotal {s:.1f} us over {sum(v[0] for v in tot.values())} launches")
    open(os.path.join(HERE, f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")


KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "regs/thread"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (inst) %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (inst) %")]


def kernels(rep, tag, frames=8):
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    out = [f"# {tag}: `ncu --set full --clock-control none` of the top kernels (one launch each, {frames} 4K frames per launch: `tools/prof_step.py {frames}`)\n"]
    traffic = {}
    stats = {}
    seen = set()
    for row in r[2:]:
        name = short(row[hdr.index("Kernel Name")])
        if name in seen:
            continue
        seen.add(name)
        out.append(f"## `{name}`  grid {row[hdr.index('Grid Size')]} block {row[hdr.index('Block Size')]}\n")
        out.append("| metric | value |")
        out.append("|---|---|")
        rd = wr = None
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"| {label} (`{k}`) | {row[i]} {units[i]} |")
                if k == "dram__bytes_read.sum":
                    rd = float(row[i]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(units[i], 1)
                if k == "dram__bytes_write.sum":
                    wr = float(row[i]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(units[i], 1)
        if rd is not None and wr is not None:
            traffic[name] = rd + wr
        st = {}
        for k, label in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                         ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
                         ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
                         ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                         ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
                         ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
                         ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefronts_pct")):
            if k in hdr and row[hdr.index(k)] not in ("", "n/a"):
                st[label] = float(row[hdr.index(k)])
        stats[name] = st
        out.append("")
    open(os.path.join(HERE, f"{tag}_kernels.md"), "w").write("\n".join(out) + "\n")
    traffic["_frames_per_launch"] = frames
    json.dump(traffic, open(os.path.join(HERE, f"{tag}_traffic.json"), "w"), indent=1, sort_keys=True)
    json.dump(stats, open(os.path.join(HERE, f"{tag}_kernel_stats.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[3])
    kernels(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 8)
