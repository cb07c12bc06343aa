#!/usr/bin/env python
"""Turns an `ncu --set full` report + an ncu launch list into the tracked summaries under profiles/.

    python tools/summarize_profile.py gpurun_out/prof_r4.ncu-rep gpurun_out/launches_r4.csv r01 bmfr_1080p
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared/MUFU)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (i-cache)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
]


def main():
    rep, launches, tag, workload = sys.argv[1:5]
    out = ROOT / "profiles"
    out.mkdir(exist_ok=True)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    md = [f"# ncu --set full summary ({tag}, workload {workload})", "",
          f"Source report: `{Path(rep).name}` (scratch, not tracked); command: see tools/gpu_round.sh.", ""]
    traffic = {}
    instructions = {}
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        md += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
        for k, label in KEYS:
            if k in idx:
                md.append(f"| {label} (`{k}`) | {r[idx[k]]} | {units[idx[k]]} |")
        md.append("")

        def num(k):
            v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        short = name.split("(")[0].replace("void ", "").replace("vkpbrt::", "").replace(" ", "")
        traffic[short] = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
        if "smsp__inst_executed.sum" in idx:
            instructions[short] = int(float(r[idx["smsp__inst_executed.sum"]].replace(",", "")))
    (out / f"{tag}_ncu_summary_{workload}.md").write_text("\n".join(md))
    # launch list: per-kernel statistics + the raw csv
    lr = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
    d = collections.defaultdict(list)
    for r in lr:
        d[r[4]].append(float(r[-1]) / 1e3)
    tot = sum(sum(v) for v in d.values())
    lines = [f"# ncu launch list ({tag}, {workload}): gpu__time_duration.sum per launch, --clock-control none", "",
             "(cold-cache, serialised launches: compare SHARES with bench.py's CUDA-event shares, not absolutes)", "",
             "| kernel | launches | median us | min us | share of profiled time |", "|---|---|---|---|---|"]
    for k, v in d.items():
        lines.append(f"| `{k}` | {len(v)} | {sorted(v)[len(v) // 2]:.1f} | {min(v):.1f} | {sum(v) / tot:.3f} |")
    (out / f"{tag}_launches_{workload}.md").write_text("\n".join(lines) + "\n")
    Path(out / f"{tag}_launches_{workload}.csv").write_text(open(launches).read())
    tp = out / "roofline_traffic.json"
    allt = json.loads(tp.read_text()) if tp.exists() else {}
    allt[workload] = {("k_bmfr_block<32,256>" if "bmfr" in k else k): v for k, v in traffic.items()}
    tp.write_text(json.dumps(allt, indent=1))
    ip = out / "roofline_instructions.json"       # warp instructions per launch: bench.py's issue-slot roofline
    alli = json.loads(ip.read_text()) if ip.exists() else {}
    alli[workload] = {("k_bmfr_block<32,256>" if "bmfr" in k else k): v for k, v in instructions.items()}
    ip.write_text(json.dumps(alli, indent=1))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
