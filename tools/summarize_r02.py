#!/usr/bin/env python
"""Turns the round-2 evidence visit (tools/gpu_final.sh: one `ncu --set full` report per kernel and workload, the ncu
launch list of the bench command) into the tracked summaries under profiles/, and writes the SASS opcode histograms of
the shipped library.

    python tools/summarize_r02.py [tag]
"""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
from summarize_profile import KEYS  # noqa: E402

EXTRA = [("lts__t_bytes.sum", "L2 bytes"), ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: lg throttle"),
         ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
         ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch")]


def short_name(name):
    s = name.split("(")[0].replace("void ", "").replace("vkpbrt::", "").replace(" ", "").replace("(int)", "").replace("(bool)", "")
    m = re.match(r"k_bmfr_block<(\d+),(\d+)", s)
    if m:
        return f"k_bmfr_block<{m.group(1)},{m.group(2)}>"
    m = re.match(r"k_bfr_block<(\d+)", s)
    if m:
        return f"k_bfr_block<{m.group(1)}>"
    return re.sub(r"<.*", "", s)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = ROOT / "profiles"
    reps = sorted((ROOT / "gpurun_out").glob(f"prof_{tag}_*.ncu-rep"))
    by_workload = collections.defaultdict(list)
    for rep in reps:
        m = re.match(rf"prof_{tag}_(.*?)_(k_.*)\.ncu-rep", rep.name)
        by_workload[m.group(1)].append(rep)
    tp, ip = out / "roofline_traffic.json", out / "roofline_instructions.json"
    allt = json.loads(tp.read_text()) if tp.exists() else {}
    alli = json.loads(ip.read_text()) if ip.exists() else {}
    for workload, files in by_workload.items():
        md = [f"# ncu --set full summary ({tag}, workload {workload})", "",
              "One report per kernel (scratch, not tracked): `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 6 -c 3",
              f"python bench.py --workload {workload} --steps 6 --warmup 3 --resident-frames 10 --cpu-budget 0` (tools/gpu_profile.sh); the launch with the",
              "median duration of each kernel is shown.  Durations under ncu are cold-cache and serialised: bench.py's CUDA-event times are the",
              "numbers that count; the counters explain them.", ""]
        traffic, instr = {}, {}
        for rep in files:
            raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(raw.splitlines()))
            hdr, units = rows[0], rows[1]
            idx = {h: i for i, h in enumerate(hdr)}
            groups = collections.defaultdict(list)
            for r in rows[2:]:
                groups[r[idx["Kernel Name"]]].append(r)
            for name, rs in groups.items():
                rs.sort(key=lambda r: float(r[idx["gpu__time_duration.sum"]].replace(",", "")))
                r = rs[len(rs) // 2]
                sn = short_name(name)
                md += [f"## {sn}  (`{name}`; grid {r[idx['Grid Size']]}, block {r[idx['Block Size']]}, {len(rs)} launches captured)", "",
                       "| metric | value | unit |", "|---|---|---|"]
                for k, label in KEYS + EXTRA:
                    if k in idx:
                        md.append(f"| {label} (`{k}`) | {r[idx[k]]} | {units[idx[k]]} |")
                md.append("")

                def num(k):
                    v, u = float(r[idx[k]].replace(",", "")), units[idx[k]]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic[sn] = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
                instr[sn] = int(float(r[idx["smsp__inst_executed.sum"]].replace(",", "")))
        (out / f"{tag}_ncu_summary_{workload}.md").write_text("\n".join(md))
        allt[workload], alli[workload] = traffic, instr
        print(workload, {k: (round(v / 1e6, 1), "MB") for k, v in traffic.items()}, {k: round(v / 1e6, 1) for k, v in instr.items()})
    tp.write_text(json.dumps(allt, indent=1))
    ip.write_text(json.dumps(alli, indent=1))
    # ---- launch list of the bench command
    launches = ROOT / "gpurun_out" / f"launches_{tag}.csv"
    if launches.exists():
        lr = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
        d = collections.defaultdict(list)
        for r in lr:
            d[r[4]].append(float(r[-1]) / 1e3)
        tot = sum(sum(v) for v in d.values())
        lines = [f"# ncu launch list ({tag}): `python bench.py --gpus 1 --steps 4 --warmup 3 --no-also --cpu-budget 0` (the driver's command, fewer steps)", "",
                 "gpu__time_duration.sum per launch, --clock-control none.  Cold-cache, serialised launches: compare SHARES with bench.py's",
                 "CUDA-event shares (`kernels.*.share`), not absolutes.  Every launch is one of the repository's own kernels.", "",
                 "| kernel | launches | median us | min us | share of profiled time |", "|---|---|---|---|---|"]
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| `{k}` | {len(v)} | {sorted(v)[len(v) // 2]:.1f} | {min(v):.1f} | {sum(v) / tot:.3f} |")
        (out / f"{tag}_launches_bmfr_taa_4k.md").write_text("\n".join(lines) + "\n")
        (out / f"{tag}_launches_bmfr_taa_4k.csv").write_text(open(launches).read())
        print("\n".join(lines[6:]))
    # ---- SASS opcode histograms of the shipped library
    so = ROOT / "vulkanpbrt_b200" / "lib" / "libvkpbrt_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
    fn, hist = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    lines = [f"# SASS opcode histograms of vulkanpbrt_b200/lib/libvkpbrt_b200.so ({tag})", "",
             "`cuobjdump -sass` of the shipped library, static instruction counts per kernel (top opcodes + the ones that identify the",
             "Blackwell paths).  `UTMALDG` = TMA tile load (cp.async.bulk.tensor), `SYNCS` = mbarrier, `FFMA2` / `FMUL2` = packed fp32 pairs",
             "(fma.rn.f32x2 / mul.rn.f32x2).  No `UTC*MMA` / `LDTM` (tcgen05) by design: the path is batched 1024x13 Householder QR and",
             "streaming stencils, not a GEMM.", ""]
    mark = ("UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FMNMX3", "MUFU", "BAR", "SHFL", "LDG", "STG", "LDS", "STS")
    for f in sorted(hist):
        if "k_" not in f:
            continue
        c = hist[f]
        total = sum(c.values())
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(10))
        flagged = ", ".join(f"**{k} {c[k]}**" for k in mark if c.get(k))
        lines += [f"## `{f.split('(')[0]}`", "", f"{total} instructions.  Top: {top}.", "", f"Markers: {flagged}.", ""]
    (out / f"{tag}_sass_opcodes.md").write_text("\n".join(lines))
    print("sass histograms:", len([f for f in hist if "k_" in f]), "kernels")


if __name__ == "__main__":
    main()
