"""Turn ncu artefacts brought back in gpurun_out/ into the tracked summaries
under profiles/ (run in the build container, no GPU needed):

    python profiles/summarize.py launches gpurun_out/r01_launches_c4.csv profiles/r01_launches_c4_summary.md
    python profiles/summarize.py kernel   gpurun_out/r01_fused2_eval_c4.ncu-rep c4 eval
    python profiles/summarize.py multi    gpurun_out/r01_c3_kernels.ncu-rep c3 "C3 multi-pass kernels"

`kernel` appends/updates profiles/fused_kernel_ncu.json (read by bench.py for
roofline.traffic) and writes profiles/<report>_summary.md.
"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
            rows.append((r["Kernel Name"], ns))
    agg = {}
    for name, ns in rows:
        short = name.split("(")[0][:110]
        c, t = agg.get(short, (0, 0.0))
        agg[short] = (c + 1, t + ns)
    total = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({os.path.basename(src)})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 "
                "--warmup 3 --no-cpu-baseline --no-e2e --cuda-graph 0` (eager launches; includes synthetic-data "
                "generation, warm-up and both timed modes). Per-launch times are cold-cache and serialised: "
                "compare shares, not absolutes.\n\n")
        f.write(f"{len(rows)} launches, {total / 1e6:.3f} ms total device time\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| `{name}` | {c} | {t / 1e6:.3f} | {t / total * 100:.1f}% | {t / c / 1e3:.1f} |\n")
    print("wrote", dst)


def kernel(rep, workload, mode):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, d = rows[0], rows[1], rows[2]
    rec = {"kernel": d[hdr.index("Kernel Name")], "report": os.path.basename(rep)}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            rec[k] = {"value": float(d[i].replace(",", "")), "unit": units[i]}
    stalls = {h.split("issue_stalled_")[1].split("_per_")[0]: float(d[i])
              for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("ratio")}
    rec["stall_ratio_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))

    def bytes_of(k):
        v, u = rec[k]["value"], rec[k]["unit"]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    dram = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
    rec["dram_bytes_per_launch"] = dram
    js = os.path.join(HERE, "fused_kernel_ncu.json")
    allrec = json.load(open(js)) if os.path.exists(js) else {}
    allrec.setdefault(workload, {})[mode] = rec
    json.dump(allrec, open(js, "w"), indent=1)
    md = os.path.join(HERE, os.path.basename(rep).replace(".ncu-rep", "_summary.md"))
    with open(md, "w") as f:
        f.write(f"# ncu --set full summary: {rec['kernel']}\n\nworkload {workload}, mode {mode}, report "
                f"`{os.path.basename(rep)}` (`--clock-control none`, 1 launch after warm-up)\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in KEYS:
            if k in rec:
                f.write(f"| {k} | {rec[k]['value']:.6g} | {rec[k]['unit']} |\n")
        f.write(f"| dram bytes per launch (read+write) | {dram:.6g} | byte |\n\n## warp stall reasons (per issued instruction)\n\n")
        for k, v in rec["stall_ratio_per_issue"].items():
            f.write(f"- {k}: {v:.3f}\n")
    print("wrote", md, "and", js)


def multi(rep, workload, title):
    """Every kernel of a multi-kernel `ncu --set full` report -> one markdown summary
    and profiles/multipass_kernels_ncu.json (keyed by workload / kernel)."""
    if rep.endswith(".csv"):   # raw page exported on the GPU box (`ncu -i rep --page raw --csv`)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    js = os.path.join(HERE, "multipass_kernels_ncu.json")
    allrec = json.load(open(js)) if os.path.exists(js) else {}
    md = os.path.join(HERE, os.path.basename(rep).replace(".ncu-rep", "_summary.md").replace("_raw.csv", "_summary.md"))
    with open(md, "w") as f:
        f.write(f"# ncu --set full summary: {title}\n\nworkload {workload}, report `{os.path.basename(rep)}` "
                f"(`--clock-control none`, launches after warm-up)\n")
        seen = set()
        for d in rows[2:]:
            name = d[hdr.index("Kernel Name")]
            short = name.split("(")[0]
            if short in seen:
                continue
            seen.add(short)
            rec = {"kernel": name, "report": os.path.basename(rep)}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    rec[k] = {"value": float(d[i].replace(",", "")), "unit": units[i]}
            stalls = {h.split("issue_stalled_")[1].split("_per_")[0]: float(d[i])
                      for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("ratio")}
            rec["stall_ratio_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
            allrec.setdefault(workload, {})[short] = rec
            f.write(f"\n## `{short}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in rec:
                    f.write(f"| {k} | {rec[k]['value']:.6g} | {rec[k]['unit']} |\n")
            top = list(rec["stall_ratio_per_issue"].items())[:6]
            f.write("\nstalls per issued instruction: " + ", ".join(f"{k} {v:.2f}" for k, v in top) + "\n")
    json.dump(allrec, open(js, "w"), indent=1)
    print("wrote", md, "and", js)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "multi":
        multi(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4])
