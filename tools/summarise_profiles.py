"""Turn gpurun_out/{launches.csv, full_*.ncu-rep} into the tracked summaries under profiles/.

    python tools/summarise_profiles.py [round_tag]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    return {n: (v[h.index(n)], u[h.index(n)]) for n in WANT if n in h}


def stall_top(rep, n=8):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for r in data:
        for i in cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    tot = sum(agg.values()) or 1
    return [(k, round(100.0 * v / tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:n]]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    summary = {}
    lines = ["# ncu summaries (%s)\n" % tag,
             "Command: `python tools/profile_run.py --blocks 2960 --reps 1` (one full wave, FCX_LANES=1), see tools/gpu_round2_final1.sh.\n"]
    # launch list
    src = os.path.join(OUT, "launches.csv")
    if os.path.exists(src):
        rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
        tot = sum(float(r[-1]) for r in rows) or 1.0
        lines.append("\n## Launch list (gpu__time_duration, cold, serialised: compare shares)\n\n| kernel | grid | block | ms | share |\n|---|---|---|---|---|\n")
        with open(os.path.join(PROF, "launches_%s.csv" % tag), "w") as f:
            f.write("kernel,grid,block,duration_ms\n")
            for r in rows:
                name = r[4].split("(")[0].replace("void ", "")
                unit = r[-2]
                ms = float(r[-1]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
                f.write("%s,%s,%s,%.4f\n" % (name, r[8].replace(",", " "), r[7].replace(",", " "), ms))
        tot_ms = 0.0
        parsed = []
        for r in rows:
            unit = r[-2]
            ms = float(r[-1]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            parsed.append((r[4].split("(")[0].replace("void ", ""), r[8], r[7], ms))
            tot_ms += ms
        for name, grid, block, ms in parsed:
            lines.append("| %s | %s | %s | %.3f | %.1f %% |\n" % (name, grid, block, ms, 100 * ms / tot_ms))
    for fn in sorted(os.listdir(OUT)):
        if fn.startswith("full_") and fn.endswith(".ncu-rep"):
            k = fn[5:-8]
            rep = os.path.join(OUT, fn)
            m = raw_metrics(rep)
            st = stall_top(rep)
            dram = 0.0
            for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if n in m:
                    dram += float(m[n][0].replace(",", "")) * TO_BYTES.get(m[n][1], 1)
            summary[k] = {"dram_bytes": dram, "metrics": {a: " ".join(b) for a, b in m.items()}, "stalls_pct": st}
            lines.append("\n## %s (ncu --set full)\n\n" % k)
            for a, b in m.items():
                lines.append("* `%s` = %s %s\n" % (a, b[0], b[1]))
            lines.append("* top stall reasons (%% of samples): %s\n" % ", ".join("%s %.1f" % x for x in st))
    # compute-sanitizer logs of the same session (tools/sanitize_run.py)
    for tool in ("memcheck", "racecheck"):
        fn = os.path.join(OUT, "sanitize_%s.log" % tool)
        if os.path.exists(fn):
            txt = open(fn, errors="replace").read().splitlines()
            keep = [l for l in txt if "ERROR SUMMARY" in l or "SANITIZE RUN" in l or l.startswith(tool + " rc=") or "RACECHECK SUMMARY" in l]
            lines.append("\n## compute-sanitizer --tool %s (python tools/sanitize_run.py)\n\n" % tool)
            lines.extend("* `%s`\n" % l.strip() for l in keep[-6:])
    # pairs in the profiled launch (from the log of the launch-list run)
    log = os.path.join(OUT, "launches.log")
    pairs = None
    if os.path.exists(log):
        for l in open(log):
            if l.startswith("workload:"):
                pairs = int(l.split("blocks,")[1].split("pairs")[0])
    for k in summary:
        summary[k]["pairs"] = pairs
    json.dump(summary, open(os.path.join(PROF, "ncu_summary.json"), "w"), indent=1, sort_keys=True)
    open(os.path.join(PROF, "ncu_%s.md" % tag), "w").writelines(lines)
    print("".join(lines))


if __name__ == "__main__":
    main()
