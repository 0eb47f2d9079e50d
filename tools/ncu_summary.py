"""Summarise an `ncu --set full` report of one kernel launch (development tool).

    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [--cell-steps N] [--json profiles/traffic_rNN.json --key fk4096]

Prints a markdown table of the metrics DESIGN.md quotes and, with --json, records the DRAM traffic of the launch
(dram__bytes_read.sum + dram__bytes_write.sum) for bench.py's `roofline.traffic`.
"""
import argparse
import csv
import io
import json
import subprocess

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "smsp__cycles_active.max", "sm__cycles_elapsed.max",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--cell-steps", type=float, default=0.0, help="cell-steps one launch produces (for per-cell-step figures)")
    ap.add_argument("--json", default=None)
    ap.add_argument("--key", default="fk4096")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    name = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
    print("kernel: `%s`\n" % name)
    print("| metric | value | unit |\n|---|---|---|")
    got = {}
    for m in METRICS:
        if m in col:
            got[m] = vals[col[m]]
            print("| %s | %s | %s |" % (m, vals[col[m]], units[col[m]]))
    stalls = sorted(((float(vals[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h, i in col.items()
                     if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h), reverse=True)
    tot = sum(v for v, _ in stalls) or 1.0
    print("\nwarp-state samples: " + ", ".join("%s %.1f %%" % (n, 100 * v / tot) for v, n in stalls if v > 0))

    def num(m, unit_scale=None):
        v, u = float(got[m].replace(",", "")), units[col[m]].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        return v * scale
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    print("\nDRAM traffic per launch: %.1f MB" % (traffic / 1e6))
    if args.cell_steps:
        print("per cell-step: %.2f B DRAM (algorithmic 28 B: %.0f %% saved), %.1f thread-instructions" % (
            traffic / args.cell_steps, 100 * (1 - traffic / args.cell_steps / 28.0),
            float(got["smsp__inst_executed.sum"].replace(",", "")) * 32 / args.cell_steps))
    if args.json:
        try:
            d = json.load(open(args.json))
        except Exception:
            d = {}
        d[args.key] = {"dram_bytes_per_launch": traffic, "kernel": name, "report": args.report,
                       "cell_steps_per_launch": args.cell_steps or None,
                       "gpu_time_us": float(got["gpu__time_duration.sum"].replace(",", ""))}
        json.dump(d, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
