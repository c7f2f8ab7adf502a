"""Per-launch table of the camera-solve kernels from one `ncu --set full` capture (raw page as CSV).
  python tools/ncu_panel_steps.py gpurun_out/prof_r02e_panel_raw.csv profiles/ncu_full_panel_r02e.md [append|new] [preamble.md]
One table row per captured launch (steps of the factorisation, k_panel_step; the backward solve, k_backward_flow).
"""
import csv
import sys

src, dst = sys.argv[1:3]
append = len(sys.argv) > 3 and sys.argv[3] == "append"
rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us", 1e-3 if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0),
        ("launch__grid_size", "CTAs", 1),
        ("launch__registers_per_thread", "regs", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %", 1),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe %", 1),
        ("launch__waves_per_multiprocessor", "waves", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %", 1),
        ("dram__bytes_read.sum", "DRAM rd", 1),
        ("dram__bytes_write.sum", "DRAM wr", 1)]
cols = [c for c in cols if c[0] in idx]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


with open(dst, "a" if append else "w") as f:
    f.write("# ncu --set full, the camera solve launch by launch (%s)\n\n" % src)
    f.write(open(sys.argv[4]).read() if len(sys.argv) > 4 else "")
    f.write("| # | kernel | " + " | ".join("%s" % c[1] + (" (%s)" % units[idx[c[0]]] if c[1].startswith("DRAM") else "") for c in cols) + " |\n")
    f.write("|---|---|" + "---|" * len(cols) + "\n")
    tot = {}
    stalls = []
    for n, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = [num(r[idx[c[0]]]) * c[2] for c in cols]
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1; t[1] += vals[0]
        f.write("| %d | %s | " % (n, name) + " | ".join(("%.1f" % v if v < 1e4 else "%.3g" % v) for v in vals) + " |\n")
        st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: num(r[idx[h]]) for h in hdr
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
        ssum = sum(st.values()) or 1.0
        stalls.append("launch %d: " % n + ", ".join("%s %.0f%%" % (k, 100 * v / ssum) for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
    f.write("\n")
    for k, (n, us) in tot.items():
        f.write("* `%s`: %d launches, %.1f us in total under ncu\n" % (k, n, us))
    f.write("\nWarp-state samples (share of all samples of the launch):\n\n")
    for s_ in stalls:
        f.write("* " + s_ + "\n")
