"""Summarise ncu outputs into profiles/ (tracked).
  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/launches_r01_summary.md
  python tools/ncu_summary.py full gpurun_out/prof_r1b.ncu-rep profiles/ncu_full_r01.md
"""
import collections
import csv
import io
import subprocess
import sys

mode, src, dst = sys.argv[1:4]
if mode == "launches":
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        if name.startswith("cub::"):
            name = name.split("<")[0] + "<...> (setup)"
        v = float(r[vi].replace(",", ""))
        v_us = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1; t[1] += v_us
    total = sum(t[1] for t in tot.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3 "
                "--no-e2e --no-cpu-baseline` (includes the one-off device structure build) (cold-cache, serialised launches: compare SHARES, not absolutes).\n\n" % src)
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, us) in sorted(tot.items(), key=lambda x: -x[1][1]):
            f.write("| %s | %d | %.1f | %.2f | %.1f%% |\n" % (k, n, us, us / n, 100 * us / total))
else:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    seen = set()
    with open(dst, "w") as f:
        f.write("# ncu --set full summary (%s)\n\nOne capture per kernel (first launch of each), `--clock-control none`.\n" % src)
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
            if name in seen:
                continue
            seen.add(name)
            f.write("\n## %s\n\n| metric | value | unit |\n|---|---|---|\n" % name)
            for w in want:
                if w in idx:
                    f.write("| %s | %s | %s |\n" % (w, r[idx[w]], units[idx[w]]))
            vals = [(h, float(r[idx[h]].replace(",", ""))) for h in hdr
                    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[idx[h]] not in ("", "n/a")]
            t = sum(v for _, v in vals) or 1
            f.write("\nstall samples: " + ", ".join("%s %.0f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / t)
                                                   for h, v in sorted(vals, key=lambda x: -x[1])[:6]) + "\n")
