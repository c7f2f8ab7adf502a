"""SASS census of libpsba_b200.so: per kernel, how many of the instructions that matter on sm_100a it holds
(cuobjdump -sass; run anywhere nvcc's tools are installed, no GPU needed).
  python tools/sass_census.py > profiles/sass_census_rNN.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "psba_b200", "libpsba_b200.so")
OPS = [("DMMA", r"\bDMMA"), ("DFMA", r"\bDFMA"), ("UBLKCP", r"\bUBLKCP"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"),
       ("LDG.256", r"\bLDG\.[A-Z0-9.]*256"), ("LDG", r"\bLDG\b"), ("STG", r"\bSTG\b"), ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"),
       ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\."), ("ATOMG/RED", r"\b(ATOMG|RED|REDG)\b"), ("ATOMS", r"\bATOMS"),
       ("LDL", r"\bLDL\b"), ("STL", r"\bSTL\b"), ("MUFU", r"\bMUFU"), ("ACQBULK/GDC", r"\b(ACQBULK|PREEXIT)\b")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        ins = line.split("*/", 1)[-1]
        if not re.search(r"^\s+[@!A-Z]", ins):
            continue
        per[cur]["total"] += 1
        for name, pat in OPS:
            if re.search(pat, ins):
                per[cur][name] += 1
    dem = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
    def short(d):
        d = d.replace("void ", "")
        depth = 0
        for k, ch in enumerate(d):
            if ch == "<": depth += 1
            elif ch == ">": depth -= 1
            elif ch == "(" and depth == 0: return d[:k]
        return d
    names = [short(d) for d in dem] if len(dem) == len(per) else list(per)
    print("# SASS census of psba_b200/libpsba_b200.so\n")
    print("`cuobjdump -sass` instruction counts per kernel; cubin architectures in the library: %s.\n" % ", ".join(arch))
    print("| kernel | instr | " + " | ".join(n for n, _ in OPS) + " |")
    print("|---|---|" + "---|" * len(OPS))
    for (k, cnt), nm in sorted(zip(per.items(), names), key=lambda x: x[1]):
        if "cub" in nm or "thrust" in nm or nm.startswith("__cuda_sm"):
            continue
        print("| `%s` | %d | " % (nm, cnt["total"]) + " | ".join(str(cnt[n]) if cnt[n] else "" for n, _ in OPS) + " |")
    lib = sum(1 for nm in names if "cub" in nm or "thrust" in nm)
    print("\n%d CUB kernels (radix sort / scan / run-length encode of the set-up) are not listed.  tcgen05 (UTC*MMA) cannot appear:"
          " the path is FP64 and tcgen05 has no f64 kind; the dense contraction of the camera solve uses DMMA (mma.sync m8n8k4 f64)." % lib)


if __name__ == "__main__":
    main()
