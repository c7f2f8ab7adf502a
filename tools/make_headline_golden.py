"""Per-iteration LM costs of the bench workloads on ONE GPU -> tests/golden/headline_lm_costs.json.
bench.py compares the costs of every run (any number of GPUs) with these values (parity_vs_1gpu).
  python tools/make_headline_golden.py [out.json]      (needs a CUDA device)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import psba_b200

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "headline_lm_costs.json")
try:
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or "worktree"
except Exception:
    commit = "worktree"
res = {}
for name in ("ring-2000-1M-5M-w64", "ring-500-250k-1.25M-w64", "ring-16-2k-10k"):
    prob = bench.make_problem(name)
    G = psba_b200.PSBA(prob)
    G.set_option("itno", 0); G.set_option("max_iter", 40); G.set_option("lm_only", 1)
    flag, fe = G.levmar()
    tr = [r for r in G.trace() if r["phase"] == 0]
    res[name] = {"commit": commit, "costs": [r["err"] for r in tr if r["accepted"]], "mu": [r["mu"] for r in tr if r["accepted"]],
                 "pattern": "".join("A" if r["accepted"] else "x" for r in tr), "flag": flag, "final_cost": fe}
    G.close()
    print(name, len(res[name]["costs"]), "iterations, final cost %.15e" % fe)
json.dump(res, open(out, "w"), indent=1)
