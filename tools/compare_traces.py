"""Print oracle vs GPU run logs side by side (diagnostic; needs a GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle, psba_b200
from util import dataset_paths

follow = "--follow" in sys.argv
keys = [a for a in sys.argv[1:] if not a.startswith("--")] or ["7", "54", "T21"]
for key in keys:
    c, p, cnp = dataset_paths(key)
    prob = psba_b200.read_sba(c, p, cnp)
    O = oracle.Problem(prob); t0 = time.time(); fo = O.solve(); to_ = time.time() - t0
    to = O.trace()
    G = psba_b200.PSBA(prob)
    if follow:
        G.force_lambda([r["mu"] for r in to if r["phase"] == 2])
    t0 = time.time(); rg = G.solve(); tg_ = time.time() - t0
    tg = G.trace()
    print("== %s: oracle flag %d itno %d final %.15E (%.3fs) | gpu flag %d itno %d final %.15E (%.3fs) launches %d" % (
        key, fo, O.get("itno"), O.get("finalErr"), to_, rg["flag"], rg["itno"], rg["finalErr"], tg_, G.stat("launches")))
    for k in range(max(len(to), len(tg))):
        a = to[k] if k < len(to) else None
        b = tg[k] if k < len(tg) else None
        f = lambda r: "--" if r is None else "%d it%2d err %.12E rho %+.6f mu %.6E dk %8.4f %s" % (
            r["phase"], r["itno"], r["err"], r["rho"], r["mu"], r["delta"], "A" if r["accepted"] else "x")
        rel = "" if (a is None or b is None or a["phase"] == 2 or not np.isfinite(a["err"])) else " rel %.1e" % (abs(a["err"] - b["err"]) / abs(a["err"]))
        print("  ", f(a), "|", f(b), rel)
