"""Trust-region traces of the fused and of the explicit scalar evaluation side by side (psba_set_option "tr_fused")."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import psba_b200
from util import dataset_paths
key = sys.argv[1] if len(sys.argv) > 1 else "54"
prob = psba_b200.read_sba(*dataset_paths(key))
tr = []
for fused in (1, 0):
    G = psba_b200.PSBA(prob)
    G.set_option("tr_fused", fused)
    r = G.solve()
    tr.append([q for q in G.trace()])
    print("fused", fused, r)
    G.close()
for a, b in zip(*tr):
    print("ph %d it %2d acc %d | err %.12e %.12e rel %.1e | rho %.6f %.6f | pnorm %.6e %.6e | mu/lambda %.6e %.6e | delta %g %g" % (
        a["phase"], a["itno"], a["accepted"], a["err"], b["err"], abs(a["err"] - b["err"]) / max(abs(b["err"]), 1e-300), a["rho"], b["rho"], a["pnorm"], b["pnorm"], a["mu"], b["mu"], a["delta"], b["delta"]))
