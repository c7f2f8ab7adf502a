"""Generate tests/golden/*.json|npz from the REFERENCE'S OWN kernel bodies (oracle/_ref, compiled in place
from /root/reference by oracle/Makefile) driven by the restated LM / trust-region drivers.

Run in the build container (needs /root/reference):   python tools/make_golden.py
The fixtures pin the oracle restatement (tests/test_oracle_golden.py) and travel to the GPU box.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from util import DATASETS, dataset_paths, pattern  # noqa: E402

oracle.build(ref=True)
out = {}
for key in ("7", "9", "54", "54KD", "T21"):
    c, p, cnp = dataset_paths(key)
    prob = oracle.read_sba(c, p, cnp, kind="reference")
    P = oracle.Problem(prob, kind="reference")
    rec = dict(m=prob["m"], n=prob["n"], o=prob["o"], files=list(DATASETS[key][:2]), origin_cnp=cnp)
    rec["init_err"] = P.call("exQT")
    P.call("jacobiQT"); P.call("U", 1); P.call("V", 1); P.call("Wblks", 1); P.call("g", 1)
    uv = P.buf("UVdiag")
    rec["max_diag"] = float(uv.max())
    mu0 = 1e-3 * rec["max_diag"]
    rec["mu0"] = mu0
    rec["g_norm2"] = float(np.dot(P.buf("g"), P.buf("g")))
    P.call("update_UV", mu0); P.call("Vinv"); P.call("Yblks"); P.call("S"); P.call("ea")
    rec["S_fro2"] = float((P.buf("S") ** 2).sum())
    rec["ea_norm2"] = float(np.dot(P.buf("eab")[:P.N], P.buf("eab")[:P.N]))
    rec["sum_JA"] = float(P.buf("JA").sum()); rec["sum_JB"] = float(P.buf("JB").sum())
    rec["sum_W"] = float(P.buf("W").sum()); rec["sum_U"] = float(P.buf("U").sum()); rec["sum_V_mixed"] = float(P.buf("V").sum())
    if key == "7":   # small per-stage arrays for element-wise checks
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stage_7.npz"),
                            ex=P.buf("ex").copy(), JA=P.buf("JA")[:64].copy(), JB=P.buf("JB")[:64].copy(), W=P.buf("W")[:64].copy(),
                            U=P.buf("U").copy(), V_mixed=P.buf("V")[:32].copy(), Y=P.buf("Y")[:64].copy(), S=P.buf("S").copy(),
                            g=P.buf("g").copy(), ea=P.buf("eab")[:P.N].copy(), iidx=prob["iidx"], jidx=prob["jidx"],
                            blk_idx=prob["blk_idx"], comm3DIdxCnt=prob["comm3DIdxCnt"])
    P.close()
    # full solve (fresh state)
    P = oracle.Problem(prob, kind="reference")
    flag = P.solve()
    tr = P.trace()
    rec["flag"] = int(flag); rec["itno"] = int(P.get("itno")); rec["final_err"] = P.get("finalErr")
    rec["pattern"] = pattern(tr)
    rec["lm"] = [dict(itno=r["itno"], err=r["err"], rho=r["rho"], mu=r["mu"]) for r in tr if r["phase"] == 0][:5]
    rec["lambda"] = [r["mu"] for r in tr if r["phase"] == 2]
    rec["cholmod_scalar_blocks"] = [int(r["err"]) for r in tr if r["phase"] == 2]
    rec["tr_accepted_err"] = [r["err"] for r in tr if r["phase"] == 1 and r["accepted"]]
    P.close()
    out[key] = rec
    print(key, rec["init_err"], rec["final_err"], rec["itno"], rec["pattern"])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "reference_runs.json"), "w"), indent=1)
