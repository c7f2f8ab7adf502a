"""PSBA_ND_ROOT=1 on the GPU: the opt-in root rule of the tile ordering (kernels_chol.cu: degree of a root candidate counted inside
the subgraph) gives a different -- CPU-checked, tests/test_tile_plan_cpu.py -- plan for band-shaped camera systems.  This script
runs the engine under that plan: S and dpa = S^-1 ea of a damped system against the oracle's S and numpy, then the LM trajectory and a whole
LM + trust-region solve (modified Cholesky on the tile pool) against the default plan.  It runs in its own process (the GPU test calls it
with a time limit) because no GPU run had seen that plan when round 2 ended.
  PSBA_ND_ROOT=1 python tools/nd_root_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle
import psba_b200
from psba_b200 import synth
from util import pattern, relerr


def steps_of(prob, root):
    os.environ["PSBA_ND_ROOT"] = root
    G = psba_b200.PSBA(prob)
    s = int(G.stat("n_steps")), int(G.stat("nt"))
    G.close()
    return s


def main():
    # ---- a banded ring whose plan the switch changes (70 tiles: 29 -> 26 dependent steps), N = 3360 > 1536
    prob = synth.ring_problem(m=560, n=10000, d=4, w=40, seed=5)
    s0, s1 = steps_of(prob, "0"), steps_of(prob, "1")
    print("ring 560/40: steps %d (default) -> %d (PSBA_ND_ROOT=1) of %d tiles" % (s0[0], s1[0], s1[1]))
    assert s1[0] < s0[0] < s0[1]
    # stage check against the oracle's S and numpy: S, S^-1, dpa of a damped system
    os.environ["PSBA_ND_ROOT"] = "1"
    O = oracle.Problem(prob)
    O.set("nthreads", min(16, len(os.sched_getaffinity(0))))
    G = psba_b200.PSBA(prob)
    O.call("exQT"); O.call("jacobiQT"); O.call("U", 1); O.call("V", 1); O.call("Wblks", 1); O.call("g", 1)
    mu = 1e-3 * float(np.max(O.buf("UVdiag")))
    O.call("update_UV", mu); O.call("Vinv"); O.call("Yblks"); O.call("S"); O.call("ea")
    So = O.buf("S").copy()
    O.close()
    G.compute_jacobiQT(); G.compute_U(1.0); G.update_UV(mu); G.compute_Vinv()
    S = G.compute_S()
    assert relerr(np.tril(S), np.tril(So)) < 1e-10
    ea = G.compute_ea()
    assert G.SPDinv() == 0.0
    assert relerr(G.matVec_mul(), np.linalg.solve(np.tril(So) + np.tril(So, -1).T, ea)) < 1e-8
    G.restore_UVdiag(); G.close()
    print("S and dpa = S^-1 ea under the new plan: ok")
    # LM trajectory and the whole LM + trust-region solve (modified Cholesky on the tile pool) against the DEFAULT plan, which the
    # GPU suite pins to the oracle (the oracle's dense camera solve takes minutes at N = 3360)
    runs = {}
    for root in ("0", "1"):
        os.environ["PSBA_ND_ROOT"] = root
        G = psba_b200.PSBA(prob)
        G.set_option("lm_only", 1); G.set_option("max_iter", 8)
        flag, fe = G.levmar()
        lm = (flag, G.trace())
        G.close()
        G = psba_b200.PSBA(prob)
        r = G.solve()
        runs[root] = (lm, r, int(G.stat("cholmod_events")))
        G.close()
    (lm0, r0, ev0), (lm1, r1, ev1) = runs["0"], runs["1"]
    assert lm0[0] == lm1[0] and pattern(lm0[1]) == pattern(lm1[1]) and len(lm1[1]) >= 5
    for a, b in zip(lm0[1], lm1[1]):
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
    print("LM trajectory equal to the default plan's: ok (%d tries)" % len(lm1[1]))
    # the damping of a trust-region phase is a rounding-noise quantity, the converged cost is not
    print("exit flags %s / %s, outer iterations %d / %d" % (r0["flag"], r1["flag"], r0["itno"], r1["itno"]))
    assert abs(r1["finalErr"] - r0["finalErr"]) / r0["finalErr"] < 1e-6, (r1["finalErr"], r0["finalErr"])
    print("whole solve with the tile-pool modified Cholesky: ok (final cost %.9e against %.9e, %d / %d events)" % (r1["finalErr"], r0["finalErr"], ev1, ev0))
    print("ND_ROOT_CHECK PASS")


if __name__ == "__main__":
    main()
