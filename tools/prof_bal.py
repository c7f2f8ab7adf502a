"""Per-kernel device time of a whole solve (levmar + trust_region) on the Venice-52 synthetic-structure problem
and on Trafalgar-21: where the time of a SMALL problem goes (launch latency, host round trips, camera solve)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import psba_b200
from psba_b200 import synth
from util import data_file, dataset_paths
KN = ["k_cam_prep", "k_cost", "k_lin_points", "k_lin_cams", "k_cam_reduce", "k_vinv", "memset_S", "k_schur_pairs", "k_S_finalize",
      "chol_graph", "k_tri_solve", "k_newcams", "k_backsub", "k_reduce", "k_Jdot", "k_vec", "k_cholmod", "allreduce"]
which = sys.argv[1] if len(sys.argv) > 1 else "venice"
prob = (synth.bal_structure_problem(data_file("Venice-52-64053-cams.txt"), 64053, synth.BAL_OBS["Venice-52-64053"])
        if which == "venice" else psba_b200.read_sba(*dataset_paths("T21")))
G = psba_b200.PSBA(prob)
G.solve()
for prof in (0, 1):
    G.set_params(prob["cams"], prob["pts"])
    G.set_option("stats_reset", 0)
    G.set_option("profile", prof)
    G.set_option("timer_start", 0)
    t0 = time.perf_counter()
    r = G.solve()
    ms = G.stat("timer_ms")
    wall = (time.perf_counter() - t0) * 1e3
    print("profile=%d: device span %.3f ms, wall %.3f ms, itno %d, tries %d, exqt %d, launches %d, final %.9e | chains as graphs: %d keys, %d graphs, %d replays" %
          (prof, ms, wall, r["itno"], int(G.stat("tries")), int(G.stat("exqt")), int(G.stat("launches")), r["finalErr"],
           int(G.stat("seq_keys")), int(G.stat("seq_graphs")), int(G.stat("seq_replays"))))
tot = 0.0
for k in KN:
    n = G.stat("n." + k)
    if n:
        t = G.stat("ms." + k); tot += t
        print("  %-14s %4d launches  %8.3f ms total  %7.2f us avg" % (k, n, t, 1e3 * t / n))
print("  sum of kernels %.3f ms" % tot)
G.close()
