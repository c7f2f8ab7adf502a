"""Full LM + trust-region solve (psba_solve = PSBA/main.cpp:192-209) of a bench workload on one GPU: device time, iterations,
modified-Cholesky events.  python tools/full_solve_headline.py [workload]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import psba_b200

name = sys.argv[1] if len(sys.argv) > 1 else "ring-2000-1M-5M-w64"
prob = bench.make_problem(name)
G = psba_b200.PSBA(prob)
out = []
for rep in range(2):
    G.set_params(prob["cams"], prob["pts"])
    G.set_option("stats_reset", 0); G.set_option("lm_only", 0); G.set_option("max_iter", 50)
    G.set_option("timer_start", 0)
    t0 = time.perf_counter()
    r = G.solve()
    ms = G.stat("timer_ms")
    tr = G.trace()
    out.append({"workload": name, "ms": ms, "wall_ms": (time.perf_counter() - t0) * 1e3, "itno": r["itno"], "flag": r["flag"],
                "init_cost": r["initErr"], "final_cost": r["finalErr"], "tries": int(G.stat("tries")), "exqt": int(G.stat("exqt")),
                "cholmod_events": int(G.stat("cholmod_events")), "cholmod_max_l_over_beta": G.stat("cholmod_max_l_over_beta"),
                "pattern": "".join("C" if q["phase"] == 2 else ("A" if q["accepted"] else "x") for q in tr),
                "lambda": [q["mu"] for q in tr if q["phase"] == 2]})
    print(json.dumps(out[-1]), flush=True)
G.close()
