"""Short run for ncu: ring problem, a few LM iterations through the C ABI (no timing claims)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import psba_b200
from psba_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "full"
its = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m, n, w = {"full": (2000, 1_000_000, 64), "mid": (500, 250_000, 64), "small": (64, 20_000, 64)}[name]
prob = synth.ring_problem(m=m, n=n, d=5, w=w, seed=20262000)
G = psba_b200.PSBA(prob)
G.set_option("lm_only", 1); G.set_option("max_iter", its)
flag, fe = G.levmar()
print("ok", name, "its", int(G.stat("itno")), "final", fe, "launches", int(G.stat("launches")))
G.close()
