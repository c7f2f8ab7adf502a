"""Device time of the phases of one LM iteration through the ABI (timer_start / timer_ms around repeated calls)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import psba_b200
from psba_b200 import synth
prob = synth.ring_problem()
G = psba_b200.PSBA(prob)
G.compute_exQT()
G.linearize()
mu = 1e-3 * G.maxElmOfUV()[0]
G.try_step(mu)
def t(name, f, reps=20):
    f()
    G.set_option("timer_start", 0)
    for _ in range(reps): f()
    print("%-28s %.4f ms" % (name, G.stat("timer_ms") / reps), flush=True)
t("linearize (both passes)", lambda: G.linearize())
t("try_step (whole try)", lambda: G.try_step(mu))
G.set_option("profile", 1)
t("linearize, profile mode (serial)", lambda: G.linearize())
G.close()
