import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import psba_b200
from psba_b200 import synth
prob = synth.ring_problem(m=2000, n=200_000, d=5, w=64, seed=1)
G = psba_b200.PSBA(prob)
G.set_option("lm_only", 1); G.set_option("max_iter", 2)
G.levmar()
buf = (ctypes.c_longlong * 64)()
psba_b200.lib().psba_debug_panel_clocks(buf)
a = np.array(buf[:]).reshape(4, 16)
for b in range(2):
    t = a[b, :6]
    print("block", b, "loads %d  sts+syrk %d  transpose %d  sweep %d  store %d  total %d" % tuple([int(t[i + 1] - t[i]) for i in range(5)] + [int(t[5] - t[0])]))
