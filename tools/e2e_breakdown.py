"""Where the end-to-end time of bench.py's e2e leg goes: wall-clock of every ABI call from page-locked host arrays
(setup_cl, fill_initBuffer2, fill_idxBuffer, K LM iterations, get_params), twice (first use / warm pool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, ctypes as C, torch
import psba_b200
from psba_b200 import synth, _d, _i
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
prob = synth.ring_problem()
hp = psba_b200.pinned_problem(prob)
L = psba_b200.lib()
m, n, o = prob["m"], prob["n"], prob["o"]
outb = (psba_b200.pinned_array(np.zeros((m, 6))), psba_b200.pinned_array(np.zeros((n, 3))))
for rep in range(3):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    G = psba_b200.PSBA.__new__(psba_b200.PSBA)
    G.L = L; G.m, G.n, G.o = m, n, o; G.N = 6 * m
    G.h = C.c_void_p(L.psba_setup_cl(6, 3, 2, m, n, o)); t.append(time.perf_counter())
    L.psba_fill_initBuffer2(G.h, 6, 3, 2, m, n, o, _d(hp["K"]), _d(hp["impts"]), _d(hp["initrot"]), _d(hp["cams"]), _d(hp["pts"])); t.append(time.perf_counter())
    L.psba_fill_idxBuffer(G.h, m, n, o, _i(hp["iidx"]), _i(hp["jidx"])); t.append(time.perf_counter())
    G.n_loc = int(G.stat("n_local")); G.o_loc = int(G.stat("o_local")); G.T_loc = G.N + 3 * G.n_loc; G._dims = (6, 3, 2, n, m, o)
    G.set_option("itno", 0); G.set_option("max_iter", K); G.set_option("lm_only", 1)
    G.levmar(); torch.cuda.synchronize(); t.append(time.perf_counter())
    G.get_params(out=outb); torch.cuda.synchronize(); t.append(time.perf_counter())
    G.close(); t.append(time.perf_counter())
    names = ["setup_cl", "fill_initBuffer2", "fill_idxBuffer", "levmar x%d" % K, "get_params", "close"]
    print("rep %d: total (without close) %.2f ms: " % (rep, 1e3 * (t[5] - t[0])) + ", ".join("%s %.2f" % (a, 1e3 * (t[i + 1] - t[i])) for i, a in enumerate(names)), flush=True)
