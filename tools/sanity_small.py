"""Small runs for compute-sanitizer: a ring problem with many segments / tasks through LM, and a whole LM + trust-region solve
of Trafalgar-21 (fused trust region, chains as graphs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import psba_b200
from psba_b200 import synth
from util import dataset_paths
prob = synth.ring_problem(m=160, n=6000, d=4, w=12, seed=7)
G = psba_b200.PSBA(prob)
G.set_option("lm_only", 1); G.set_option("max_iter", 3)
print("ring", G.levmar(), int(G.stat("n_seg")), int(G.stat("ring_rows")))
G.close()
prob = psba_b200.read_sba(*dataset_paths("7"))
G = psba_b200.PSBA(prob)
print("7cams", G.solve(), int(G.stat("seq_replays")))
G.close()
