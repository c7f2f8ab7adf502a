"""Engine-level multi-rank parity (run under torchrun, one rank per GPU, NCCL): the point-sharded engine against the SAME
engine on one rank -- cost, U, g, S, ea, dpa, dp of the first damped try and the LM trace, at 1e-12.  Exercises comm.cu
(NCCL init, all-reduces), the packed [S tiles | ea] and [U | ga] exchanges and the per-rank slicing of psba_api.cu.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/mgpu_stage_check.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import psba_b200
from psba_b200 import synth
from util import dataset_paths, pattern, relerr

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
L = psba_b200.lib()
TOL = 1e-12


def stages(prob):
    """first linearisation + first damped try + 6 LM iterations; everything a rank can see"""
    G = psba_b200.PSBA(prob)
    out = {"p_off": int(G.stat("p_off")), "n_loc": G.n_loc, "N": G.N}
    out["cost"] = G.compute_exQT()
    G.compute_jacobiQT()
    out["U"] = G.compute_U(1.0)
    g = G.compute_g(1.0)
    out["ga"], out["gb"] = g[:G.N].copy(), g[G.N:].copy()
    mx, _ = G.maxElmOfUV()
    out["maxdiag"] = mx
    mu = 1e-3 * mx
    G.update_UV(mu)
    G.compute_Vinv()
    out["S"] = np.tril(G.compute_S())
    out["ea"] = G.compute_ea()
    assert G.SPDinv() == 0.0
    out["dpa"] = G.matVec_mul()
    G.compute_eb()
    dp = G.compute_dpb()
    out["dpb"] = dp[G.N:].copy()
    G.restore_UVdiag()
    G.close()
    G = psba_b200.PSBA(prob)
    G.set_option("max_iter", 6); G.set_option("lm_only", 1)
    flag, fe = G.levmar()
    out["flag"], out["final"], out["trace"] = flag, fe, G.trace()
    G.close()
    return out


problems = {"54": psba_b200.read_sba(*dataset_paths("54")), "T21": psba_b200.read_sba(*dataset_paths("T21")),
            "ring-160": synth.ring_problem(m=160, n=6000, d=4, w=12, seed=7)}
single = {k: stages(p) for k, p in problems.items()}          # every rank alone, before the communicator exists

idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = ctypes.create_string_buffer(128)
    L.psba_comm_unique_id(buf)
    idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
dist.broadcast(idt, 0)
L.psba_comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))

ok = True
for key, prob in problems.items():
    a, b = single[key], stages(prob)
    lo, hi = b["p_off"], b["p_off"] + b["n_loc"]
    errs = {"cost": abs(a["cost"] - b["cost"]) / a["cost"], "maxdiag": abs(a["maxdiag"] - b["maxdiag"]) / a["maxdiag"],
            "U": relerr(b["U"], a["U"]), "ga": relerr(b["ga"], a["ga"]), "gb": relerr(b["gb"], a["gb"][3 * lo:3 * hi]),
            "S": relerr(b["S"], a["S"]), "ea": relerr(b["ea"], a["ea"]), "dpa": relerr(b["dpa"], a["dpa"]),
            "dpb": relerr(b["dpb"], a["dpb"][3 * lo:3 * hi])}
    # dpa / dpb go through the camera solve: conditioning amplifies the 1e-16 differences of the summation order
    tol = {k: TOL for k in errs}
    tol["dpa"] = tol["dpb"] = 1e-9
    tr = max(abs(x["err"] - y["err"]) / x["err"] for x, y in zip(a["trace"], b["trace"]))
    good = all(errs[k] < tol[k] for k in errs) and tr < TOL * 1e2 and pattern(a["trace"]) == pattern(b["trace"]) and a["flag"] == b["flag"]
    t = torch.tensor([0.0 if good else 1.0], device="cuda")
    dist.all_reduce(t)
    ok = ok and t.item() == 0.0
    print("rank %d/%d %s: points [%d,%d)  %s  trace %.1e  pattern %s -> %s" % (
        rank, world, key, lo, hi, " ".join("%s %.1e" % kv for kv in errs.items()), tr, pattern(b["trace"]), "OK" if good else "MISMATCH"), flush=True)
dist.barrier()
L.psba_comm_finalize()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_STAGE_CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
