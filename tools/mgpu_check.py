"""Multi-GPU parity check (run under torchrun, one rank per GPU): the point-sharded engine with NCCL
all-reduces must reproduce the single-process CPU oracle's LM trajectory.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
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

import oracle
import psba_b200
from util import dataset_paths, pattern

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
L = psba_b200.lib()
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = ctypes.create_string_buffer(128)
    L.psba_comm_unique_id(buf)
    idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
dist.broadcast(idt, 0)
L.psba_comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
ok = True
for key in ("54", "T21"):
    prob = psba_b200.read_sba(*dataset_paths(key))
    G = psba_b200.PSBA(prob)
    res = G.solve()
    tg = G.trace()
    if rank == 0:
        O = oracle.Problem(prob)
        fo = O.solve()
        to = O.trace()
        lm_o = [r for r in to if r["phase"] == 0][:5]
        lm_g = [r for r in tg if r["phase"] == 0][:5]
        worst = max(abs(a["err"] - b["err"]) / a["err"] for a, b in zip(lm_o, lm_g))
        fin = abs(res["finalErr"] - O.get("finalErr")) / O.get("finalErr")
        good = worst < 1e-9 and fin < 1e-6 and pattern(tg) == pattern(to) and res["itno"] == int(O.get("itno")) and res["flag"] == fo
        ok = ok and good
        print("%s: world %d local points %d/%d  LM-phase max rel %.2e  final rel %.2e  pattern %s  itno %d -> %s" % (
            key, world, G.n_loc, prob["n"], worst, fin, pattern(tg), res["itno"], "OK" if good else "MISMATCH"))
    G.close()
dist.barrier()
L.psba_comm_finalize()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)
