"""CPU test of the N>1 path (world_size 2, gloo): the engine shards POINTS across ranks
(psba_local_range) and sum-all-reduces the camera-side quantities.  Here every rank runs the CPU
oracle on its own shard, the partial U / ga / S / ea / cost are all-reduced over gloo, and the result
must equal the unsharded computation ("partial-sum + reduce == unsharded to 1e-12")."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(prob, rank, world):
    import psba_b200
    p0, p1, o0, o1 = psba_b200.local_range(prob["n"], prob["o"], prob["iidx"], rank, world)
    sub = dict(prob)
    sub.update(n=p1 - p0, o=o1 - o0, pts=prob["pts"][p0:p1], impts=prob["impts"][o0:o1],
               iidx=(prob["iidx"][o0:o1] - p0).astype(np.int32), jidx=prob["jidx"][o0:o1])
    return sub, (p0, p1, o0, o1)


def _stage(P, mu):
    cost = P.call("exQT")
    P.call("jacobiQT"); P.call("U", 1); P.call("V", 1); P.call("Wblks", 1); P.call("g", 1)
    U = P.buf("U").copy(); ga = P.buf("g")[:P.N].copy()
    # damping is added AFTER the reduction in the engine (k_add_U); points carry it locally
    r = P.buf("V").reshape(-1, 9)
    r[:, 0] += mu; r[:, 4] += mu; r[:, 8] += mu
    P.call("Vinv"); P.call("Yblks"); P.call("S"); P.call("ea")
    return cost, U, ga, P.buf("S").copy(), P.buf("eab")[:P.N].copy()


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import oracle
    import psba_b200
    from util import dataset_paths
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = psba_b200.read_sba(*dataset_paths("54"))
    mu = 1.0e7
    sub, rng = _shard(prob, rank, world)
    P = oracle.Problem(sub)
    cost, U, ga, S, ea = _stage(P, mu)
    pack = torch.from_numpy(np.concatenate([[cost], U.ravel(), ga, S.ravel(), ea]))
    dist.all_reduce(pack, op=dist.ReduceOp.SUM)
    if rank == 0:
        F = oracle.Problem(prob)
        c0, U0, ga0, S0, ea0 = _stage(F, mu)
        ref = np.concatenate([[c0], U0.ravel(), ga0, S0.ravel(), ea0])
        got = pack.numpy()
        den = np.maximum(np.abs(ref), np.max(np.abs(S0)) * 1e-6)
        q.put((float(np.max(np.abs(got - ref) / den)), rng))
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharded_partial_sums_allreduce_equal_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, rng = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-10, err


def test_shards_cover_the_problem():
    import psba_b200
    from util import dataset_paths
    prob = psba_b200.read_sba(*dataset_paths("T21"))
    tot_o, tot_n = 0, 0
    for r in range(4):
        sub, (p0, p1, o0, o1) = _shard(prob, r, 4)
        tot_o += sub["o"]; tot_n += sub["n"]
        assert sub["iidx"].min() == 0 and sub["iidx"].max() == sub["n"] - 1
        assert abs(sub["o"] - prob["o"] / 4) < 20
    assert tot_o == prob["o"] and tot_n == prob["n"]
