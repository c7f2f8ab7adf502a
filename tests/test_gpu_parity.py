"""GPU parity tests: the CUDA path (through the C ABI of libpsba_b200.so) against the CPU oracle on
the same inputs.  Tolerances: FP64, per-iteration cost 1e-9 relative, final cost 1e-6 relative
(north star); stage outputs at 1e-10..1e-8 relative to the largest entry (different but fixed
summation orders); index structure bit-exact."""
import numpy as np
import pytest

import oracle
import psba_b200
from util import dataset_paths, pattern, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["7", "54", "T21"])
def pair(request):
    c, p, cnp = dataset_paths(request.param)
    prob = psba_b200.read_sba(c, p, cnp)
    O = oracle.Problem(prob, kind="restatement")
    G = psba_b200.PSBA(prob)
    yield request.param, prob, O, G
    G.close()
    O.close()


def test_stages_first_iteration(pair):
    key, prob, O, G = pair
    # residual / cost
    c_o = O.call("exQT")
    c_g, ex = G.compute_exQT(want=True)
    assert relerr(ex, O.buf("ex")) < 1e-12
    assert abs(c_g - c_o) / c_o < 1e-12
    # Jacobian (materialised only for this check)
    O.call("jacobiQT")
    JA, JB = G.compute_jacobiQT(want=True)
    assert relerr(JA, O.buf("JA")) < 1e-11
    assert relerr(JB, O.buf("JB")) < 1e-11
    # block assembly
    O.call("U", 1); O.call("V", 1); O.call("Wblks", 1); O.call("g", 1)
    assert relerr(G.compute_U(1.0), O.buf("U")) < 1e-11
    assert relerr(G.compute_V(1.0), O.buf("V")) < 1e-11
    assert relerr(G.compute_Wblks(1.0), O.buf("W")) < 1e-11
    assert relerr(G.compute_g(1.0), O.buf("g")) < 1e-11
    mx, uv = G.maxElmOfUV()
    assert relerr(uv, O.buf("UVdiag")) < 1e-11
    mu = 1e-3 * float(np.max(O.buf("UVdiag")))
    assert abs(1e-3 * mx - mu) / mu < 1e-11
    # damping + Schur
    O.call("update_UV", mu); O.call("Vinv"); O.call("Yblks"); O.call("S"); O.call("ea")
    G.update_UV(mu)
    ret, Vmix = G.compute_Vinv()
    assert ret == 0.0
    assert relerr(Vmix, O.buf("V")) < 1e-10
    assert relerr(G.compute_Yblks(), O.buf("Y")) < 1e-10
    S = G.compute_S()
    So = O.buf("S")
    assert relerr(np.tril(S), np.tril(So)) < 1e-10
    assert relerr(G.compute_ea(), O.buf("eab")[:O.N]) < 1e-10
    # camera solve
    assert G.SPDinv() == 0.0
    dpa = G.matVec_mul()
    assert O.call("SPDinv") == 0.0
    O.call("matVec")
    assert relerr(dpa, O.buf("dp")[:O.N]) < 1e-7
    # back substitution
    O.call("eb"); O.call("dpb")
    eab = G.compute_eb()
    dp = G.compute_dpb()
    assert relerr(eab[O.N:], O.buf("eab")[O.N:]) < 1e-7
    assert relerr(dp, O.buf("dp")) < 1e-7
    # candidate + cost
    O.call("newp")
    newp = G.compute_newp()
    assert relerr(newp[:O.N], O.buf("newcams").ravel()) < 1e-9
    assert relerr(newp[O.N:], O.buf("newpts").ravel()) < 1e-9
    cn_o = O.call("exQT_new")
    cn_g = G.compute_exQT(psba_b200.PARAMS_NEW)
    assert abs(cn_g - cn_o) / cn_o < 1e-9
    # J*x with fused dot product
    jg_o = O.call("Jmultiply_g")
    jg_g = G.compute_Jmultiply(psba_b200.VEC_G)
    assert abs(jg_g - jg_o) / jg_o < 1e-11
    G.restore_UVdiag()


def test_explicit_inverse_small():
    c, p, cnp = dataset_paths("7")
    prob = psba_b200.read_sba(c, p, cnp)
    O = oracle.Problem(prob); G = psba_b200.PSBA(prob)
    for X in (O,):
        X.call("exQT"); X.call("jacobiQT"); X.call("U", 1); X.call("V", 1); X.call("Wblks", 1); X.call("g", 1)
    mu = 1e-3 * float(np.max(O.buf("UVdiag")))
    O.call("update_UV", mu); O.call("Vinv"); O.call("Yblks"); O.call("S")
    G.compute_jacobiQT(); G.compute_U(1.0); G.update_UV(mu); G.compute_Vinv(); G.compute_S(want=False)
    ret, Sinv = G.SPDinv(want=True)
    assert ret == 0.0 and O.call("SPDinv") == 0.0
    assert relerr(Sinv, O.buf("S")) < 1e-7
    G.close(); O.close()


def test_full_solve_lm_phase_and_final(pair):
    """LM phase: per-iteration cost 1e-9, identical accept pattern; whole solve: same iteration count,
    same accept/shrink pattern, final cost 1e-6 (north star tolerances)."""
    key, prob, O, G = pair
    O2 = oracle.Problem(prob, kind="restatement")
    G2 = psba_b200.PSBA(prob)
    fo = O2.solve()
    to = O2.trace()
    rg = G2.solve()
    tg = G2.trace()
    lm_o = [r for r in to if r["phase"] == 0]
    lm_g = [r for r in tg if r["phase"] == 0]
    n_first = 5   # the first LM phase: 5 accepted steps, then TR
    for a, b in zip(lm_o[:n_first], lm_g[:n_first]):
        assert a["accepted"] == b["accepted"]
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
        assert abs(a["mu"] - b["mu"]) / a["mu"] < 1e-9
    assert abs(rg["initErr"] - O2.get("initErr")) / O2.get("initErr") < 1e-12
    assert rg["flag"] == fo
    assert rg["itno"] == int(O2.get("itno"))
    assert pattern(tg) == pattern(to)
    assert abs(rg["finalErr"] - O2.get("finalErr")) / O2.get("finalErr") < 1e-6
    G2.close(); O2.close()


def test_tr_phase_lambda_follow(pair):
    """TR phase with the oracle's lambda injected (lambda-follow mode).

    At the LM->TR switch S is numerically singular (7-dim gauge null space, eigenvalues at rounding
    level), and lambda ~ 1e-5..1e-4 does not lift it above rounding: the Gauss-Newton step P_B then
    carries rounding-determined gauge components, so intermediate TR costs are NOT reproducible between
    two builds of the reference's own arithmetic either (~1e-4 relative; the oracle's
    variants P and R differ the same way).  What is reproducible -- and asserted -- is the accept/shrink
    sequence, the iteration count, every ACCEPTED cost to 2e-2 and the final cost to 1e-9."""
    key, prob, O, G = pair
    O2 = oracle.Problem(prob, kind="restatement")
    fo = O2.solve()
    to = O2.trace()
    lams = [r["mu"] for r in to if r["phase"] == 2]
    G2 = psba_b200.PSBA(prob)
    G2.force_lambda(lams)
    rg = G2.solve()
    tg = G2.trace()
    assert rg["flag"] == fo
    assert pattern(tg) == pattern(to)
    for a, b in zip(to, tg):
        if a["phase"] == 1 and a["accepted"]:
            assert abs(a["err"] - b["err"]) / a["err"] < 2e-2, (a, b)
        if a["phase"] == 2:
            assert a["mu"] == b["mu"]
    assert abs(rg["finalErr"] - O2.get("finalErr")) / O2.get("finalErr") < 1e-9
    G2.close(); O2.close()


def test_cholmod_matches_oracle():
    c, p, cnp = dataset_paths("7")
    prob = psba_b200.read_sba(c, p, cnp)
    O = oracle.Problem(prob); G = psba_b200.PSBA(prob)
    # TR-style linearisation at the start point, lambda = 0
    O.call("exQT"); O.call("jacobiQT"); O.call("g", -2); O.call("U", 2); O.call("V", 2); O.call("Wblks", 2)
    O.call("update_UV", 0.0); O.call("Vinv"); O.call("Yblks"); O.call("S")
    So = O.buf("S").copy()
    G.compute_jacobiQT(); G.compute_g(-2.0); G.compute_U(2.0)
    S = G.compute_S()
    assert relerr(np.tril(S), np.tril(So)) < 1e-10
    # run both modified Choleskys on the SAME matrix entries: feed the oracle the GPU's S
    O.buf("S")[:] = S
    sum_o = O.call("cholmod")
    res = G.cholmod_blk()
    assert res["n_scalar_blocks"] == int(O.get("ret"))
    # E_i = sum_k L_ik^2 - S_ii is a difference of O(S_ii) numbers: compare at the scale of
    # the diagonal, not of E itself (nvcc contracts a*b+c into FMA, the oracle is built without)
    scale = float(np.max(np.abs(np.diag(S))))
    assert float(np.max(np.abs(res["E"] - O.buf("E")))) < 1e-9 * scale
    G.close(); O.close()


def test_two_gpu_sharded_solve_matches_oracle():
    """N>1 on real GPUs (skipped on a 1-GPU box): tools/mgpu_check.py under torchrun, NCCL all-reduces"""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tools", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MGPU_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["Venice-52-64053", "Ladybug-138-19878"])
def test_bal_synthetic_structure_full_solve(name):
    """BASELINE configs 3-4: full LM + trust-region solve on the shipped BAL cameras with seeded synthetic
    structure (the BAL pts files are missing from the reference checkout).  Same tolerances
    as the shipped datasets; the CPU oracle runs the reference's algorithm with CSR index lookups."""
    import os
    from psba_b200 import synth
    from util import data_file
    n = int(name.split("-")[2])
    prob = synth.bal_structure_problem(data_file(name + "-cams.txt"), n, synth.BAL_OBS[name], name=name)
    O = oracle.Problem(prob)
    O.set("nthreads", min(16, len(os.sched_getaffinity(0))))
    G = psba_b200.PSBA(prob)
    if name.startswith("Ladybug"):
        # N = 828: the oracle's serial dense inverse is slow; compare the LM phase (5 accepted iterations) only
        fo = O.levmar()
        G.set_option("max_iter", 50)
        fg, fe = G.levmar()
        assert fo == fg == 2          # ITER_TURN_TO_TR
    else:
        fo = O.solve()
        rg = G.solve()
        assert rg["flag"] == fo and rg["itno"] == int(O.get("itno"))
        assert abs(rg["finalErr"] - O.get("finalErr")) / O.get("finalErr") < 1e-6
    to, tg = O.trace(), G.trace()
    assert pattern(tg) == pattern(to)
    for a, b in zip([r for r in to if r["phase"] == 0][:5], [r for r in tg if r["phase"] == 0][:5]):
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
        assert abs(a["mu"] - b["mu"]) / a["mu"] < 1e-9
    G.close(); O.close()


@pytest.mark.parametrize("nd_min", ["0", "2", "12"])
def test_banded_ring_nested_dissection_solve(nd_min, monkeypatch):
    """Block-banded camera system (ring of 160 cameras, window 12): the solver orders the 48x48 tiles by nested
    dissection and runs independent panels in the same step.  Every ordering (natural, deepest, default) must
    give the reference's S, S^-1, dpa and LM costs."""
    from psba_b200 import synth
    monkeypatch.setenv("PSBA_ND_MIN", nd_min)
    prob = synth.ring_problem(m=160, n=6000, d=4, w=12, seed=7)
    O = oracle.Problem(prob)
    O.set("nthreads", 8)
    G = psba_b200.PSBA(prob)
    nsteps, nt = int(G.stat("n_steps")), int(G.stat("nt"))
    assert nt == 20
    assert nsteps == nt if nd_min == "0" else nsteps < nt
    O.call("exQT"); O.call("jacobiQT"); O.call("U", 1); O.call("V", 1); O.call("Wblks", 1); O.call("g", 1)
    mu = 1e-3 * float(np.max(O.buf("UVdiag")))
    O.call("update_UV", mu); O.call("Vinv"); O.call("Yblks"); O.call("S"); O.call("ea")
    So = O.buf("S").copy()
    G.compute_jacobiQT(); G.compute_U(1.0); G.update_UV(mu); G.compute_Vinv()
    S = G.compute_S()
    assert relerr(np.tril(S), np.tril(So)) < 1e-10
    ea = G.compute_ea()
    ret, Sinv = G.SPDinv(want=True)
    assert ret == 0.0
    Sfull = np.tril(So) + np.tril(So, -1).T
    ref_inv = np.linalg.inv(Sfull)
    assert relerr(Sinv, ref_inv) < 1e-8
    dpa = G.matVec_mul()
    assert relerr(dpa, ref_inv @ ea) < 1e-8
    G.restore_UVdiag()
    G.close()
    # LM iterations: costs and damping sequence against the oracle
    G = psba_b200.PSBA(prob)
    assert O.levmar() == G.levmar()[0]          # 5 accepted LM iterations, then ITER_TURN_TO_TR
    to, tg = O.trace(), G.trace()
    assert pattern(tg) == pattern(to) and len(to) >= 5
    for a, b in zip(to, tg):
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
        assert abs(a["mu"] - b["mu"]) / a["mu"] < 1e-9
    G.close(); O.close()


@pytest.mark.parametrize("key", ["7", "54"])
def test_device_built_index_structure_is_bit_exact(key):
    """The index structure is built on the device (structure.cu).  It must reproduce the reference's tables of
    generate_idxs (PSBA/misc.cpp:178-218) exactly: blk_idx (observation of point i in camera j), comm3DIdx /
    comm3DIdxCnt (common points of every camera pair, ascending), and the camera-major scan order of compute_U."""
    c, p, cnp = dataset_paths(key)
    prob = psba_b200.read_sba(c, p, cnp)
    O = oracle.Problem(prob, kind="restatement", dense=True)
    G = psba_b200.PSBA(prob)
    m, n, o = prob["m"], prob["n"], prob["o"]
    blk = O.ibuf("blk_idx", n * m).reshape(n, m)
    comm = O.ibuf("comm3DIdx", n * m * m).reshape(m, m, n)
    cnt = O.ibuf("comm3DIdxCnt", m * m).reshape(m, m)
    # CSR by point and camera-major order
    ptr = G.index("pt_ptr")
    assert np.array_equal(ptr, np.concatenate([[0], np.cumsum(np.bincount(prob["iidx"], minlength=n))]))
    assert np.array_equal(G.index("cam_obs"), np.argsort(prob["jidx"], kind="stable"))
    # camera pairs k >= l: every pair with a common point + every diagonal, sorted by (k, l)
    pk, pl = G.index("pair_k"), G.index("pair_l")
    want = [(k, l) for k in range(m) for l in range(k + 1) if cnt[k, l] > 0 or k == l]
    assert list(zip(pk.tolist(), pl.tolist())) == want
    # triples of every pair = the rows of comm3DIdx, with the observation ids blk_idx gives
    oa, ob, pt = G.index("tri_oa"), G.index("tri_ob"), G.index("tri_pt")
    # chunks (pair-major ids): the triples of one pair inside one segment of the camera row; contiguous per pair
    pcp, cb, ce = G.index("pair_chunk_ptr"), G.index("chunk_beg"), G.index("chunk_end")
    assert np.array_equal(cb[1:], ce[:-1]) and cb[0] == 0 and ce[-1] == len(oa)
    beg = {}; end = {}
    for pid in range(len(want)):
        if pcp[pid + 1] > pcp[pid]:
            beg[pid] = int(cb[pcp[pid]]); end[pid] = int(ce[pcp[pid + 1] - 1])
    # segments: consecutive visits of one camera (camera-major positions); every off-diagonal chunk is scheduled exactly
    # once, in the segment that holds the visits of its triples, largest first; the diagonal chunk is the segment's own
    sd = G.index("seg_desc").reshape(-1, 6)
    sched = G.index("sched_chunk")
    cam_obs = G.index("cam_obs")
    cam_pos = np.empty(o, dtype=np.int64); cam_pos[cam_obs] = np.arange(o)
    seen = np.zeros(len(cb), dtype=int)
    for row, v0, v1, dchunk, s0, s1 in sd.tolist():
        assert 0 < v1 - v0 <= int(G.stat("seg_v")) and np.all(prob["jidx"][cam_obs[v0:v1]] == row)
        sizes = []
        for cidx in [dchunk] + sched[s0:s1].tolist():
            pos = cam_pos[oa[cb[cidx]:ce[cidx]]]
            assert np.all((pos >= v0) & (pos < v1))
            seen[cidx] += 1
            sizes.append(int(ce[cidx] - cb[cidx]))
        assert oa[cb[dchunk]] == ob[cb[dchunk]] and sizes[0] == v1 - v0
        assert sizes[1:] == sorted(sizes[1:], reverse=True)
    assert np.all(seen == 1)
    assert len(oa) == sum(int(cnt[k, l]) for k, l in want)
    for pid, (k, l) in enumerate(want):
        c_kl = int(cnt[k, l])
        assert cnt[l, k] == c_kl
        if c_kl == 0:
            assert pid not in beg
            continue
        b, e = beg[pid], end[pid]
        assert e - b == c_kl
        pts_kl = comm[k, l, :c_kl]
        assert np.array_equal(pts_kl, comm[l, k, :c_kl]) and np.all(np.diff(pts_kl) > 0)
        assert np.array_equal(pt[b:e], pts_kl)
        assert np.array_equal(oa[b:e], blk[pts_kl, k]) and np.array_equal(ob[b:e], blk[pts_kl, l])
    G.close(); O.close()


def _mixed_track_problem():
    """160 ring cameras; 300 points seen by 150 cameras each (more observations than one 128-wide wave: the
    'oversize chunk' path of the point-major kernels) followed by 2000 ordinary points with 4 observations."""
    from psba_b200 import synth
    a = synth.ring_problem(m=160, n=300, d=150, w=160, seed=11)
    b = synth.ring_problem(m=160, n=2000, d=4, w=12, seed=12)
    prob = dict(a)
    prob["n"] = a["n"] + b["n"]; prob["o"] = a["o"] + b["o"]
    prob["pts"] = np.concatenate([a["pts"], b["pts"]]); prob["impts"] = np.concatenate([a["impts"], b["impts"]])
    prob["iidx"] = np.concatenate([a["iidx"], b["iidx"] + a["n"]]).astype(np.int32)
    prob["jidx"] = np.concatenate([a["jidx"], b["jidx"]]).astype(np.int32)
    prob["name"] = "ring-160-mixed-tracks"
    return prob


def _check_try_against_oracle(prob, G):
    """one linearisation + one damped solve + LM trajectory against the oracle"""
    O = oracle.Problem(prob)
    O.set("nthreads", 8)
    O.call("exQT"); O.call("jacobiQT"); O.call("U", 1); O.call("V", 1); O.call("Wblks", 1); O.call("g", 1)
    G.compute_jacobiQT()
    assert relerr(G.compute_U(1.0), O.buf("U")) < 1e-11
    assert relerr(G.compute_V(1.0), O.buf("V")) < 1e-11
    assert relerr(G.compute_Wblks(1.0), O.buf("W")) < 1e-11
    assert relerr(G.compute_g(1.0), O.buf("g")) < 1e-11
    mu = 1e-3 * float(np.max(O.buf("UVdiag")))
    O.call("update_UV", mu); O.call("Vinv"); O.call("Yblks"); O.call("S"); O.call("ea")
    So, eao = O.buf("S").copy(), O.buf("eab")[:O.N].copy()
    G.update_UV(mu)
    ret, Vmix = G.compute_Vinv()
    assert ret == 0.0 and relerr(Vmix, O.buf("V")) < 1e-10
    assert relerr(np.tril(G.compute_S()), np.tril(So)) < 1e-10
    assert relerr(G.compute_ea(), eao) < 1e-10
    assert G.SPDinv() == 0.0 and O.call("SPDinv") == 0.0
    G.matVec_mul(); O.call("matVec")
    O.call("eb"); O.call("dpb"); O.call("newp")
    assert relerr(G.compute_eb()[O.N:], O.buf("eab")[O.N:]) < 1e-7
    assert relerr(G.compute_dpb(), O.buf("dp")) < 1e-7
    G.compute_newp()
    cn_o, cn_g = O.call("exQT_new"), G.compute_exQT(psba_b200.PARAMS_NEW)
    assert abs(cn_g - cn_o) / cn_o < 1e-9
    G.restore_UVdiag()
    return O


@pytest.mark.parametrize("mode,segv,G_", [("6", "", ""), ("6", "40", "0:1"), ("6", "100", "1:2"), ("6", "1400", "2:64"), ("6", "33", "4:3"), ("6", "200", "5:5"),
                                          ("5", "", ""), ("5", "48", "1"), ("5", "100", "2"), ("5", "100", "8"), ("5", "1400", "32"), ("0", "", "")])
def test_pair_pass_variants_agree_with_oracle(mode, segv, G_, monkeypatch):
    """PSBA_PAIR_MODE selects the pair pass: 6 = ring kernel (default; launch shape `cfg` and rows per task `rt` varied here as
    "cfg:rt", segment lengths too), 5 = segment kernel (Y staged per camera-row segment; segment lengths and lane-group sizes
    varied here), 0 = the pair-major gather kernel.  All must give the reference's S and ea (compute_S.cl / compute_ea.cl)
    and the same LM trajectory."""
    from psba_b200 import synth
    monkeypatch.setenv("PSBA_PAIR_MODE", mode)
    if segv:
        monkeypatch.setenv("PSBA_SEG_V", segv)
    if G_ and mode == "6":
        monkeypatch.setenv("PSBA_RING_CFG", G_.split(":")[0]); monkeypatch.setenv("PSBA_RING_RT", G_.split(":")[1])
    elif G_:
        monkeypatch.setenv("PSBA_SEG_G", G_)
    for prob in (synth.ring_problem(m=160, n=6000, d=4, w=12, seed=7), psba_b200.read_sba(*dataset_paths("54"))):
        G = psba_b200.PSBA(prob)
        assert int(G.stat("pair_mode")) == int(mode)
        if mode in ("5", "6"):
            assert int(G.stat("n_seg")) >= prob["m"] - 1
        O = _check_try_against_oracle(prob, G)
        G.close()
        G = psba_b200.PSBA(prob)
        assert O.levmar() == G.levmar()[0]
        to, tg = O.trace(), G.trace()
        assert pattern(tg) == pattern(to)
        for a, b in zip(to, tg):
            assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
        G.close(); O.close()


@pytest.mark.parametrize("mode", ["6", "5", "0"])
def test_points_with_more_observations_than_one_wave(mode, monkeypatch):
    """Tracks of 150 observations (> 128, one CTA wave) next to ordinary ones: the pipelined point-major kernels
    take the one-wave chunks, the wave-loop kernels the oversize ones; the pair pass has no limit on track length."""
    monkeypatch.setenv("PSBA_PAIR_MODE", mode)
    prob = _mixed_track_problem()
    G = psba_b200.PSBA(prob)
    O = _check_try_against_oracle(prob, G)
    G.close()
    G = psba_b200.PSBA(prob)
    assert O.levmar() == G.levmar()[0]
    to, tg = O.trace(), G.trace()
    assert pattern(tg) == pattern(to)
    for a, b in zip(to, tg):
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
    G.close(); O.close()


def _ragged_problem():
    """ring of 24 cameras where camera 5 sees nothing, every seventh point has ONE observation and the others two or
    three: empty camera rows / pair runs, rank-deficient V_i (regularised by the damping only)"""
    from psba_b200 import synth
    a = synth.ring_problem(m=24, n=600, d=3, w=8, seed=5)
    keep = np.ones(a["o"], dtype=bool)
    keep[a["jidx"] == 5] = False
    first = np.concatenate([[True], a["iidx"][1:] != a["iidx"][:-1]])
    keep[(a["iidx"] % 7 == 0) & ~first] = False
    # a point whose only remaining candidates were camera 5 keeps its first observation
    cnt = np.bincount(a["iidx"][keep], minlength=a["n"])
    for i in np.nonzero(cnt == 0)[0]:
        keep[np.nonzero(a["iidx"] == i)[0][0]] = True
    prob = dict(a)
    prob["o"] = int(keep.sum())
    prob["iidx"] = a["iidx"][keep].astype(np.int32); prob["jidx"] = a["jidx"][keep].astype(np.int32)
    prob["impts"] = a["impts"][keep]
    # observations kept "to save a point" may belong to camera 5: move them to camera 6 if that keeps cameras ascending
    bad = prob["jidx"] == 5
    prob["jidx"][bad] = 4
    order_ok = np.all((prob["iidx"][1:] > prob["iidx"][:-1]) | ((prob["iidx"][1:] == prob["iidx"][:-1]) & (prob["jidx"][1:] > prob["jidx"][:-1])))
    assert order_ok
    prob["name"] = "ring-24-ragged"
    return prob


@pytest.mark.parametrize("mode", ["6", "5", "0"])
def test_ragged_structure_empty_camera_and_single_observation_points(mode, monkeypatch):
    monkeypatch.setenv("PSBA_PAIR_MODE", mode)
    prob = _ragged_problem()
    assert not np.any(prob["jidx"] == 5) and np.bincount(prob["iidx"]).min() == 1
    G = psba_b200.PSBA(prob)
    O = _check_try_against_oracle(prob, G)
    G.close()
    G = psba_b200.PSBA(prob)
    assert O.levmar() == G.levmar()[0]
    to, tg = O.trace(), G.trace()
    assert pattern(tg) == pattern(to) and len(to) >= 2
    for a, b in zip(to, tg):
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
    G.close(); O.close()


@pytest.mark.parametrize("key", ["7", "T21"])
def test_indefinite_camera_system_is_reported_not_factored(key):
    """Failure semantics of the camera solve (SPD_inv.cl:66-107, cl_spdinv.cpp:18-40): a system that is not positive
    definite makes SPDinv return 1.0 -- here forced with a NEGATIVE damping term -- and the fused try reports
    solve_status 1 with NaN costs, so that levmar.cpp:227-244 can raise mu and try again.  A regular try on the same
    context afterwards must still work."""
    prob = psba_b200.read_sba(*dataset_paths(key))
    O = oracle.Problem(prob); G = psba_b200.PSBA(prob)
    O.call("exQT"); O.call("jacobiQT"); O.call("U", 1); O.call("V", 1); O.call("Wblks", 1); O.call("g", 1)
    mx = float(np.max(O.buf("UVdiag")))
    O.call("update_UV", -0.5 * mx); O.call("Vinv"); O.call("Yblks"); O.call("S"); O.call("ea")
    assert O.call("SPDinv") == 1.0
    G.compute_exQT(); G.linearize(1.0, 1.0)
    G.update_UV(-0.5 * mx); G.compute_Vinv(); G.compute_S()
    assert G.SPDinv() == 1.0
    G.restore_UVdiag()
    res = G.try_step(-0.5 * mx)
    assert res["solve_status"] == 1.0 and np.isnan(res["cost_new"])
    res = G.try_step(1e-3 * mx)
    assert res["solve_status"] == 0.0 and np.isfinite(res["cost_new"]) and res["cost_new"] > 0
    G.close(); O.close()


def test_page_locked_host_arrays_give_the_same_solve():
    """psba_host_alloc / psba_host_free: the same problem handed over in page-locked arrays, parameters read back into
    page-locked arrays -- identical trajectory and parameters (the ABI takes any host pointer)."""
    prob = psba_b200.read_sba(*dataset_paths("54"))
    G = psba_b200.PSBA(prob)
    flag_a, fe_a = G.levmar()
    cams_a, pts_a = G.get_params()
    G.close()
    hp = psba_b200.pinned_problem(prob)
    G = psba_b200.PSBA(hp)
    flag_b, fe_b = G.levmar()
    out = (psba_b200.pinned_array(np.zeros((prob["m"], 6))), psba_b200.pinned_array(np.zeros((G.n_loc, 3))))
    cams_b, pts_b = G.get_params(out=out)
    G.close()
    assert flag_a == flag_b and fe_a == fe_b
    assert np.array_equal(cams_a, cams_b) and np.array_equal(pts_a, pts_b)
    del hp, out, cams_b, pts_b
    psba_b200.free_pinned()


def _numpy_residuals(prob, cams, pts):
    """e = measured - projected for every observation with the reference's camera model (compute_exQT.cl:33-69),
    written independently of the CUDA code: ql = (sqrt(1-|v|^2), v), q = ql (x) q0, Xc = R(q) X + t,
    x = (fu Xc + s Yc + u0 Zc) / Zc, y = (fu ar Yc + v0 Zc) / Zc."""
    j, i = prob["jidx"], prob["iidx"]
    v = cams[:, :3]
    ql = np.concatenate([np.sqrt(1.0 - (v * v).sum(1, keepdims=True)), v], axis=1)
    q0 = prob["initrot"]
    a0, a1, a2, a3 = ql.T
    b0, b1, b2, b3 = q0.T
    q = np.stack([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                  a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0], axis=1)
    s, x, y, z = q.T
    R = np.empty((len(q), 3, 3))
    R[:, 0, 0] = s * s + x * x - y * y - z * z; R[:, 0, 1] = 2 * (x * y - s * z); R[:, 0, 2] = 2 * (x * z + s * y)
    R[:, 1, 0] = 2 * (x * y + s * z); R[:, 1, 1] = s * s - x * x + y * y - z * z; R[:, 1, 2] = 2 * (y * z - s * x)
    R[:, 2, 0] = 2 * (x * z - s * y); R[:, 2, 1] = 2 * (y * z + s * x); R[:, 2, 2] = s * s - x * x - y * y + z * z
    Xc = np.einsum("oab,ob->oa", R[j], pts[i]) + cams[j, 3:6]
    K = prob["K"][j]
    px = (K[:, 0] * Xc[:, 0] + K[:, 4] * Xc[:, 1] + K[:, 1] * Xc[:, 2]) / Xc[:, 2]
    py = (K[:, 0] * K[:, 3] * Xc[:, 1] + K[:, 2] * Xc[:, 2]) / Xc[:, 2]
    return prob["impts"] - np.stack([px, py], axis=1)


def test_full_size_properties_of_the_headline_workload():
    """BASELINE.json's single-GPU configuration (ring: 2 000 cameras, 1 M points, 5 M observations) is out of the
    oracle's reach (its dense tables are 14.6 TB); at that size the CUDA path is checked through size-independent
    properties: (1) every residual against an independent numpy statement of the camera model, before and after a
    step; (2) the damped normal equations -- the step the Schur path returns satisfies BOTH block rows of
    (J^T J + mu I) dp = J^T e built from the library's own U, V, W, g; (3) the cost the fused try reports is the cost
    of the parameters it produced; (4) two runs give the same bits (fixed summation order everywhere)."""
    from psba_b200 import synth
    prob = synth.ring_problem(m=2000, n=1_000_000, d=5, w=64, seed=20262000)
    m, n, o, N = prob["m"], prob["n"], prob["o"], 6 * prob["m"]
    G = psba_b200.PSBA(prob)
    cost0, ex = G.compute_exQT(want=True)
    e_np = _numpy_residuals(prob, prob["cams"], prob["pts"])
    assert relerr(ex, e_np) < 1e-11
    assert abs(cost0 - float((e_np * e_np).sum())) / cost0 < 1e-12
    G.linearize(1.0, 1.0)
    mx, _ = G.maxElmOfUV()
    mu = 1e-3 * mx
    res = G.try_step(mu)
    assert res["solve_status"] == 0.0
    dp = G.compute_dpb()                                   # [dpa (6m) | dpb (3n)] of this try
    U, V, W, g = G.compute_U(1.0), G.compute_V(1.0), G.compute_Wblks(1.0), G.compute_g(1.0)
    dpa, dpb = dp[:N].reshape(m, 6), dp[N:].reshape(n, 3)
    i, j = prob["iidx"], prob["jidx"]
    # point rows:  (V_i + mu I) dpb_i + sum_j W_ij^T dpa_j = gb_i
    wt_dpa = np.einsum("orc,or->oc", W, dpa[j])
    rb = np.einsum("nab,nb->na", V, dpb) + mu * dpb - g[N:].reshape(n, 3)
    for cc in range(3):
        rb[:, cc] += np.bincount(i, weights=wt_dpa[:, cc], minlength=n)
    # camera rows: (U_j + mu I) dpa_j + sum_i W_ij dpb_i = ga_j
    w_dpb = np.einsum("orc,oc->or", W, dpb[i])
    ra = np.einsum("mab,mb->ma", U, dpa) + mu * dpa - g[:N].reshape(m, 6)
    for rr in range(6):
        ra[:, rr] += np.bincount(j, weights=w_dpb[:, rr], minlength=m)
    gn = float(np.linalg.norm(g))
    assert float(np.linalg.norm(rb)) / gn < 1e-9 and float(np.linalg.norm(ra)) / gn < 1e-9
    # the candidate the try evaluated: its cost is the cost of p + dp
    cams1, pts1 = prob["cams"] + dpa, prob["pts"] + dpb
    e1 = _numpy_residuals(prob, cams1, pts1)
    assert abs(res["cost_new"] - float((e1 * e1).sum())) / res["cost_new"] < 1e-11
    assert res["cost_new"] < cost0
    G.close()
    # bit-reproducible: two fresh runs of three LM iterations
    finals = []
    for _ in range(2):
        G = psba_b200.PSBA(prob)
        G.set_option("lm_only", 1); G.set_option("max_iter", 3)
        flag, fe = G.levmar()
        finals.append((flag, fe, tuple(r["err"] for r in G.trace())))
        G.close()
    assert finals[0] == finals[1] and finals[0][1] < res["cost_new"]


def test_full_size_index_structure_against_numpy():
    """The device-built index structure at the full size of the headline workload (15 M camera-pair triples; the
    reference's comm3DIdx would be 14.6 TB): CSR by point, camera-major order and the pair-sorted triples are rebuilt
    with numpy -- stable sorts in the enumeration order of generate_idxs (PSBA/misc.cpp:189-217) -- and compared
    bit for bit."""
    from psba_b200 import synth
    prob = synth.ring_problem(m=2000, n=1_000_000, d=5, w=64, seed=20262000)
    m, n, o, d = prob["m"], prob["n"], prob["o"], 5
    G = psba_b200.PSBA(prob)
    assert np.array_equal(G.index("pt_ptr"), np.arange(n + 1, dtype=np.int64) * d)
    assert np.array_equal(G.index("cam_obs"), np.argsort(prob["jidx"], kind="stable"))
    # triples of a point in emission order: a = 0..d-1, b = 0..a (cameras ascend inside a point)
    la = np.concatenate([np.full(a + 1, a) for a in range(d)]); lb = np.concatenate([np.arange(a + 1) for a in range(d)])
    base = (np.arange(n, dtype=np.int64) * d)[:, None]
    oa, ob = (base + la[None, :]).ravel(), (base + lb[None, :]).ravel()
    key = prob["jidx"][oa].astype(np.int64) * m + prob["jidx"][ob]
    order = np.argsort(key, kind="stable")                  # by pair (k, l); ascending point inside a pair
    assert int(G.stat("ntriples")) == len(order)
    assert np.array_equal(G.index("tri_oa"), oa[order]) and np.array_equal(G.index("tri_ob"), ob[order])
    assert np.array_equal(G.index("tri_pt"), prob["iidx"][oa[order]])
    uk = np.unique(np.concatenate([key, np.arange(m, dtype=np.int64) * (m + 1)]))     # pairs present + every diagonal
    assert np.array_equal(G.index("pair_k").astype(np.int64) * m + G.index("pair_l"), uk)
    G.close()
