"""CPU tests (no GPU): the oracle restatement against
  (a) the golden fixtures generated from the reference's OWN kernel bodies (tests/golden/reference_runs.json,
      stage_7.npz; made by tools/make_golden.py from oracle/_ref),
  (b) the known answers of SURVEY.md App. B.3 (hard-coded below, produced independently by the survey),
  (c) oracle/_ref itself, stage by stage, when that library is present.
"""
import json
import os

import numpy as np
import pytest

import oracle
from util import HERE, dataset_paths, pattern, relerr

GOLD = json.load(open(os.path.join(HERE, "golden", "reference_runs.json")))

# SURVEY.md App. B.3 (g++ -O2, no FMA contraction): dataset -> (init Err, max diag, ||g||^2, ||S||_F^2, ||ea||^2)
SURVEY_B3 = {
    "7": (3.658538570112267E+04, 1.189450690578604E+12, 1.577128965824467E+16, 8.128663312295264E+24, 8.728991143503093E+15),
    "9": (1.980233694849334E+04, 1.189930855124764E+12, 5.773442356328752E+15, 1.001215188134122E+25, 3.105807473246120E+15),
    "54": (5.283715930511977E+04, 1.090125282010044E+13, 2.984488838741704E+16, 2.586160994998789E+27, 1.877112986735656E+16),
    "T21": (2.081774076035723E+08, 3.551511953702390E+11, 2.994485769824892E+18, 6.358530685451889E+23, 2.957543371238794E+18),
}
# LM-phase costs (it0..it4) and exit: SURVEY App. B.3
SURVEY_LM = {
    "7": [3.564918636089513E+03, 2.558932097227388E+03, 2.097645800113637E+03, 1.845072556927081E+03, 1.630909782506780E+03],
    "54": [1.508308973656909E+04, 9.927799887419431E+03, 8.134379459136615E+03, 7.409506164349314E+03, 7.044021971412757E+03],
    "T21": [5.738307010905697E+06, 1.566735983845842E+06, 1.182840710331750E+06, 9.759618512038145E+05, 7.928565741559345E+05],
}
SURVEY_FINAL = {"7": (1.294058936035844E+03, 11), "9": (1.500570640935653E+03, 12),
                "54": (4.342837182193854E+03, 10), "T21": (3.034073752504546E+05, 24)}


def load(key, kind="restatement"):
    c, p, cnp = dataset_paths(key)
    return oracle.read_sba(c, p, cnp, kind=kind)


@pytest.mark.parametrize("key", ["7", "9", "54", "T21"])
def test_first_iteration_known_answers(key):
    prob = load(key)
    P = oracle.Problem(prob)
    g = GOLD[key]
    assert (prob["m"], prob["n"], prob["o"]) == (g["m"], g["n"], g["o"])
    init = P.call("exQT")
    P.call("jacobiQT"); P.call("U", 1); P.call("V", 1); P.call("Wblks", 1); P.call("g", 1)
    mx = float(P.buf("UVdiag").max())
    g2 = float(np.dot(P.buf("g"), P.buf("g")))
    P.call("update_UV", 1e-3 * mx); P.call("Vinv"); P.call("Yblks"); P.call("S"); P.call("ea")
    S2 = float((P.buf("S") ** 2).sum())
    ea2 = float(np.dot(P.buf("eab")[:P.N], P.buf("eab")[:P.N]))
    for got, ref_run, survey in zip((init, mx, g2, S2, ea2),
                                    (g["init_err"], g["max_diag"], g["g_norm2"], g["S_fro2"], g["ea_norm2"]), SURVEY_B3[key]):
        assert abs(got - ref_run) / ref_run < 1e-12
        assert abs(got - survey) / survey < 1e-12
    for name, buf in (("sum_JA", "JA"), ("sum_JB", "JB"), ("sum_W", "W"), ("sum_U", "U"), ("sum_V_mixed", "V")):
        assert abs(float(P.buf(buf).sum()) - g[name]) <= 1e-9 * max(abs(g[name]), float(np.abs(P.buf(buf)).max()))
    P.close()


def test_stage_arrays_7():
    """element-wise against arrays produced by the reference kernel bodies"""
    z = np.load(os.path.join(HERE, "golden", "stage_7.npz"))
    prob = load("7")
    assert np.array_equal(prob["iidx"], z["iidx"]) and np.array_equal(prob["jidx"], z["jidx"])
    P = oracle.Problem(prob, dense=True)
    # bit-exact index structure (generate_idxs semantics, misc.cpp:178-218)
    assert np.array_equal(P.ibuf("blk_idx", P.n * P.m).reshape(P.n, P.m), z["blk_idx"])
    assert np.array_equal(P.ibuf("comm3DIdxCnt", P.m * P.m).reshape(P.m, P.m), z["comm3DIdxCnt"])
    P.call("exQT"); P.call("jacobiQT"); P.call("U", 1); P.call("V", 1); P.call("Wblks", 1); P.call("g", 1)
    assert relerr(P.buf("ex"), z["ex"]) < 1e-13
    assert relerr(P.buf("JA")[:64], z["JA"]) < 1e-12
    assert relerr(P.buf("JB")[:64], z["JB"]) < 1e-12
    assert relerr(P.buf("W")[:64], z["W"]) < 1e-12
    assert relerr(P.buf("g"), z["g"]) < 1e-12
    mu = 1e-3 * float(P.buf("UVdiag").max())
    P.call("update_UV", mu); P.call("Vinv"); P.call("Yblks"); P.call("S"); P.call("ea")
    assert relerr(P.buf("U"), z["U"]) < 1e-12            # the fixture holds U after update_UV(mu0)
    assert relerr(P.buf("V")[:32], z["V_mixed"]) < 1e-12
    assert relerr(P.buf("Y")[:64], z["Y"]) < 1e-11
    assert relerr(P.buf("S"), z["S"]) < 1e-11
    assert relerr(P.buf("eab")[:P.N], z["ea"]) < 1e-11
    P.close()


@pytest.mark.parametrize("key", ["7", "9", "54", "54KD", "T21"])
def test_full_solve_matches_reference_runs(key):
    prob = load(key)
    P = oracle.Problem(prob)
    flag = P.solve()
    tr = P.trace()
    g = GOLD[key]
    assert flag == g["flag"] and int(P.get("itno")) == g["itno"]
    assert pattern(tr) == g["pattern"]
    lm = [r for r in tr if r["phase"] == 0][:5]
    for a, b in zip(lm, g["lm"]):
        assert abs(a["err"] - b["err"]) / b["err"] < 1e-12
        assert abs(a["mu"] - b["mu"]) / b["mu"] < 1e-12
        assert abs(a["rho"] - b["rho"]) < 1e-9
    assert abs(P.get("finalErr") - g["final_err"]) / g["final_err"] < 1e-9
    assert [int(r["err"]) for r in tr if r["phase"] == 2] == g["cholmod_scalar_blocks"]
    k = "54" if key == "54KD" else key
    assert abs(P.get("finalErr") - SURVEY_FINAL[k][0]) / SURVEY_FINAL[k][0] < 1e-9
    assert int(P.get("itno")) == SURVEY_FINAL[k][1]
    if k in SURVEY_LM:
        for a, b in zip(lm, SURVEY_LM[k]):
            assert abs(a["err"] - b) / b < 1e-12
    P.close()


def test_variant_P_equals_variant_R_in_LM_phase():
    """replacing the explicit inverse by potrf + potrs is parity-safe"""
    prob = load("7")
    A = oracle.Problem(prob); B = oracle.Problem(prob)
    B.set("use_explicit_inverse", 0)
    A.solve(); B.solve()
    la = [r for r in A.trace() if r["phase"] == 0][:5]
    lb = [r for r in B.trace() if r["phase"] == 0][:5]
    for x, y in zip(la, lb):
        assert abs(x["err"] - y["err"]) / x["err"] < 1e-13
    assert pattern(A.trace()) == pattern(B.trace())
    assert abs(A.get("finalErr") - B.get("finalErr")) / A.get("finalErr") < 1e-12
    A.close(); B.close()


def test_spd_inverse_chain_is_an_inverse():
    rng = np.random.default_rng(3)
    N = 18
    M = rng.normal(size=(N, N))
    S = M @ M.T + N * np.eye(N)
    L = oracle.lib()
    mat = S.copy()
    aux = np.zeros(3 * N)
    ret = L.orc_SPDinv(oracle._d(mat), oracle._d(aux), N)
    assert ret == 0.0
    assert relerr(mat @ S, np.eye(N)) < 1e-12
    bad = S.copy(); bad[4, 4] = -1.0
    assert L.orc_SPDinv(oracle._d(bad), oracle._d(aux), N) == 1.0


def test_cholmod_on_spd_matrix_is_plain_cholesky():
    """a comfortably positive-definite matrix takes the block fast path everywhere and E ~ 0"""
    rng = np.random.default_rng(5)
    N = 12
    M = rng.normal(size=(N, N))
    S = M @ M.T + 5 * N * np.eye(N)
    L = oracle.lib()
    mat = S.copy(); aux = np.zeros(3 * N); dinv = np.zeros(3 * N); diag = np.zeros(N)
    delta, beta = oracle.C.c_double(), oracle.C.c_double()
    L.orc_get_delta_beta(oracle._d(mat), N, oracle.C.byref(delta), oracle.C.byref(beta))
    ns = oracle.C.c_int()
    L.orc_cholmod_blk(oracle._d(mat), oracle._d(aux), oracle._d(dinv), oracle._d(diag), N, beta.value, delta.value, oracle.C.byref(ns))
    assert ns.value == 0
    Lf = np.tril(mat)
    assert relerr(Lf @ Lf.T, S) < 1e-12
    L.orc_cholmod_E(oracle._d(mat), oracle._d(diag), N)
    assert np.max(np.abs(diag)) < 1e-9 * np.max(np.diag(S))


@pytest.mark.parametrize("kind,N", [("indefinite", 12), ("indefinite", 30), ("rank-deficient", 42), ("negative definite", 24), ("one bad pivot", 36)])
def test_cholmod_defining_properties_on_non_positive_matrices(kind, N):
    """cholmod_blk.cl cannot be compiled here, so its restatement (orc_chol.c) is pinned by what the algorithm DEFINES (SURVEY App. A.4,
    PSBA/cl_cholmod.cpp:25-202): L L^T = S + diag(E) with the off-diagonal part reproduced to rounding, E >= 0, S + diag(E) positive
    (semi)definite, the modified path taken exactly when S is not safely positive definite, and E = what compute_cholmod_E returns."""
    rng = np.random.default_rng(N)
    M = rng.normal(size=(N, N))
    if kind == "indefinite":
        S = M + M.T
    elif kind == "rank-deficient":
        B = rng.normal(size=(N, N - 7))                   # 7 = the gauge null space of a bundle-adjustment camera system
        S = B @ B.T
    elif kind == "negative definite":
        S = -(M @ M.T) - np.eye(N)
    else:
        S = M @ M.T + 1e-12 * np.eye(N)
        S[3, 3] -= 1.5 * S[3, 3]
    L = oracle.lib()
    mat = S.copy(); aux = np.zeros(3 * N); dinv = np.zeros(3 * N); E = np.zeros(N)
    delta, beta = oracle.C.c_double(), oracle.C.c_double()
    L.orc_get_delta_beta(oracle._d(mat), N, oracle.C.byref(delta), oracle.C.byref(beta))
    assert delta.value > 0 and beta.value > 0
    ns = oracle.C.c_int()
    L.orc_cholmod_blk(oracle._d(mat), oracle._d(aux), oracle._d(dinv), oracle._d(E), N, beta.value, delta.value, oracle.C.byref(ns))
    Lf = np.tril(mat)
    L.orc_cholmod_E(oracle._d(mat), oracle._d(E), N)
    assert ns.value >= 1                                    # at least one 3-column block left the fast path
    scale = np.abs(S).max()
    R = Lf @ Lf.T - S
    assert np.abs(R - np.diag(np.diag(R))).max() < 1e-13 * scale
    assert np.abs(np.diag(R) - E).max() < 1e-12 * max(scale, np.abs(E).max())
    assert E.min() > -1e-12 * scale
    assert np.linalg.eigvalsh(S + np.diag(E)).min() > -1e-12 * max(scale, np.abs(E).max())
    if kind != "rank-deficient":
        assert E.max() > 0.1 * abs(np.linalg.eigvalsh(S).min())      # a real shift, not rounding


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("key", ["7", "54"])
def test_restatement_vs_reference_kernels_stagewise(key):
    prob = load(key, kind="reference")
    A = oracle.Problem(prob, kind="reference")
    B = oracle.Problem(prob, kind="restatement")
    for X in (A, B):
        X.call("exQT"); X.call("jacobiQT"); X.call("U", 1); X.call("V", 1); X.call("Wblks", 1); X.call("g", 1)
    for b in ("ex", "JA", "JB", "U", "V", "W", "g", "UVdiag"):
        assert relerr(B.buf(b), A.buf(b)) < 1e-12, b
    mu = 1e-3 * float(A.buf("UVdiag").max())
    for X in (A, B):
        X.call("update_UV", mu); X.call("Vinv"); X.call("Yblks"); X.call("S"); X.call("ea")
    for b in ("V", "Y", "S"):
        assert relerr(B.buf(b), A.buf(b)) < 1e-11, b
    assert relerr(B.buf("eab")[:A.N], A.buf("eab")[:A.N]) < 1e-11
    for X in (A, B):
        assert X.call("SPDinv") == 0.0
        X.call("matVec"); X.call("eb"); X.call("dpb"); X.call("newp")
    assert relerr(B.buf("dp"), A.buf("dp")) < 1e-8
    ca, cb = A.call("exQT_new"), B.call("exQT_new")
    assert abs(ca - cb) / ca < 1e-10
    ja, jb = A.call("Jmultiply_g"), B.call("Jmultiply_g")
    assert abs(ja - jb) / ja < 1e-12
    A.close(); B.close()
