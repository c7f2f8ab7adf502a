"""Extended camera model: lens distortion of the sba "varKD" camera and image-point covariances.
The reference parses both (data/54camsvarKD.txt columns 6-10 through quat2vec, PSBA/misc.cpp:27-29; covimgpts in
PSBA/readparams.cpp:272-283, 380-413) and none of its kernels uses them, so parity is pinned differently:
  * at kc = 0 / Sigma = I the extended code must reproduce the reference model (oracle and engine);
  * at kc != 0 the analytic Jacobian is checked against finite differences of the residual (oracle, CPU);
  * the engine is then checked against the oracle stage by stage and over an LM run (GPU)."""
import numpy as np
import pytest

import oracle
import psba_b200
from util import dataset_paths, pattern, relerr


def _ext_inputs(prob, seed=1):
    rng = np.random.default_rng(seed)
    kc = rng.standard_normal((prob["m"], 5)) * np.array([2e-2, 2e-3, 1e-3, 1e-3, 2e-4])
    a = 0.5 + rng.random(prob["o"]); b = 0.5 + rng.random(prob["o"]); r = 0.6 * (2 * rng.random(prob["o"]) - 1)
    cov = np.stack([a * a, r * a * b, r * a * b, b * b], 1)                 # SPD 2x2, row-major (FULLCOV)
    return kc, cov


def _weights(cov):
    """W = L^-1 with Sigma = L L^T (lower Cholesky): rows (w00, w10, w11)"""
    s00, s01, s11 = cov[:, 0], cov[:, 1], cov[:, -1]
    l00 = np.sqrt(s00); l10 = s01 / l00; l11 = np.sqrt(s11 - l10 * l10)
    return np.stack([1 / l00, -l10 / (l00 * l11), 1 / l11], 1)


def test_oracle_extended_model_reduces_to_the_reference_model():
    prob = oracle.read_sba(*dataset_paths("54"))
    O = oracle.Problem(prob)
    c0 = O.call("exQT"); O.call("jacobiQT")
    JA0, JB0 = O.buf("JA").copy(), O.buf("JB").copy()
    O.set_ext(kc=np.zeros((prob["m"], 5)), wgt=np.tile([1.0, 0.0, 1.0], (prob["o"], 1)))
    c1 = O.call("exQT"); O.call("jacobiQT")
    assert abs(c1 - c0) / c0 < 1e-13
    assert relerr(O.buf("JA"), JA0) < 1e-13 and relerr(O.buf("JB"), JB0) < 1e-13
    O.close()


def test_oracle_extended_jacobian_matches_finite_differences():
    prob = oracle.read_sba(*dataset_paths("54"))
    kc, cov = _ext_inputs(prob)
    O = oracle.Problem(prob)
    O.set_ext(kc=kc, wgt=_weights(cov))
    O.call("exQT"); O.call("jacobiQT")
    JA, JB = O.buf("JA").copy(), O.buf("JB").copy()
    cams, pts = O.buf("cams"), O.buf("pts")
    h = 1e-6
    for arr, J, ncol in ((cams, JA, 6), (pts, JB, 3)):
        for col in range(ncol):
            arr[:, col] += h; O.call("exQT"); ep = O.buf("ex").copy()
            arr[:, col] -= 2 * h; O.call("exQT"); em = O.buf("ex").copy()
            arr[:, col] += h
            fd = -(ep - em) / (2 * h)                                     # e = W (measured - projected)
            assert relerr(fd, J[:, :, col]) < 1e-6
    O.close()


def test_reader_hands_out_distortion_and_covariances(tmp_path):
    c, p, cnp = dataset_paths("54KD")
    a = psba_b200.read_sba(c, p, cnp, ext=True)
    assert a["kc"].shape == (54, 5) and not np.any(a["kc"]) and a["cov"] is None      # the shipped file: kc = 0, no covariances
    cams = tmp_path / "c.txt"; pts = tmp_path / "p.txt"
    cams.write_text("1000 0 0 1 0  0.01 0.002 0.003 0.004 0.005  1 0 0 0  0 0 5\n900 1 2 1 0  -0.01 0 0 0 0  0.9 0.1 0 0  1 2 3\n")
    pts.write_text("0 0 1 2 0 10 20 4 1 1 9 1 30 40 1 0 0 1\n1 1 2 1 1 5 6 2 0.5 0.5 3\n")     # FULLCOV: x y s00 s01 s10 s11
    b = psba_b200.read_sba(str(cams), str(pts), 16, ext=True)
    assert np.array_equal(b["kc"], [[0.01, 0.002, 0.003, 0.004, 0.005], [-0.01, 0, 0, 0, 0]])
    assert np.array_equal(b["cov"], [[4, 1, 1, 9], [1, 0, 0, 1], [2, 0.5, 0.5, 3]])
    assert b["impts"].tolist() == [[10, 20], [30, 40], [5, 6]]
    pts.write_text("0 0 1 2 0 10 20 4 1 9 1 30 40 1 0 1\n1 1 2 1 1 5 6 2 0.5 3\n")               # TRICOV: x y s00 s01 s11
    b = psba_b200.read_sba(str(cams), str(pts), 16, ext=True)
    assert np.array_equal(b["cov"], [[4, 1, 9], [1, 0, 1], [2, 0.5, 3]])


@pytest.mark.gpu
def test_extended_kernels_reproduce_the_plain_model(monkeypatch):
    """the extended kernels forced on with kc = 0 (PSBA_FORCE_EXT) against the default kernels: same products, same LM run"""
    prob = psba_b200.read_sba(*dataset_paths("54"))
    runs = []
    for force in ("0", "1"):
        monkeypatch.setenv("PSBA_FORCE_EXT", force)
        G = psba_b200.PSBA(prob)
        G.set_distortion(np.zeros((prob["m"], 5)))
        cost = G.compute_exQT(); G.compute_jacobiQT()
        U, V, W, g = G.compute_U(1.0), G.compute_V(1.0), G.compute_Wblks(1.0), G.compute_g(1.0)
        G.close()
        G = psba_b200.PSBA(prob)
        G.set_distortion(np.zeros((prob["m"], 5)))
        r = G.solve()
        runs.append((cost, U, V, W, g, r, pattern(G.trace())))
        G.close()
    a, b = runs
    assert abs(a[0] - b[0]) / a[0] < 1e-13
    for q in range(1, 5):
        assert relerr(b[q], a[q]) < 1e-12
    assert a[6] == b[6] and a[5]["itno"] == b[5]["itno"] and abs(a[5]["finalErr"] - b[5]["finalErr"]) / a[5]["finalErr"] < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["kc", "cov", "both"])
def test_distortion_and_covariances_against_oracle(what):
    prob = psba_b200.read_sba(*dataset_paths("54"))
    kc, cov = _ext_inputs(prob)
    if what == "kc": cov = None
    if what == "cov": kc = None
    O = oracle.Problem(prob); G = psba_b200.PSBA(prob)
    O.set_ext(kc=kc, wgt=None if cov is None else _weights(cov))
    G.set_distortion(kc); G.set_covariances(cov)
    c_o = O.call("exQT"); c_g, ex = G.compute_exQT(want=True)
    assert abs(c_g - c_o) / c_o < 1e-12 and relerr(ex, O.buf("ex")) < 1e-12
    O.call("jacobiQT")
    JA, JB = G.compute_jacobiQT(want=True)
    assert relerr(JA, O.buf("JA")) < 1e-11 and relerr(JB, O.buf("JB")) < 1e-11
    O.call("U", 1); O.call("V", 1); O.call("Wblks", 1); O.call("g", 1)
    assert relerr(G.compute_U(1.0), O.buf("U")) < 1e-11 and relerr(G.compute_V(1.0), O.buf("V")) < 1e-11
    assert relerr(G.compute_Wblks(1.0), O.buf("W")) < 1e-11 and relerr(G.compute_g(1.0), O.buf("g")) < 1e-11
    mu = 1e-3 * float(np.max(O.buf("UVdiag")))
    O.call("update_UV", mu); O.call("Vinv"); O.call("Yblks"); O.call("S"); O.call("ea")
    G.update_UV(mu); G.compute_Vinv()
    assert relerr(np.tril(G.compute_S()), np.tril(O.buf("S"))) < 1e-10
    assert G.SPDinv() == 0.0 and O.call("SPDinv") == 0.0
    G.matVec_mul(); O.call("matVec"); O.call("eb"); O.call("dpb"); O.call("newp")
    G.compute_eb(); assert relerr(G.compute_dpb(), O.buf("dp")) < 1e-7
    G.compute_newp()
    cn_o, cn_g = O.call("exQT_new"), G.compute_exQT(psba_b200.PARAMS_NEW)
    assert abs(cn_g - cn_o) / cn_o < 1e-9
    jg_o, jg_g = O.call("Jmultiply_g"), G.compute_Jmultiply(psba_b200.VEC_G)
    assert abs(jg_g - jg_o) / jg_o < 1e-11
    G.restore_UVdiag(); G.close()
    # whole LM phase + the fused candidate evaluation of the back-substitution kernel
    G = psba_b200.PSBA(prob)
    G.set_distortion(kc); G.set_covariances(cov)
    assert O.levmar() == G.levmar()[0]
    to, tg = O.trace(), G.trace()
    assert pattern(tg) == pattern(to) and len(to) >= 3
    for a, b in zip(to, tg):
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
    G.close(); O.close()


@pytest.mark.gpu
def test_covariance_must_be_positive_definite():
    prob = psba_b200.read_sba(*dataset_paths("7"))
    G = psba_b200.PSBA(prob)
    cov = np.tile([1.0, 2.0, 2.0, 1.0], (prob["o"], 1))
    with pytest.raises(ValueError):
        G.set_covariances(cov)
    G.close()
